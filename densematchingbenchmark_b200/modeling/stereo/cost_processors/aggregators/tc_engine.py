"""Routing of the PSM/Acf trunk onto the tcgen05 kernels (csrc/conv3d_tc.cu).

Activations stay in the blocked channels-last 16-bit layout [B][C/8][D][H][W][8] between layers.
`precision` selects the element format and whether values travel as a (hi, lo) split pair:
  'fp16x3' : IEEE-half split pair, 3 MMAs per product (hi*hi + hi*lo + lo*hi), fp32 accumulate --
             ~21 significant bits, the mode that meets the 1e-3 px tolerance vs the fp32 reference;
             values must stay below the fp16 range (65504) -- true for BatchNorm'ed cost volumes
  'bf16x3' : bfloat16 split pair (~16 bits, full fp32 exponent range)
  'fp16' / 'bf16' : single 16-bit plane, 1 MMA per product (fastest, 11 / 8 significant bits)
BatchNorm is folded (eval mode) exactly as on the direct path; packed weight blobs are cached per
layer until a parameter changes."""
import torch

from ..... import _cabi as C
from .....ops import functional as F_
from ...layers.basic_layers import FusedConvUnit


PRECISIONS = {          # name -> (split, fp16)
    "fp16x3": (True, True),
    "bf16x3": (True, False),
    "fp16": (False, True),
    "bf16": (False, False),
}


class Blocked(object):
    """A [B,C,D,H,W] activation held as blocked 16-bit planes."""
    __slots__ = ("hi", "lo", "B", "C", "dims", "fp16")

    def __init__(self, hi, lo, B, Cc, dims, fp16):
        self.hi, self.lo, self.B, self.C, self.dims, self.fp16 = hi, lo, B, Cc, tuple(dims), fp16

    @property
    def split(self):
        return self.lo is not None

    @staticmethod
    def empty(B, Cc, dims, split, fp16, device):
        n = B * Cc * dims[0] * dims[1] * dims[2]
        dt = torch.float16 if fp16 else torch.bfloat16
        hi = torch.empty(n, dtype=dt, device=device)
        lo = torch.empty(n, dtype=dt, device=device) if split else None
        return Blocked(hi, lo, B, Cc, dims, fp16)

    @staticmethod
    def from_ncdhw(x, split, fp16=False):
        x = C.f32(x)
        B, Cc, D, H, W = x.shape
        out = Blocked.empty(B, Cc, (D, H, W), split, fp16, x.device)
        C.call("dmb_b200_ncdhw_to_blocked", C.ptr(x), C.ptr(out.hi), C.ptr(out.lo), B, Cc, D, H, W,
               1 if fp16 else 0, C.stream(x.device))
        return out

    def to_ncdhw(self):
        D, H, W = self.dims
        y = torch.empty(self.B, self.C, D, H, W, dtype=torch.float32, device=self.hi.device)
        C.call("dmb_b200_blocked_to_ncdhw", C.ptr(self.hi), C.ptr(self.lo), C.ptr(y), self.B, self.C, D, H, W,
               1 if self.fp16 else 0, C.stream(y.device))
        return y


def tc_available():
    try:
        return bool(C.load().dmb_b200_conv3d_tc_available())
    except Exception:
        return False


def tc_supported(trunk, raw_cost):
    if not raw_cost.is_cuda or raw_cost.dim() != 5:
        return False
    if trunk.in_planes % 32 != 0 or not tc_available():
        return False
    B, Cc, D, H, W = raw_cost.shape
    return D % 4 == 0 and H % 4 == 0 and W % 4 == 0


def _blob(layer, split, fp16):
    """(packed tcgen05 weight blob, folded bias, Cin, Cout, scale) of a FusedConvUnit or a bare
    nn.Conv3d, cached until the layer's parameters change."""
    cache = layer.__dict__.setdefault("_dmb_b200_tc_cache", {})
    mode = (split, fp16)
    if isinstance(layer, FusedConvUnit):
        w, b = layer.folded()                 # same tensor objects while the parameters are unchanged
        key = (id(w),) + mode
    else:
        tensors = [layer.weight, layer.bias]
        key = tuple((t.data_ptr(), t._version) if t is not None else None for t in tensors) + mode
        w = b = None
    hit = cache.get(mode)
    if hit is not None and hit[0] == key:
        return hit[1]
    if w is None:
        w = F_.pack_conv_weight(layer.weight.detach(), False)
        b = layer.bias.detach().float().contiguous() if layer.bias is not None else None
    K3, Cin, Cout = w.shape
    if K3 != 27:
        raise NotImplementedError("the tcgen05 path implements 3x3x3 kernels only")
    scale = 1.0
    if fp16:
        # power-of-two pre-scale: largest |w| lands in [4, 8) so that the `lo` halves stay normal
        wmax = float(w.abs().max())
        if wmax > 0:
            import math
            scale = 2.0 ** max(-14, min(14, math.floor(math.log2(8.0 / wmax))))
    nbytes = C.load().dmb_b200_conv3d_tc_weight_bytes(Cin, Cout, 1 if split else 0)
    blob = torch.empty(nbytes // 2, dtype=torch.float16 if fp16 else torch.bfloat16, device=w.device)
    C.call("dmb_b200_conv3d_tc_pack_weights", C.ptr(w), C.ptr(blob), Cin, Cout, 1 if split else 0,
           1 if fp16 else 0, float(scale), C.stream(w.device))
    val = (blob, b, Cin, Cout, scale, w)      # `w` is kept alive so that id(w) stays unique
    cache[mode] = (key, val)
    return val


def conv_tc(layer, x, residual=None, relu=False, res_f32=None):
    """x: Blocked.  Returns Blocked (Cout % 32 == 0) or a float32 [B,1,D,H,W] tensor (Cout == 1)."""
    blob, bias, Cin, Cout, scale, _ = _blob(layer, x.split, x.fp16)
    if Cin != x.C:
        raise ValueError("layer expects %d input channels, activation has %d" % (Cin, x.C))
    D, H, W = x.dims
    dev = x.hi.device
    fp16 = 1 if x.fp16 else 0
    if Cout == 1:
        y = torch.empty(x.B, 1, D, H, W, dtype=torch.float32, device=dev)
        C.call("dmb_b200_conv3d_tc", C.ptr(x.hi), C.ptr(x.lo), Cin, C.ptr(blob), float(scale), C.ptr(bias),
               None, None, None, None, 1, C.ptr(y), C.ptr(res_f32), x.B, D, H, W, 1 if relu else 0, fp16, C.stream(dev))
        return y
    y = Blocked.empty(x.B, Cout, x.dims, x.split, x.fp16, dev)
    C.call("dmb_b200_conv3d_tc", C.ptr(x.hi), C.ptr(x.lo), Cin, C.ptr(blob), float(scale), C.ptr(bias),
           C.ptr(residual.hi) if residual is not None else None,
           C.ptr(residual.lo) if residual is not None else None,
           C.ptr(y.hi), C.ptr(y.lo), Cout, None, None, x.B, D, H, W, 1 if relu else 0, fp16, C.stream(dev))
    return y


def _unit(layer, x, residual=None, relu_after=False):
    """FusedConvUnit semantics (own ReLU, or residual + optional ReLU) on the tc kernel."""
    if layer.training and layer.bn is not None:
        raise NotImplementedError("training-mode BatchNorm is not implemented on the CUDA path yet")
    return conv_tc(layer, x, residual, relu=layer._has_relu or relu_after)


def _hourglass_direct(hg, x_blk, presqu, postsqu, out_residual_blk):
    """Interim: stride-2 / transposed layers of the hourglass run on the fp32 direct kernels."""
    x = x_blk.to_ncdhw()
    out, pre, post = hg(x, presqu, postsqu, out_residual=out_residual_blk.to_ncdhw())
    return Blocked.from_ncdhw(out, x_blk.split, x_blk.fp16), pre, post


def run_trunk_tc(trunk, raw_cost):
    """PSMTrunk.trunk() on tcgen05: returns (cost1, cost2, cost3) float32 [B,1,D,H,W]."""
    split, fp16 = PRECISIONS[trunk.precision]
    x = Blocked.from_ncdhw(raw_cost, split, fp16)
    c0 = _unit(trunk.dres0[1], _unit(trunk.dres0[0], x))
    cost0 = _unit(trunk.dres1[1], _unit(trunk.dres1[0], c0), residual=c0)
    out1, pre1, post1 = _hourglass_direct(trunk.dres2, cost0, None, None, cost0)
    out2, pre2, post2 = _hourglass_direct(trunk.dres3, out1, pre1, post1, cost0)
    out3, pre3, post3 = _hourglass_direct(trunk.dres4, out2, pre2, post2, cost0)
    cost1 = conv_tc(trunk.classif1[1], _unit(trunk.classif1[0], out1))
    cost2 = conv_tc(trunk.classif2[1], _unit(trunk.classif2[0], out2), res_f32=cost1)
    cost3 = conv_tc(trunk.classif3[1], _unit(trunk.classif3[0], out3), res_f32=cost2)
    return cost1, cost2, cost3
