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
import os

import torch

from ..... import _cabi as C
from .....ops import functional as F_
from ...layers.basic_layers import FusedConvUnit


# stride-1 / stride-2 layers: kinds 3 / 4 (kw taps merged into the MMA N dimension, the fast kernels) or
# kinds 0 / 1 (one MMA per tap; kept as the simpler reference implementations, DMB_B200_TC_KW_MERGE=0)
KW_MERGE = os.environ.get("DMB_B200_TC_KW_MERGE", "1") != "0"
# transposed layers with 64 input channels: kind 2 (two 32-input-channel passes, the second accumulating in place)
# or kind 5 (one K=64 pass per 16 output channels, every output written once; DMB_B200_TC_DECONV_K64=1).
# Measured (profiles/README.md): kind 5 is SLOWER (conv6: 2 x 141 us against 2 x 115 us) although it moves 40 %
# fewer bytes -- a tcgen05.mma with M=128 costs ~110 cycles however small N is (its A tile streams from shared
# memory), and kind 5 issues twice as many.  Kept, tested and off by default.
DECONV_K64 = os.environ.get("DMB_B200_TC_DECONV_K64", "0") == "1"
# kind 6 (default for 64-input-channel transposed layers): one K = 64 pass per 32 output channels, run as three
# class-group launches that each write their output parity classes once (csrc/conv3d_tc.cu, KIND 6..8) -- the MMA
# count of kind 2 with half its output traffic.  DMB_B200_TC_DECONV_GROUPS=0 falls back to kind 2.
DECONV_GROUPS = os.environ.get("DMB_B200_TC_DECONV_GROUPS", "1") != "0"


def _transposed_kind(cin, cout):
    if DECONV_GROUPS and cin == 64 and cout % 32 == 0:
        return 6
    return 5 if (DECONV_K64 and cin % 64 == 0 and cout % 32 == 0) else 2

# fp16 element formats saturate: |x| > 65504 becomes inf in the hi plane and `lo = x - inf` poisons everything
# downstream.  BatchNorm'ed cost volumes stay orders of magnitude below that, so production runs unchecked;
# DMB_B200_CHECK_FINITE=1 verifies every tensor the tcgen05 path produces or ingests (one host synchronisation per
# layer: a debugging aid) and raises with the layer's name.  'bf16x3' has the full fp32 exponent range.
CHECK_FINITE = os.environ.get("DMB_B200_CHECK_FINITE", "0") == "1"


def check_finite(blocked_or_tensor, what):
    if not CHECK_FINITE:
        return
    t = blocked_or_tensor.hi if isinstance(blocked_or_tensor, Blocked) else blocked_or_tensor
    if not bool(torch.isfinite(t).all()):
        raise FloatingPointError(
            "%s: non-finite values on the tcgen05 path -- activations beyond the IEEE-half range (65504)? "
            "Use precision='bf16x3' (full fp32 exponent range) for un-normalised inputs" % what)


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
    def from_ncdhw(x, split, fp16=False, wsplit=False):
        """wsplit: every row W-parity-split ([even | odd] halves) -- only the stride-2 weight gradient reads that."""
        x = C.f32(x)
        B, Cc, D, H, W = x.shape
        out = Blocked.empty(B, Cc, (D, H, W), split, fp16, x.device)
        C.call("dmb_b200_ncdhw_to_blocked_wsplit" if wsplit else "dmb_b200_ncdhw_to_blocked", C.ptr(x), C.ptr(out.hi),
               C.ptr(out.lo), B, Cc, D, H, W, 1 if fp16 else 0, C.stream(x.device))
        check_finite(out, "ncdhw_to_blocked")
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
    if isinstance(raw_cost, Blocked):
        return True
    if not raw_cost.is_cuda or raw_cost.dim() != 5:
        return False
    return tc_shape_ok(trunk, raw_cost.shape[1], tuple(raw_cost.shape[2:]))


def tc_shape_ok(trunk, channels, dhw):
    if trunk.in_planes % 32 != 0 or channels != trunk.in_planes or not tc_available():
        return False
    return all(n % 4 == 0 for n in dhw)


def _kind_of(layer):
    """0/3: stride-1 conv, 1/4: stride-2 conv, 2/5: stride-2 transposed conv (k3, p1, op1)."""
    conv = layer.conv if isinstance(layer, FusedConvUnit) else layer
    if tuple(conv.kernel_size) != (3, 3, 3) or tuple(conv.padding) != (1, 1, 1) or tuple(conv.dilation) != (1, 1, 1):
        raise NotImplementedError("the tcgen05 path implements 3x3x3, padding 1, dilation 1 only")
    if isinstance(conv, torch.nn.ConvTranspose3d):
        if tuple(conv.stride) != (2, 2, 2) or tuple(conv.output_padding) != (1, 1, 1):
            raise NotImplementedError("transposed conv on tcgen05: stride 2, output_padding 1 only")
        return _transposed_kind(conv.in_channels, conv.out_channels)
    if tuple(conv.stride) == (1, 1, 1):
        return 3 if KW_MERGE else 0
    if tuple(conv.stride) == (2, 2, 2):
        return 4 if KW_MERGE else 1
    raise NotImplementedError("stride %s is not supported on tcgen05" % (tuple(conv.stride),))


def _blob(layer, split, fp16):
    """(packed tcgen05 weight blob, folded bias, Cin, Cout, scale, kind) of a FusedConvUnit or a
    bare nn.Conv3d, cached until the layer's parameters change."""
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
    kind = _kind_of(layer)
    if w is None:
        w = F_.pack_conv_weight(layer.weight.detach(), isinstance(layer, torch.nn.ConvTranspose3d))
        b = layer.bias.detach().float().contiguous() if layer.bias is not None else None
    blob, Cin, Cout, scale = pack_blob(w, kind, split, fp16)
    val = (blob, b, Cin, Cout, scale, kind, w)      # `w` is kept alive so that id(w) stays unique
    cache[mode] = (key, val)
    return val


def weight_scale(w):
    """Power-of-two pre-scale of an fp16 weight blob: the largest |w| lands in [4, 8) so that the `lo` halves stay
    normal.  Reads max|w| back to the host (one synchronisation)."""
    import math
    wmax = float(w.abs().max())
    return 2.0 ** max(-14, min(14, math.floor(math.log2(8.0 / wmax)))) if wmax > 0 else 1.0


def pack_blob(w, kind, split, fp16, scale=None):
    """w: packed fp32 weight [27,Cin,Cout] (ops.functional.pack_conv_weight) -> (tcgen05 blob, Cin, Cout, scale).
    `scale`: a pre-computed weight_scale() -- training re-packs every step and must not synchronise every time;
    None computes it here.  bfloat16 blobs are never scaled (full fp32 exponent range)."""
    K3, Cin, Cout = w.shape
    if not fp16:
        scale = 1.0
    elif scale is None:
        scale = weight_scale(w)
    nbytes = C.load().dmb_b200_conv3d_tc_weight_bytes(Cin, Cout, 1 if split else 0, kind)
    blob = torch.empty(nbytes // 2, dtype=torch.float16 if fp16 else torch.bfloat16, device=w.device)
    C.call("dmb_b200_conv3d_tc_pack_weights", C.ptr(w), C.ptr(blob), Cin, Cout, 1 if split else 0,
           1 if fp16 else 0, float(scale), kind, C.stream(w.device))
    return blob, Cin, Cout, scale


def conv_tc(layer, x, residual=None, relu=False, res_f32=None):
    """x: Blocked.  Returns Blocked (Cout % 32 == 0) or a float32 [B,1,D,H,W] tensor (Cout == 1)."""
    blob, bias, Cin, Cout, scale, kind, _ = _blob(layer, x.split, x.fp16)
    return conv_tc_raw(x, blob, bias, Cin, Cout, scale, kind, residual, relu, res_f32)


def conv_tc_raw(x, blob, bias, Cin, Cout, scale, kind, residual=None, relu=False, res_f32=None):
    if Cin != x.C:
        raise ValueError("layer expects %d input channels, activation has %d" % (Cin, x.C))
    D, H, W = x.dims
    if kind in (1, 4) and (D % 2 or H % 2 or W % 2):
        raise ValueError("stride-2 convolution on tcgen05 needs even extents, got %s" % (x.dims,))
    odims = x.dims if kind in (0, 3) else (tuple(n // 2 for n in x.dims) if kind in (1, 4) else tuple(2 * n for n in x.dims))   # 2, 5, 6: transposed
    dev = x.hi.device
    fp16 = 1 if x.fp16 else 0
    if Cout == 1:
        y = torch.empty(x.B, 1, *odims, dtype=torch.float32, device=dev)
        C.call("dmb_b200_conv3d_tc", C.ptr(x.hi), C.ptr(x.lo), Cin, C.ptr(blob), float(scale), C.ptr(bias),
               None, None, None, None, 1, C.ptr(y), C.ptr(res_f32), x.B, D, H, W, kind, 1 if relu else 0, fp16,
               C.stream(dev))
        check_finite(y, "conv3d_tc (%d->1)" % Cin)
        return y
    if residual is not None and (residual.dims != odims or residual.C != Cout):
        raise ValueError("residual geometry %s x%d does not match the output %s x%d"
                         % (residual.dims, residual.C, odims, Cout))
    y = Blocked.empty(x.B, Cout, odims, x.split, x.fp16, dev)
    C.call("dmb_b200_conv3d_tc", C.ptr(x.hi), C.ptr(x.lo), Cin, C.ptr(blob), float(scale), C.ptr(bias),
           C.ptr(residual.hi) if residual is not None else None,
           C.ptr(residual.lo) if residual is not None else None,
           C.ptr(y.hi), C.ptr(y.lo), Cout, None, None, x.B, D, H, W, kind, 1 if relu else 0, fp16, C.stream(dev))
    check_finite(y, "conv3d_tc (%d->%d, kind %d)" % (Cin, Cout, kind))
    return y


def cached_weight_scale(param, w_packed, refresh=64):
    """weight_scale() of a trainable weight, cached on the Parameter and refreshed every `refresh` in-place updates
    (optimizer steps): the scale only has to keep |w| * scale inside fp16's range with the `lo` halves normal, and
    max|w| drifts slowly -- one host synchronisation per layer every `refresh` steps instead of every step."""
    hit = param.__dict__.get("_dmb_b200_wscale")
    ver = param._version
    if hit is not None and 0 <= ver - hit[0] < refresh:
        return hit[1]
    sc = weight_scale(w_packed)
    param.__dict__["_dmb_b200_wscale"] = (ver, sc)
    return sc


def conv3d_ncdhw_tc(x, w_packed, bias, stride, transposed, precision, residual=None, relu=False, scale=None,
                    x_blocked=None):
    """One 3x3x3 / pad 1 convolution (stride 1 | stride 2 | transposed stride 2 with output_padding 1) of a float32
    NCDHW tensor on the tcgen05 kernels: layout conversion in, conv, layout conversion out.  Used by the training
    path (ops/autograd.py), whose weights change every step -- nothing is cached.  Returns float32 NCDHW."""
    split, fp16 = PRECISIONS[precision]
    kind = (_transposed_kind(w_packed.shape[1], w_packed.shape[2]) if transposed
            else ((3 if KW_MERGE else 0) if stride == 1 else (4 if KW_MERGE else 1)))
    blob, Cin, Cout, scale = pack_blob(w_packed, kind, split, fp16, scale)
    xb = x_blocked if x_blocked is not None else Blocked.from_ncdhw(x, split, fp16)   # (a conversion the caller shares)
    if Cout == 1:
        return conv_tc_raw(xb, blob, bias, Cin, Cout, scale, kind, None, relu, residual)
    rb = Blocked.from_ncdhw(residual, split, fp16) if residual is not None else None
    return conv_tc_raw(xb, blob, bias, Cin, Cout, scale, kind, rb, relu).to_ncdhw()


def wgrad_tc(a, g, stride=1, a_blocked=None, g_blocked=None):
    """Weight gradient of a 3x3x3 / pad 1 convolution of stride 1 or 2 on tcgen05 (csrc/wgrad_tc.cu):
    dw[tap][ca][cg] = sum a[s*v + tap - 1] * g[v].  a, g: float32 NCDHW (or already Blocked bf16 split pairs; with
    stride 2 `a` W-parity-split) -> [27, Ca, Cg] float32.  bfloat16 split pairs: gradients can be arbitrarily small,
    so the fp32 exponent range matters more than the three extra mantissa bits of IEEE half."""
    ab = a_blocked if a_blocked is not None else Blocked.from_ncdhw(a, True, False, wsplit=(stride == 2))
    gb = g_blocked if g_blocked is not None else Blocked.from_ncdhw(g, True, False)
    if ab.dims != tuple(stride * n for n in gb.dims) or ab.B != gb.B or ab.fp16 or gb.fp16 or not (ab.split and gb.split):
        raise ValueError("wgrad_tc: a and g must be bf16 split pairs with dims(a) == stride * dims(g)")
    D, H, W = gb.dims
    dw = torch.zeros(27, ab.C, gb.C, device=ab.hi.device, dtype=torch.float32)
    C.call("dmb_b200_conv3d_wgrad_tc", C.ptr(ab.hi), C.ptr(ab.lo), C.ptr(gb.hi), C.ptr(gb.lo), C.ptr(dw), ab.B, ab.C, gb.C,
           D, H, W, stride, 0, C.stream(dw.device))
    return dw


def wgrad_tc_eligible(a, g, ksize, stride, pad, pad_g=False):
    """a: the tensor the taps slide over (the conv input; for a transposed conv the output gradient), g the other one.
    pad_g: g will be zero-padded to a multiple of 32 channels by the caller."""
    return (stride in (1, 2) and pad == 1 and tuple(ksize) == (3, 3, 3) and a.is_cuda
            and a.shape[1] % 32 == 0 and (pad_g or g.shape[1] % 32 == 0)
            and tuple(a.shape[2:]) == tuple(stride * n for n in g.shape[2:]) and tc_available())


def conv3d_tc_eligible(x, Cin, Cout, ksize, stride, pad, transposed, opad, out_dims=None):
    """Geometry the tcgen05 kernels implement (see _kind_of) on a CUDA tensor of an sm_100 device."""
    if not x.is_cuda or tuple(ksize) != (3, 3, 3) or pad != 1 or Cin % 32 or not (Cout % 32 == 0 or Cout == 1):
        return False
    dims = tuple(x.shape[2:])
    if transposed:
        if stride != 2 or Cout == 1:
            return False
        want = tuple(2 * n for n in dims)
        if out_dims is not None:
            if tuple(int(v) for v in out_dims) != want:
                return False
        elif opad != 1:
            return False
    elif stride == 2:
        if any(n % 2 for n in dims) or Cout == 1:
            return False
    elif stride != 1:
        return False
    return tc_available()


# 32->1 classifier heads: fused into the epilogue of the preceding 32->32 layer + a 27-term gather
# (DMB_B200_TC_FUSED_HEAD=0: the head as its own tensor-core launch, kept for A/B)
FUSED_HEAD = os.environ.get("DMB_B200_TC_FUSED_HEAD", "1") != "0"


def _head_weight(conv):
    """Conv3d(32,1,3,1,1,bias=False) weight as [27][32] float32, cached until the parameter changes."""
    cache = conv.__dict__.setdefault("_dmb_b200_head_cache", {})
    key = (conv.weight.data_ptr(), conv.weight._version)
    if cache.get("key") != key:
        w = F_.pack_conv_weight(conv.weight.detach(), False)           # [27, 32, 1]
        cache["w"] = w.reshape(w.shape[0], w.shape[1]).contiguous()
        cache["key"] = key
    return cache["w"]


def classif_head(seq, x, res_f32=None):
    """classifN = Sequential(conv3d_bn_relu(32,32), Conv3d(32,1)) (+ the previous cost) -> float32 [B,1,D,H,W]."""
    unit, conv = seq[0], seq[1]
    fusable = (FUSED_HEAD and KW_MERGE and x.C == 32 and conv.bias is None and conv.out_channels == 1
               and conv.in_channels == 32 and _kind_of(unit) == 3 and _kind_of(conv) == 3)
    if not fusable:
        return conv_tc(conv, _unit(unit, x), res_f32=res_f32)
    if unit.training and unit.bn is not None:
        raise NotImplementedError("training-mode BatchNorm is not implemented on the tcgen05 path")
    blob, bias, Cin, Cout, scale, kind, _ = _blob(unit, x.split, x.fp16)
    D, H, W = x.dims
    dev = x.hi.device
    with torch.cuda.device(dev):          # the schedule (depth segments -> spill planes) depends on the device's SM count
        nfloats = C.load().dmb_b200_conv3d_tc_head_floats(x.B, D, H, W)
    taps = torch.empty(nfloats, dtype=torch.float32, device=dev)
    C.call("dmb_b200_conv3d_tc_head", C.ptr(x.hi), C.ptr(x.lo), C.ptr(blob), float(scale), C.ptr(bias),
           C.ptr(_head_weight(conv)), C.ptr(taps), x.B, D, H, W, 1 if unit._has_relu else 0, 1 if x.fp16 else 0,
           C.stream(dev))
    y = torch.empty(x.B, 1, D, H, W, dtype=torch.float32, device=dev)
    C.call("dmb_b200_head_gather", C.ptr(taps), C.ptr(res_f32), C.ptr(y), x.B, D, H, W, C.stream(dev))
    check_finite(y, "classifier head")
    return y


def _unit(layer, x, residual=None, relu_after=False):
    """FusedConvUnit semantics (own ReLU, or residual + optional ReLU) on the tc kernel."""
    if layer.training and layer.bn is not None:
        raise NotImplementedError("training-mode BatchNorm is not implemented on the CUDA path yet")
    return conv_tc(layer, x, residual, relu=layer._has_relu or relu_after)


def _hourglass(hg, x, presqu, postsqu, out_residual):
    """Hourglass.forward (utils/hourglass.py:62-86) entirely on tcgen05, Blocked in / Blocked out."""
    out = _unit(hg.conv1, x)                                               # stride 2
    pre = _unit(hg.conv2, out, residual=postsqu, relu_after=True)
    out = _unit(hg.conv3, pre)                                             # stride 2
    out = _unit(hg.conv4, out)
    post = _unit(hg.conv5, out, residual=presqu if presqu is not None else pre, relu_after=True)   # transposed
    out = _unit(hg.conv6, post, residual=out_residual)                     # transposed
    return out, pre, post


def cat_volume_blocked(reference_fm, target_fm, max_disp, start_disp, dilation, precision):
    """cat_fms (cat_fms.py:7-48) written straight into the trunk's blocked 16-bit layout: the fp32
    NCDHW volume (401 MB at 544x960) and its conversion pass are never materialised."""
    split, fp16 = PRECISIONS[precision]
    l, r = C.f32(reference_fm), C.f32(target_fm)
    B, Cc, H, W = l.shape
    idx = F_.disp_indices(max_disp, start_disp, dilation)
    out = Blocked.empty(B, 2 * Cc, (len(idx), H, W), split, fp16, l.device)
    C.call("dmb_b200_cat_volume_blocked", C.ptr(l), C.ptr(r), C.ptr(out.hi), C.ptr(out.lo), B, Cc, H, W,
           C.int_array(idx), len(idx), 1 if fp16 else 0, C.stream(l.device))
    check_finite(out, "cat_volume_blocked")
    return out


def run_trunk_tc(trunk, raw_cost):
    """PSMTrunk.trunk() on tcgen05: returns (cost1, cost2, cost3) float32 [B,1,D,H,W].
    `raw_cost` is the NCDHW float32 volume or an already Blocked one."""
    split, fp16 = PRECISIONS[trunk.precision]
    x = raw_cost if isinstance(raw_cost, Blocked) else Blocked.from_ncdhw(raw_cost, split, fp16)
    c0 = _unit(trunk.dres0[1], _unit(trunk.dres0[0], x))
    cost0 = _unit(trunk.dres1[1], _unit(trunk.dres1[0], c0), residual=c0)
    out1, pre1, post1 = _hourglass(trunk.dres2, cost0, None, None, cost0)
    out2, pre2, post2 = _hourglass(trunk.dres3, out1, pre1, post1, cost0)
    out3, pre3, post3 = _hourglass(trunk.dres4, out2, pre2, post2, cost0)
    cost1 = classif_head(trunk.classif1, out1)
    cost2 = classif_head(trunk.classif2, out2, res_f32=cost1)
    cost3 = classif_head(trunk.classif3, out3, res_f32=cost2)
    return cost1, cost2, cost3


# ---- the other GeneralizedStereoModel aggregators on the same kernels --------------------
def blocked_cat(a, b):
    """torch.cat([a, b], dim=1) of two Blocked activations: channel blocks are the second-slowest axis of the layout."""
    if a.dims != b.dims or a.B != b.B or a.fp16 != b.fp16 or a.split != b.split:
        raise ValueError("blocked_cat: geometry / format mismatch")
    n = a.dims[0] * a.dims[1] * a.dims[2] * 8

    def cat(x, y):
        return torch.cat([x.view(a.B, a.C // 8, n), y.view(b.B, b.C // 8, n)], dim=1).reshape(-1)

    return Blocked(cat(a.hi, b.hi), cat(a.lo, b.lo) if a.split else None, a.B, a.C + b.C, a.dims, a.fp16)


def blocked_add(a, b):
    if a.dims != b.dims or a.B != b.B or a.C != b.C or a.fp16 != b.fp16 or a.split != b.split:
        raise ValueError("blocked_add: geometry / format mismatch")
    y = Blocked.empty(a.B, a.C, a.dims, a.split, a.fp16, a.hi.device)
    C.call("dmb_b200_blocked_add", C.ptr(a.hi), C.ptr(a.lo), C.ptr(b.hi), C.ptr(b.lo), C.ptr(y.hi), C.ptr(y.lo),
           a.hi.numel() // 8, 1 if a.fp16 else 0, C.stream(a.hi.device))
    return y


def gc_shape_ok(agg, raw_cost):
    """GCAggregator on tcgen05: every channel count a multiple of 32 (in_planes = 64) and four clean halvings."""
    return (raw_cost.is_cuda and raw_cost.dim() == 5 and agg.in_planes % 64 == 0 and raw_cost.shape[1] == agg.in_planes
            and all(n % 16 == 0 for n in raw_cost.shape[2:]) and tc_available())


def run_gc_tc(agg, raw_cost, precision="fp16x3"):
    """GCAggregator.forward (aggregators/GCNet.py:73-120) up to layer36 on the tcgen05 kernels (stride-1, stride-2 and
    transposed 3x3x3 units in blocked 16-bit split pairs; concatenations and post-ReLU skip additions in the blocked
    layout); returns the float32 NCDHW input of layer37 (a 32 -> 1 transposed convolution: not a tensor-core shape)."""
    split, fp16 = PRECISIONS[precision]
    v18 = Blocked.from_ncdhw(raw_cost, split, fp16)
    v19 = _unit(agg.layer19, v18)
    v20 = _unit(agg.layer20, v19)
    v21 = _unit(agg.layer21, blocked_cat(v18, v20))
    v22 = _unit(agg.layer22, v21)
    v23 = _unit(agg.layer23, v22)
    v24 = _unit(agg.layer24, blocked_cat(v21, v23))
    v25 = _unit(agg.layer25, v24)
    v26 = _unit(agg.layer26, v25)
    v27 = _unit(agg.layer27, blocked_cat(v24, v26))
    v28 = _unit(agg.layer28, v27)
    v29 = _unit(agg.layer29, v28)
    v30 = _unit(agg.layer30, blocked_cat(v27, v29))
    v31 = _unit(agg.layer31, v30)
    v32 = _unit(agg.layer32, v31)
    v33 = _unit(agg.layer33, v32)
    v34 = _unit(agg.layer34, blocked_add(v33, v29))
    v35 = _unit(agg.layer35, blocked_add(v34, v26))
    v36 = _unit(agg.layer36, blocked_add(v35, v23))
    return blocked_add(v36, v20).to_ncdhw()


def run_stereonet_tc(agg, raw_cost, precision="fp16x3"):
    """StereoNetAggregator.forward (aggregators/StereoNet.py:41-55): `num` conv3d_bn_relu(32, 32) units and the
    Conv3d(32, 1) head, all on the stride-1 tcgen05 kernel; returns [B,1,D,H,W] float32."""
    split, fp16 = PRECISIONS[precision]
    x = Blocked.from_ncdhw(raw_cost, split, fp16)
    for layer in agg.classify:
        x = _unit(layer, x)
    return conv_tc(agg.lastconv, x)
