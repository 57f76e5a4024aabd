"""Training-mode autograd of the cost-volume path: every Function below is a forward AND a backward
made of the library's own kernels (csrc/train.cu, csrc/conv3d_direct.cu, csrc/regress.cu,
csrc/volumes.cu) -- torch.autograd only strings them together.

Reference behaviour being reproduced: autograd over `nn.Sequential(Conv3d|ConvTranspose3d,
BatchNorm3d (batch statistics, running-stat update), ReLU)` (layers/basic_layers.py:68-216), the
trilinear / ConvTranspose3d(1,1,8,4,2) cost upsampling (aggregators/PSMNet.py:75-88,
AcfNet.py:55-57,81-83), FasterSoftArgmin / SoftArgmin (disp_predictors/*.py) and cat_fms
(cost_processors/utils/cat_fms.py:7-48).  Synchronised BatchNorm (dmb/apis/train.py:95-97 converts the
model with apex) = all-reducing the raw per-channel sums between the two passes of each direction.
"""
import os

import torch
import torch.distributed as dist

from .. import _cabi as C
from . import functional as F_

# Training convolutions (forward and input gradient) run on the tcgen05 kernels whenever the layer geometry
# allows it (3x3x3, pad 1, stride 1 / 2 / transposed 2, Cin % 32 == 0, Cout % 32 == 0 or 1); everything else
# uses the fp32 SIMT kernel.  Forward: split IEEE-half arithmetic (fp32-grade).  Input
# gradients: split bfloat16 (full fp32 exponent range -- gradients can be arbitrarily small -- at ~2^-16
# relative precision).  DMB_B200_TRAIN_TC=0 forces the SIMT kernels (A/B and the reference for the tests).
TRAIN_TC = os.environ.get("DMB_B200_TRAIN_TC", "1") != "0"
TRAIN_TC_FWD = "fp16x3"
TRAIN_TC_BWD = "bf16x3"


def _conv(x, w_packed, bias, ksize, stride, pad, transposed, opad, residual, relu, out_dims=None, precision=None,
          param=None, x_blocked=None):
    """y = act(conv(x) + bias + residual) for the training path: tcgen05 if eligible, else conv3d_direct.
    `param`: the weight Parameter the packed weight was made from (cache key of its fp16 pre-scale)."""
    if TRAIN_TC and precision is not None:
        from ..modeling.stereo.cost_processors.aggregators import tc_engine as T
        K3, Cin, Cout = w_packed.shape
        scale = None
        if param is not None and T.PRECISIONS[precision][1]:
            scale = T.cached_weight_scale(param, w_packed)
        if transposed and stride == 1 and tuple(ksize) == (3, 3, 3) and pad == 1 and \
                (out_dims is None or tuple(int(v) for v in out_dims) == tuple(x.shape[2:])):
            # a stride-1 transposed convolution is the plain convolution with the taps mirrored
            if T.conv3d_tc_eligible(x, Cin, Cout, ksize, 1, pad, False, 0):
                return T.conv3d_ncdhw_tc(x, w_packed.flip(0).contiguous(), bias, 1, False, precision, residual, relu, scale,
                                         x_blocked)
        elif T.conv3d_tc_eligible(x, Cin, Cout, ksize, stride, pad, transposed, opad, out_dims):
            return T.conv3d_ncdhw_tc(x, w_packed, bias, stride, transposed, precision, residual, relu, scale, x_blocked)
    return F_.conv3d_fused(x, w_packed, bias, ksize, stride, pad, transposed, opad, residual, relu=relu,
                           out_dims=out_dims)


def _sync_world(group):
    if group is None or not dist.is_available() or not dist.is_initialized():
        return 1
    return dist.get_world_size(group if group is not True else None)


def _all_reduce_sums(sums, group):
    # one peer-memory kernel on the current stream when the ranks share a box (csrc/peer_comm.cu), else NCCL
    from ..utils.dist_utils import sum_over_ranks_
    sum_over_ranks_(sums, group)


def _dgrad(dz, weight, transposed, ksize, stride, pad, x_dims, dz_blocked=None):
    """Gradient w.r.t. the conv input = the forward kernel with the weight's roles swapped:
    Conv3d weight [Cout,Cin,k] read as a ConvTranspose3d weight (in=Cout, out=Cin), and vice versa."""
    w = F_.pack_conv_weight(weight.detach(), transposed=not transposed)
    return _conv(dz, w, None, ksize, stride, pad, not transposed, 0, None, False, out_dims=x_dims,
                 precision=TRAIN_TC_BWD, x_blocked=dz_blocked)


# weight gradients of the stride-1 layers on tcgen05 (csrc/wgrad_tc.cu); DMB_B200_TRAIN_TC_WGRAD=0: the SIMT kernel
TRAIN_TC_WGRAD = os.environ.get("DMB_B200_TRAIN_TC_WGRAD", "1") != "0"


def _wgrad_on_tc(x, dz, transposed, ksize, stride, pad):
    if not (TRAIN_TC and TRAIN_TC_WGRAD) or (transposed and stride == 1):
        return False
    from ..modeling.stereo.cost_processors.aggregators import tc_engine as T
    a, g = (dz, x) if transposed else (x, dz)
    if g.shape[1] < 32 and stride == 1:
        # the 32 -> 1 classifier heads (aggregators/PSMNet.py:41-52): g is padded to one 32-channel block with zeros --
        # 31/32 of the MMA columns idle, still ~25x faster than the SIMT kernel's 3.4 ms (profiles/README.md)
        return T.wgrad_tc_eligible(a, g, ksize, stride, pad, pad_g=True)
    return T.wgrad_tc_eligible(a, g, ksize, stride, pad)


def _wgrad(x, dz, transposed, ksize, stride, pad, weight_shape, dz_blocked=None):
    if tuple(ksize) != (3, 3, 3):
        raise NotImplementedError("conv weight gradient: only 3x3x3 kernels are on the training path")
    a, g = (dz, x) if transposed else (x, dz)          # transposed: the roles of input and output swap
    B, Ca = a.shape[:2]
    Cg = g.shape[1]
    if _wgrad_on_tc(x, dz, transposed, ksize, stride, pad):
        from ..modeling.stereo.cost_processors.aggregators import tc_engine as T
        if Cg < 32:                                    # zero-padded to one channel block (see _wgrad_on_tc)
            gp = torch.zeros((B, 32) + tuple(g.shape[2:]), device=g.device, dtype=torch.float32)
            gp[:, :Cg] = g
            dw = T.wgrad_tc(a, gp, stride)[:, :, :Cg]
        else:
            # dz_blocked (shared with the input-gradient kernel) is g for a plain conv; for the transposed conv dz is
            # `a` and needs the W-parity-split layout instead
            dw = T.wgrad_tc(a, g, stride, g_blocked=None if transposed else dz_blocked)
        return dw.permute(2, 1, 0).reshape(weight_shape).contiguous()
    dw = torch.zeros(27, Ca, Cg, device=x.device, dtype=torch.float32)
    C.call("dmb_b200_conv3d_wgrad", C.ptr(a), C.ptr(g), C.ptr(dw), B, Ca, Cg, C.int_array(list(a.shape[2:])),
           C.int_array(list(g.shape[2:])), stride, pad, C.stream(x.device))
    # packed [27][in-role][out-role] -> [out-role][in-role][3][3][3]: Conv3d [Cout,Cin,..]; for the transposed
    # conv out-role = its Cin, in-role = its Cout, i.e. exactly the ConvTranspose3d layout [Cin,Cout,..]
    return dw.permute(2, 1, 0).reshape(weight_shape).contiguous()


class ConvUnitFn(torch.autograd.Function):
    """y = relu?( bn_train?( conv(x, weight) + bias ) + residual? )"""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, residual, cfg):
        x = C.f32(x)
        transposed, ksize, stride, pad, opad = cfg["transposed"], cfg["ksize"], cfg["stride"], cfg["pad"], cfg["opad"]
        relu, bn, group = cfg["relu"], cfg.get("bn"), cfg.get("sync_group")
        w_packed = F_.pack_conv_weight(weight.detach(), transposed)
        b = bias.detach().float().contiguous() if bias is not None else None
        res = C.f32(residual) if residual is not None else None
        ctx.cfg = cfg
        ctx.has_bias = bias is not None
        ctx.has_res = residual is not None
        ctx.x_dims = tuple(x.shape[2:])
        if bn is None:
            y = _conv(x, w_packed, b, ksize, stride, pad, transposed, opad, res, relu, precision=TRAIN_TC_FWD, param=weight)
            ctx.save_for_backward(x, weight, y if relu else None, None, None, None, None)
            return y
        z = _conv(x, w_packed, b, ksize, stride, pad, transposed, opad, None, False, precision=TRAIN_TC_FWD, param=weight)
        B, Co = z.shape[:2]
        S = z.numel() // (B * Co)
        dev = z.device
        sums = torch.zeros(2 * Co, device=dev, dtype=torch.float64)
        C.call("dmb_b200_bn_stats", C.ptr(z), C.ptr(sums), B, Co, S, C.stream(dev))
        world = _sync_world(group)
        if world > 1:
            _all_reduce_sums(sums, group)
        count = float(B * S * world)
        mean, invstd, scale, shift = (torch.empty(Co, device=dev, dtype=torch.float32) for _ in range(4))
        track = bn.track_running_stats and bn.running_mean is not None
        momentum = bn.momentum
        if track:
            bn.num_batches_tracked += 1
            if momentum is None:                              # cumulative moving average
                momentum = 1.0 / float(bn.num_batches_tracked)
        g = gamma.detach().float().contiguous() if gamma is not None else None
        bt = beta.detach().float().contiguous() if beta is not None else None
        C.call("dmb_b200_bn_finalize", C.ptr(sums), count, C.ptr(g), C.ptr(bt), float(bn.eps),
               float(momentum if momentum is not None else 0.0), C.ptr(bn.running_mean if track else None),
               C.ptr(bn.running_var if track else None), C.ptr(mean), C.ptr(invstd), C.ptr(scale), C.ptr(shift), Co,
               C.stream(dev))
        y = torch.empty_like(z)
        C.call("dmb_b200_bn_apply", C.ptr(z), C.ptr(scale), C.ptr(shift), C.ptr(res), C.ptr(y), B, Co, S,
               1 if relu else 0, C.stream(dev))
        ctx.count = count
        ctx.save_for_backward(x, weight, y if relu else None, z, mean, invstd, g)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y_relu, z, mean, invstd, gamma = ctx.saved_tensors
        cfg = ctx.cfg
        transposed, ksize, stride, pad = cfg["transposed"], cfg["ksize"], cfg["stride"], cfg["pad"]
        bn, group = cfg.get("bn"), cfg.get("sync_group")
        dy = C.f32(dy)
        B, Co = dy.shape[:2]
        S = dy.numel() // (B * Co)
        dev = dy.device
        need_x, need_w, need_b, need_g, need_bt, need_res = ctx.needs_input_grad[:6]
        need_res = need_res and ctx.has_res
        dgamma = dbeta = dbias = dres = None
        if bn is not None:
            sums = torch.zeros(2 * Co, device=dev, dtype=torch.float64)
            C.call("dmb_b200_bn_backward_reduce", C.ptr(dy), C.ptr(y_relu), C.ptr(z), C.ptr(mean), C.ptr(invstd),
                   C.ptr(sums), B, Co, S, C.stream(dev))
            if gamma is not None:
                dbeta, dgamma = sums[:Co].float(), sums[Co:].float()        # local sums (DDP averages them later)
            if _sync_world(group) > 1:
                _all_reduce_sums(sums, group)
            dz = torch.empty_like(dy)
            dres = torch.empty_like(dy) if need_res else None
            C.call("dmb_b200_bn_backward_apply", C.ptr(dy), C.ptr(y_relu), C.ptr(z), C.ptr(mean), C.ptr(invstd),
                   C.ptr(gamma), C.ptr(sums), ctx.count, C.ptr(dz), C.ptr(dres), B, Co, S, C.stream(dev))
            if ctx.has_bias and need_b:
                dbias = torch.zeros(Co, device=dev, dtype=torch.float32)    # batch norm removes the mean: exactly 0
        else:
            if y_relu is not None:
                dz = torch.empty_like(dy)
                C.call("dmb_b200_bn_backward_apply", C.ptr(dy), C.ptr(y_relu), None, None, None, None, None, 1.0,
                       C.ptr(dz), None, B, Co, S, C.stream(dev))
            else:
                dz = dy
            dres = dz if need_res else None
            if ctx.has_bias and need_b:
                sums = torch.zeros(2 * Co, device=dev, dtype=torch.float64)
                C.call("dmb_b200_bn_backward_reduce", C.ptr(dz), None, None, None, None, C.ptr(sums), B, Co, S,
                       C.stream(dev))
                dbias = sums[:Co].float()
        dz_blocked = None
        if need_x and need_w and dz.shape[1] % 32 == 0 and not transposed and _wgrad_on_tc(x, dz, transposed, ksize, stride, pad):
            # one bf16 split-pair conversion of dz shared by the input-gradient and the weight-gradient kernels
            from ..modeling.stereo.cost_processors.aggregators import tc_engine as T
            if T.PRECISIONS[TRAIN_TC_BWD] == (True, False):
                dz_blocked = T.Blocked.from_ncdhw(dz, True, False)
        dx = _dgrad(dz, weight, transposed, ksize, stride, pad, ctx.x_dims, dz_blocked) if need_x else None
        dw = _wgrad(x, dz, transposed, ksize, stride, pad, weight.shape, dz_blocked) if need_w else None
        return dx, dw, dbias, (dgamma if need_g else None), (dbeta if need_bt else None), dres, None


class CatVolumeFn(torch.autograd.Function):
    """cat_fms forward + backward (cost_processors/utils/cat_fms.py:7-48)."""

    @staticmethod
    def forward(ctx, left, right, max_disp, start_disp, dilation):
        ctx.args = (max_disp, start_disp, dilation)
        ctx.shape = tuple(left.shape)
        return F_.cat_volume(left, right, max_disp, start_disp, dilation)

    @staticmethod
    def backward(ctx, dvol):
        dvol = C.f32(dvol)
        B, Ch, H, W = ctx.shape
        idx = F_.disp_indices(*ctx.args)
        dl = torch.empty(ctx.shape, device=dvol.device, dtype=torch.float32)
        dr = torch.empty_like(dl)
        C.call("dmb_b200_cat_volume_backward", C.ptr(dvol), C.ptr(dl), C.ptr(dr), B, Ch, H, W, C.int_array(idx),
               len(idx), C.stream(dvol.device))
        return dl, dr, None, None, None


class DifVolumeFn(torch.autograd.Function):
    """dif_fms forward + backward (cost_processors/utils/dif_fms.py:7-46): out = ref - shifted tgt, so the
    backward is the cat-volume gather with both halves reading the same channels and the target half negated."""

    @staticmethod
    def forward(ctx, left, right, max_disp, start_disp, dilation):
        ctx.args = (max_disp, start_disp, dilation)
        ctx.shape = tuple(left.shape)
        return F_.dif_volume(left, right, max_disp, start_disp, dilation)

    @staticmethod
    def backward(ctx, dvol):
        dvol = C.f32(dvol)
        B, Ch, H, W = ctx.shape
        idx = F_.disp_indices(*ctx.args)
        dl = torch.empty(ctx.shape, device=dvol.device, dtype=torch.float32)
        dr = torch.empty_like(dl)
        C.call("dmb_b200_dif_volume_backward", C.ptr(dvol), C.ptr(dl), C.ptr(dr), B, Ch, H, W, C.int_array(idx),
               len(idx), C.stream(dvol.device))
        return dl, dr, None, None, None


class UpsampleTrilinearFn(torch.autograd.Function):
    """[B,1,Dl,Hl,Wl] -> [B,D,H,W] (F.interpolate trilinear align_corners=True + squeeze, PSMNet.py:75-88)."""

    @staticmethod
    def forward(ctx, cost_low, size):
        ctx.low_shape = tuple(cost_low.shape)
        ctx.size = tuple(size)
        return F_.upsample_regress(cost_low, size, "trilinear")[0]

    @staticmethod
    def backward(ctx, dcost):
        dcost = C.f32(dcost)
        B = ctx.low_shape[0]
        Dl, Hl, Wl = ctx.low_shape[-3:]
        D, H, W = ctx.size
        dlow = torch.empty(ctx.low_shape, device=dcost.device, dtype=torch.float32)
        C.call("dmb_b200_upsample_trilinear_backward", C.ptr(dcost), C.ptr(dlow), B, Dl, Hl, Wl, D, H, W,
               C.stream(dcost.device))
        return dlow, None


class UpsampleDeconvFn(torch.autograd.Function):
    """AcfNet's learned upsampling ConvTranspose3d(1,1,8,4,2) (aggregators/AcfNet.py:55-57,81-83)."""

    @staticmethod
    def forward(ctx, cost_low, weight, size):
        low = C.f32(cost_low)
        ctx.size = tuple(size)
        ctx.save_for_backward(low, weight)
        return F_.upsample_regress(low, size, "deconv", weight.detach())[0]

    @staticmethod
    def backward(ctx, dcost):
        low, weight = ctx.saved_tensors
        dcost = C.f32(dcost)
        B = low.shape[0]
        Dl, Hl, Wl = low.shape[-3:]
        D, H, W = ctx.size
        dlow = dw = None
        if ctx.needs_input_grad[0]:
            # gradient of a transposed conv w.r.t. its input = the plain conv with the same weight
            w = F_.pack_conv_weight(weight.detach(), transposed=False)
            dlow = F_.conv3d_fused(dcost.view(B, 1, D, H, W), w, None, (8, 8, 8), 4, 2).view(low.shape)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(512, device=dcost.device, dtype=torch.float32)
            C.call("dmb_b200_upsample_deconv_wgrad", C.ptr(low), C.ptr(dcost), C.ptr(dw), B, Dl, Hl, Wl, D, H, W,
                   C.stream(dcost.device))
            dw = dw.view(weight.shape)
        return dlow, dw, None


class SoftArgminFn(torch.autograd.Function):
    """SoftArgmin / FasterSoftArgmin with a shared list of disparity samples."""

    @staticmethod
    def forward(ctx, cost, alpha, normalize, start_disp, disp_step, disp_values):
        cost = C.f32(cost)
        ctx.args = (float(alpha), bool(normalize), float(start_disp), float(disp_step))
        dv = C.f32(disp_values).reshape(-1) if disp_values is not None else None
        ctx.save_for_backward(cost, dv)
        return F_.soft_argmin(cost, alpha, normalize, start_disp, disp_step, dv)

    @staticmethod
    def backward(ctx, gdisp):
        cost, dv = ctx.saved_tensors
        alpha, normalize, start_disp, disp_step = ctx.args
        g = C.f32(gdisp)
        B, D, H, W = cost.shape
        dcost = torch.empty_like(cost)
        C.call("dmb_b200_soft_argmin_backward", C.ptr(cost), C.ptr(g), C.ptr(dcost), B, D, H, W, alpha,
               1 if normalize else 0, start_disp, disp_step, C.ptr(dv), C.stream(cost.device))
        return dcost, None, None, None, None, None


def wants_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


def forbid_grad(op, *tensors):
    """Ops whose backward is not built: fail loudly instead of returning a tensor without grad_fn (which would
    train the layers in front of it with a silent zero gradient)."""
    if wants_grad(*tensors):
        raise NotImplementedError(
            "%s has no backward on the CUDA path: call it under torch.no_grad() / on detached inputs, or use the "
            "differentiable variants (cat_fms / dif_fms 'default', SoftArgmin / FasterSoftArgmin)" % op)
