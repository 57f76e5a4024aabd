"""Functional wrappers: torch tensors in, torch tensors out, arithmetic in the CUDA library.

Each function validates shapes like its reference counterpart would fail, allocates the output
on the input's device and launches on the current stream.  None of them has a CPU path."""
import torch

from .. import _cabi as C


def disp_indices(max_disp, start_disp=0, dilation=1):
    """Integer disparities exactly as the reference enumerates them: `int()` of a float32
    `torch.linspace(start, start+max-1, n)` (cat_fms.py:27-35).  Host-side, n <= a few hundred."""
    n = (max_disp + dilation - 1) // dilation
    return [int(v) for v in torch.linspace(start_disp, start_disp + max_disp - 1, n)]


def _check_pair(reference_fm, target_fm):
    if reference_fm.dim() != 4 or target_fm.shape != reference_fm.shape:
        raise ValueError("expected two [B,C,H,W] feature maps of equal shape, got %s and %s"
                         % (tuple(reference_fm.shape), tuple(target_fm.shape)))


def cat_volume(reference_fm, target_fm, max_disp=192, start_disp=0, dilation=1):
    _check_pair(reference_fm, target_fm)
    l, r = C.f32(reference_fm), C.f32(target_fm)
    B, Ch, H, W = l.shape
    idx = disp_indices(max_disp, start_disp, dilation)
    out = torch.empty(B, 2 * Ch, len(idx), H, W, device=l.device, dtype=torch.float32)
    C.call("dmb_b200_cat_volume", C.ptr(l), C.ptr(r), C.ptr(out), B, Ch, H, W, C.int_array(idx), len(idx),
           C.stream(l.device))
    return out


def dif_volume(reference_fm, target_fm, max_disp=192, start_disp=0, dilation=1):
    _check_pair(reference_fm, target_fm)
    l, r = C.f32(reference_fm), C.f32(target_fm)
    B, Ch, H, W = l.shape
    idx = disp_indices(max_disp, start_disp, dilation)
    out = torch.empty(B, Ch, len(idx), H, W, device=l.device, dtype=torch.float32)
    C.call("dmb_b200_dif_volume", C.ptr(l), C.ptr(r), C.ptr(out), B, Ch, H, W, C.int_array(idx), len(idx),
           C.stream(l.device))
    return out


def gwc_volume(reference_fm, target_fm, num_groups, max_disp=192, start_disp=0, dilation=1):
    _check_pair(reference_fm, target_fm)
    l, r = C.f32(reference_fm), C.f32(target_fm)
    B, Ch, H, W = l.shape
    if Ch % num_groups:
        raise ValueError("channels %d not divisible by num_groups %d" % (Ch, num_groups))
    idx = disp_indices(max_disp, start_disp, dilation)
    out = torch.empty(B, num_groups, len(idx), H, W, device=l.device, dtype=torch.float32)
    C.call("dmb_b200_gwc_volume", C.ptr(l), C.ptr(r), C.ptr(out), B, Ch, H, W, num_groups, C.int_array(idx),
           len(idx), C.stream(l.device))
    return out


def warp_volume(reference_fm, target_fm, disp_sample, mode, p=1.0):
    """mode 0: fast_cat_fms, 1: fast_dif_fms, 2: fast_dif_fms(normalize=True, p)."""
    _check_pair(reference_fm, target_fm)
    l, r, ds = C.f32(reference_fm), C.f32(target_fm), C.f32(disp_sample)
    B, Ch, H, W = l.shape
    if ds.dim() != 4 or ds.shape[0] != B or tuple(ds.shape[2:]) != (H, W):
        raise ValueError("disp_sample must be [B,D,H,W]")
    D = ds.shape[1]
    shape = {0: (B, 2 * Ch, D, H, W), 1: (B, Ch, D, H, W), 2: (B, D, H, W)}[mode]
    out = torch.empty(shape, device=l.device, dtype=torch.float32)
    C.call("dmb_b200_warp_volume", C.ptr(l), C.ptr(r), C.ptr(ds), C.ptr(out), B, Ch, H, W, D, mode, float(p),
           C.stream(l.device))
    return out


def pack_conv_weight(weight, transposed=False):
    """[Cout,Cin,kd,kh,kw] (Conv3d) or [Cin,Cout,kd,kh,kw] (ConvTranspose3d) -> [K3,Cin,Cout]."""
    if transposed:
        w = weight.permute(2, 3, 4, 0, 1)
    else:
        w = weight.permute(2, 3, 4, 1, 0)
    k3 = weight.shape[2] * weight.shape[3] * weight.shape[4]
    return w.reshape(k3, w.shape[3], w.shape[4]).contiguous().float()


def conv3d_fused(x, w_packed, bias, ksize, stride=1, pad=1, transposed=False, output_padding=0,
                 residual=None, relu=False, out_dims=None):
    """y = act(conv(x) + bias + residual), fp32 NCDHW, generic direct kernel.  `out_dims` (transposed
    only) names the output extent explicitly -- per-dimension output padding, as
    ConvTranspose3d(output_size=...) / a conv input-gradient needs; the library validates it."""
    x = C.f32(x)
    B, Cin, Di, Hi, Wi = x.shape
    K3, Cin_w, Cout = w_packed.shape
    if Cin_w != Cin or K3 != ksize[0] * ksize[1] * ksize[2]:
        raise ValueError("packed weight %s does not match input channels %d / kernel %s"
                         % (tuple(w_packed.shape), Cin, ksize))
    if transposed:
        dims_out = [(n - 1) * stride - 2 * pad + k + output_padding for n, k in zip((Di, Hi, Wi), ksize)]
        if out_dims is not None:
            dims_out = [int(v) for v in out_dims]
    else:
        dims_out = [(n + 2 * pad - k) // stride + 1 for n, k in zip((Di, Hi, Wi), ksize)]
    y = torch.empty(B, Cout, *dims_out, device=x.device, dtype=torch.float32)
    if residual is not None:
        residual = C.f32(residual)
        if residual.shape != y.shape:
            raise ValueError("residual shape %s != output shape %s" % (tuple(residual.shape), tuple(y.shape)))
    C.call("dmb_b200_conv3d_direct", C.ptr(x), C.ptr(w_packed), C.ptr(bias), C.ptr(residual), C.ptr(y),
           B, Cin, Cout, C.int_array([Di, Hi, Wi]), C.int_array(dims_out), C.int_array(list(ksize)),
           stride, pad, 1 if transposed else 0, 1 if relu else 0, C.stream(x.device))
    return y


def upsample_regress(cost_low, out_dhw, mode="trilinear", up_weight=None, want_cost=True, want_disp=False,
                     alpha=1.0, normalize=True, start_disp=0.0, disp_step=1.0, disp_values=None):
    """cost_low [B,1,Dl,Hl,Wl] or [B,Dl,Hl,Wl] -> (cost [B,D,H,W] or None, disp [B,1,H,W] or None)."""
    if cost_low.dim() == 5:
        if cost_low.shape[1] != 1:
            raise ValueError("expected a single-channel low-resolution cost")
        cost_low = cost_low[:, 0]
    cl = C.f32(cost_low)
    B, Dl, Hl, Wl = cl.shape
    D, H, W = out_dhw
    cost = torch.empty(B, D, H, W, device=cl.device, dtype=torch.float32) if want_cost else None
    disp = torch.empty(B, 1, H, W, device=cl.device, dtype=torch.float32) if want_disp else None
    m = {"trilinear": 0, "deconv": 1}[mode]
    upw = C.f32(up_weight).reshape(-1) if up_weight is not None else None
    dv = C.f32(disp_values).reshape(-1) if disp_values is not None else None
    if dv is not None and dv.numel() != D:
        raise AssertionError("The number of disparity samples should be consistent!")
    C.call("dmb_b200_upsample_regress", C.ptr(cl), C.ptr(upw), C.ptr(cost), C.ptr(disp), B, Dl, Hl, Wl, D, H, W, m,
           float(alpha), 1 if normalize else 0, float(start_disp), float(disp_step), C.ptr(dv), C.stream(cl.device))
    return cost, disp


def soft_argmin(cost_volume, alpha=1.0, normalize=True, start_disp=0.0, disp_step=1.0, disp_values=None,
                disp_sample=None):
    if cost_volume.dim() != 4:
        raise ValueError('expected 4D input (got {}D input)'.format(cost_volume.dim()))
    c = C.f32(cost_volume)
    B, D, H, W = c.shape
    dv = C.f32(disp_values).reshape(-1) if disp_values is not None else None
    ds = C.f32(disp_sample) if disp_sample is not None else None
    if dv is not None:
        assert dv.numel() == D, 'The number of disparity samples should be consistent!'
    if ds is not None:
        assert ds.shape[1] == D, 'The number of disparity samples should be consistent!'
        if ds.shape != c.shape:
            ds = ds.expand_as(c).contiguous()
    out = torch.empty(B, 1, H, W, device=c.device, dtype=torch.float32)
    C.call("dmb_b200_soft_argmin", C.ptr(c), C.ptr(out), B, D, H, W, float(alpha), 1 if normalize else 0,
           float(start_disp), float(disp_step), C.ptr(dv), C.ptr(ds), C.stream(c.device))
    return out


def local_soft_argmin(cost_volume, radius, radius_dilation=1, alpha=1.0, start_disp=0.0, dilation=1.0):
    if cost_volume.dim() != 4:
        raise ValueError('expected 4D input (got {}D input)'.format(cost_volume.dim()))
    c = C.f32(cost_volume)
    B, D, H, W = c.shape
    out = torch.empty(B, 1, H, W, device=c.device, dtype=torch.float32)
    C.call("dmb_b200_local_soft_argmin", C.ptr(c), C.ptr(out), B, D, H, W, int(radius), int(radius_dilation),
           float(alpha), float(start_disp), float(dilation), C.stream(c.device))
    return out


def sga(x, guidance):
    """x [B,C,D,H,W]; guidance [B,4*5*C,H,W] or [B,4,5,C,H,W] -> [B,C,D,H,W]."""
    x = C.f32(x)
    B, Ch, D, H, W = x.shape
    g = C.f32(guidance).reshape(B, 4, 5, Ch, H, W)
    out = torch.empty_like(x)
    C.call("dmb_b200_sga", C.ptr(x), C.ptr(g), C.ptr(out), B, Ch, D, H, W, C.stream(x.device))
    return out


def lga(x, guidance, radius=2):
    """x [B,D,H,W]; guidance [B,3*K*K,H,W] -> [B,D,H,W]."""
    x = C.f32(x)
    B, D, H, W = x.shape
    K = 2 * radius + 1
    g = C.f32(guidance).reshape(B, 3 * K * K, H, W)
    out = torch.empty_like(x)
    C.call("dmb_b200_lga", C.ptr(x), C.ptr(g), C.ptr(out), B, D, H, W, int(radius), C.stream(x.device))
    return out
