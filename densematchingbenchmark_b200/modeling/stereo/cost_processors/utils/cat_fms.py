"""CAT_FUNCS -- concatenation cost volumes (reference: cost_processors/utils/cat_fms.py:7-88)."""
from .....ops import functional as F_
from .....ops.autograd import CatVolumeFn, forbid_grad, wants_grad


def cat_fms(reference_fm, target_fm, max_disp=192, start_disp=0, dilation=1, disp_sample=None):
    """[B,C,H,W] x2 -> [B,2C,D,H,W] float32; `disp_sample` is ignored like in the reference
    (cat_fms.py:7).  One fused kernel instead of ~2*D slice-assign launches plus a CPU zeros +
    H2D copy (cat_fms.py:32-45)."""
    if wants_grad(reference_fm, target_fm):          # training: the same kernel + its backward
        return CatVolumeFn.apply(reference_fm, target_fm, max_disp, start_disp, dilation)
    return F_.cat_volume(reference_fm, target_fm, max_disp, start_disp, dilation)


def fast_cat_fms(reference_fm, target_fm, max_disp=192, start_disp=0, dilation=1, disp_sample=None):
    """grid_sample-warped variant with per-pixel `disp_sample` [B,D,H,W] (cat_fms.py:51-82),
    reproducing the reference's align_corners mismatch and `(target > 0)` masking."""
    forbid_grad("fast_cat_fms", reference_fm, target_fm, disp_sample)
    if disp_sample is None:
        disp_sample = _ramp_sample(reference_fm, max_disp, start_disp, dilation)
    return F_.warp_volume(reference_fm, target_fm, disp_sample, mode=0)


def _ramp_sample(fm, max_disp, start_disp, dilation):
    import torch
    B, _, H, W = fm.shape
    D = (max_disp + dilation - 1) // dilation
    s = torch.linspace(start_disp, start_disp + max_disp - 1, D)
    return s.view(1, D, 1, 1).expand(B, D, H, W).to(fm.device).float().contiguous()


CAT_FUNCS = dict(
    default=cat_fms,
    fast_mode=fast_cat_fms,
)
