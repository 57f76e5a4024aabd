"""DIF_FUNCS -- difference cost volumes (reference: cost_processors/utils/dif_fms.py:7-92)."""
from .....ops import functional as F_
from .....ops.autograd import DifVolumeFn, forbid_grad, wants_grad
from .cat_fms import _ramp_sample


def dif_fms(reference_fm, target_fm, max_disp=192, start_disp=0, dilation=1, disp_sample=None,
            normalize=False, p=1.0):
    """[B,C,H,W] x2 -> [B,C,D,H,W]; `normalize`/`p` are ignored by the reference's default
    variant too (dif_fms.py:7-46)."""
    if wants_grad(reference_fm, target_fm):          # training (StereoNet configs): the same kernel + its backward
        return DifVolumeFn.apply(reference_fm, target_fm, max_disp, start_disp, dilation)
    return F_.dif_volume(reference_fm, target_fm, max_disp, start_disp, dilation)


def fast_dif_fms(reference_fm, target_fm, max_disp=192, start_disp=0, dilation=1, disp_sample=None,
                 normalize=False, p=1.0):
    forbid_grad("fast_dif_fms", reference_fm, target_fm, disp_sample)
    if disp_sample is None:
        disp_sample = _ramp_sample(reference_fm, max_disp, start_disp, dilation)
    return F_.warp_volume(reference_fm, target_fm, disp_sample, mode=2 if normalize else 1, p=p)


DIF_FUNCS = dict(
    default=dif_fms,
    fast_mode=fast_dif_fms,
)
