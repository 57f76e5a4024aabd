"""GWC_FUNCS -- group-wise correlation cost volume (GwcNet).  NEW table: the reference snapshot
names GwcNet (README.md:16) but ships no builder; signature follows its sibling builders."""
from .....ops import functional as F_
from .....ops.autograd import forbid_grad


def gwc_fms(reference_fm, target_fm, max_disp=192, start_disp=0, dilation=1, disp_sample=None, num_groups=40):
    """[B,C,H,W] x2 -> [B,num_groups,D,H,W]: mean over each group's channels of L(x)*R(x-d)."""
    forbid_grad("gwc_fms", reference_fm, target_fm)
    return F_.gwc_volume(reference_fm, target_fm, num_groups, max_disp, start_disp, dilation)


GWC_FUNCS = dict(
    default=gwc_fms,
)
