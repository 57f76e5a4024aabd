"""COR_FUNCS -- 1-D correlation cost volume (reference: cost_processors/utils/correlation1d_cost.py:7-31).

The reference delegates to `spatial_correlation_sampler.SpatialCorrelationSampler` (an un-vendored third-party CUDA
extension, INSTALL.md:60, no version pinned) with patch_size = (1, 2 * max_disp - 1), keeps the first max_disp
channels and applies leaky ReLU(0.1).  Restated from that package's published definition (sum over channels of
input1(x) * input2(x + displacement), zero outside the image, no normalisation): channel j holds the correlation at
disparity max_disp - 1 - j.  **Parity unpinned** (the dependency cannot be imported anywhere in this project): the
contract is kernel == oracle/dmb_oracle.py:correlation1d_cost.  `start_disp`, `dilation` and `disp_sample` are accepted
and ignored exactly like the reference does."""
import torch

from ..... import _cabi as C
from .....ops.autograd import forbid_grad


def correlation1d_cost(reference_fm, target_fm, max_disp=192, start_disp=0, dilation=1, disp_sample=None,
                       kernel_size=1, stride=1, padding=0, dilation_patch=1):
    if (kernel_size, stride, padding, dilation_patch) != (1, 1, 0, 1):
        raise NotImplementedError("correlation1d_cost: only the reference's defaults kernel_size=1, stride=1, padding=0, "
                                  "dilation_patch=1 are built")
    forbid_grad("correlation1d_cost", reference_fm, target_fm)
    if reference_fm.shape != target_fm.shape or reference_fm.dim() != 4:
        raise ValueError("reference_fm / target_fm must be [B,C,H,W] tensors of equal shape")
    l, r = C.f32(reference_fm), C.f32(target_fm)
    B, Ch, H, W = l.shape
    out = torch.empty(B, max_disp, H, W, device=l.device, dtype=torch.float32)
    C.call("dmb_b200_corr1d_volume", C.ptr(l), C.ptr(r), C.ptr(out), B, Ch, H, W, int(max_disp), 0.1, C.stream(l.device))
    return out


COR_FUNCS = dict(
    default=correlation1d_cost,
)
