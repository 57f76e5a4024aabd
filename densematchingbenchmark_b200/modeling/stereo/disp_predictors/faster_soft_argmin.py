"""FasterSoftArgmin (reference: disp_predictors/faster_soft_argmin.py:6-75).

The reference regresses with a frozen Conv3d(1,1,(D,1,1)) whose weight is the linspace of
disparity samples; that weight is a registered parameter and therefore part of every checkpoint
(`disp_predictor.disp_regression.weight`).  The same parameter exists here and its VALUES are what
the fused kernel multiplies with, so a loaded checkpoint behaves identically."""
import torch
import torch.nn as nn

from ....ops import functional as F_
from ....ops.autograd import SoftArgminFn, wants_grad
from ..cost_processors.aggregators.deferred import DeferredCost


class FasterSoftArgmin(nn.Module):

    def __init__(self, max_disp, start_disp=0, dilation=1, alpha=1.0, normalize=True):
        super(FasterSoftArgmin, self).__init__()
        self.max_disp = max_disp
        self.start_disp = start_disp
        self.dilation = dilation
        self.end_disp = start_disp + max_disp - 1
        self.disp_sample_number = (max_disp + dilation - 1) // dilation
        self.alpha = alpha
        self.normalize = normalize
        disp_sample = torch.linspace(self.start_disp, self.end_disp, self.disp_sample_number)
        self.disp_regression = nn.Conv3d(1, 1, (self.disp_sample_number, 1, 1), 1, 0, bias=False)
        self.disp_regression.weight.data = disp_sample.view(1, 1, -1, 1, 1).contiguous()
        self.disp_regression.weight.requires_grad = False

    def forward(self, cost_volume, disp_sample=None):
        if cost_volume.dim() != 4:
            raise ValueError('expected 4D input (got {}D input)'.format(cost_volume.dim()))
        values = self.disp_regression.weight.detach().reshape(-1)
        if values.device != cost_volume.device:
            raise RuntimeError("disp_predictor lives on %s but the cost volume on %s"
                               % (values.device, cost_volume.device))
        if cost_volume.shape[1] != values.numel():
            # the reference's Conv3d would fail on the depth mismatch as well
            raise RuntimeError("cost volume has %d disparity samples, predictor expects %d"
                               % (cost_volume.shape[1], values.numel()))
        kw = dict(alpha=self.alpha, normalize=self.normalize, disp_values=values)
        if isinstance(cost_volume, DeferredCost):
            return cost_volume.regress(**kw)
        if wants_grad(cost_volume):
            return SoftArgminFn.apply(cost_volume, self.alpha, self.normalize, 0.0, 1.0, values)
        return F_.soft_argmin(cost_volume, **kw)

    def __repr__(self):
        return ('{}\n    Max Disparity: {}\n    Start disparity: {}\n    Dilation rate: {}\n    Alpha: {}\n'
                '    Normalize: {}\n').format(self.__class__.__name__, self.max_disp, self.start_disp,
                                             self.dilation, self.alpha, self.normalize)

    @property
    def name(self):
        return 'FasterSoftArgmin'
