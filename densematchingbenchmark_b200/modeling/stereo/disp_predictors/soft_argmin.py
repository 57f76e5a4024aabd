"""SoftArgmin (reference: disp_predictors/soft_argmin.py:5-75)."""
import torch
import torch.nn as nn

from ....ops import functional as F_
from ....ops.autograd import SoftArgminFn, wants_grad
from ..cost_processors.aggregators.deferred import DeferredCost


class SoftArgmin(nn.Module):

    def __init__(self, max_disp=192, start_disp=0, dilation=1, alpha=1.0, normalize=True):
        super(SoftArgmin, self).__init__()
        self.max_disp = max_disp
        self.start_disp = start_disp
        self.dilation = dilation
        self.end_disp = start_disp + max_disp - 1
        self.disp_sample_number = (max_disp + dilation - 1) // dilation
        self.alpha = alpha
        self.normalize = normalize
        # [disp_sample_number] float32 samples, same linspace as the reference (:40-42)
        self.disp_sample = torch.linspace(self.start_disp, self.end_disp, self.disp_sample_number)
        self._dev_sample = None

    def _values(self, device):
        if self._dev_sample is None or self._dev_sample.device != device:
            self._dev_sample = self.disp_sample.to(device)
        return self._dev_sample

    def forward(self, cost_volume, disp_sample=None):
        if cost_volume.dim() != 4:
            raise ValueError('expected 4D input (got {}D input)'.format(cost_volume.dim()))
        D = cost_volume.shape[1]
        if disp_sample is None:
            assert D == self.disp_sample_number, 'The number of disparity samples should be consistent!'
            kw = dict(alpha=self.alpha, normalize=self.normalize, disp_values=self._values(cost_volume.device))
            if isinstance(cost_volume, DeferredCost):
                return cost_volume.regress(**kw)
            if wants_grad(cost_volume):
                return SoftArgminFn.apply(cost_volume, self.alpha, self.normalize, 0.0, 1.0, kw["disp_values"])
            return F_.soft_argmin(cost_volume, **kw)
        assert D == disp_sample.shape[1], 'The number of disparity samples should be consistent!'
        if wants_grad(cost_volume):
            raise NotImplementedError("SoftArgmin: the backward of the per-pixel disp_sample variant is not built "
                                      "(no shipped training config uses it)")
        return F_.soft_argmin(cost_volume, alpha=self.alpha, normalize=self.normalize, disp_sample=disp_sample)

    def __repr__(self):
        return ('{}\n    Max Disparity: {}\n    Start disparity: {}\n    Dilation rate: {}\n    Alpha: {}\n'
                '    Normalize: {}\n').format(self.__class__.__name__, self.max_disp, self.start_disp,
                                             self.dilation, self.alpha, self.normalize)

    @property
    def name(self):
        return 'SoftArgmin'
