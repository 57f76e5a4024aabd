"""LocalSoftArgmin (reference: disp_predictors/local_soft_argmin.py:5-105)."""
import torch.nn as nn

from ....ops import functional as F_
from ....ops.autograd import forbid_grad


class LocalSoftArgmin(nn.Module):

    def __init__(self, max_disp, radius, start_disp=0, dilation=1, radius_dilation=1, alpha=1.0, normalize=True):
        super(LocalSoftArgmin, self).__init__()
        self.max_disp = max_disp
        self.radius = radius
        self.start_disp = start_disp
        self.dilation = dilation
        self.radius_dilation = radius_dilation
        self.end_disp = start_disp + max_disp - 1
        self.disp_sample_number = (max_disp + dilation - 1) // dilation
        self.alpha = alpha
        self.normalize = normalize   # unused: the window is always soft-maxed (reference docstring :9)

    def forward(self, cost_volume, disp_sample=None):
        D = cost_volume.size()[1]
        assert D == self.disp_sample_number, 'Number of disparity sample should be same' \
                                             'with predicted disparity number in cost volume!'
        forbid_grad("LocalSoftArgmin", cost_volume)
        return F_.local_soft_argmin(cost_volume, self.radius, self.radius_dilation, self.alpha,
                                    self.start_disp, self.dilation)

    def __repr__(self):
        return ('{}\n    Max Disparity: {}\n    Local disparity sample radius: {}\n    Start disparity: {}\n'
                '    Dilation rate: {}\n    Local disparity sample dilation rate: {}\n    Alpha: {}\n'
                '    Normalize: {}\n').format(self.__class__.__name__, self.max_disp, self.radius, self.start_disp,
                                             self.dilation, self.radius_dilation, self.alpha, self.normalize)

    @property
    def name(self):
        return 'LocalSoftArgmin'
