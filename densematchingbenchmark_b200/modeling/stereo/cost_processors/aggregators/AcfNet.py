"""AcfAggregator (reference: cost_processors/aggregators/AcfNet.py:8-89): PSM trunk with conv
biases, learned ConvTranspose3d(1,1,8,4,2) upsampling (deconv1..3, :55-57,81-83)."""
import torch.nn as nn

from .PSMNet import PSMTrunk
from .deferred import DeferredCost
from .....ops import functional as F_
from .....ops.autograd import UpsampleDeconvFn


class AcfAggregator(PSMTrunk):

    def __init__(self, max_disp, in_planes=64, batch_norm=True):
        super(AcfAggregator, self).__init__(max_disp, in_planes, batch_norm, bias=True)
        self.deconv1 = nn.ConvTranspose3d(1, 1, 8, 4, 2, bias=False)
        self.deconv2 = nn.ConvTranspose3d(1, 1, 8, 4, 2, bias=False)
        self.deconv3 = nn.ConvTranspose3d(1, 1, 8, 4, 2, bias=False)
        self.defer_upsample = True

    def forward(self, raw_cost):
        D, H, W = raw_cost.dims if hasattr(raw_cost, "dims") else raw_cost.shape[2:]
        cost1, cost2, cost3 = self.trunk(raw_cost)
        size = (self.max_disp, H * 4, W * 4)
        if size[0] != 4 * D:
            raise ValueError("AcfAggregator: max_disp (%d) must be 4x the raw cost depth (%d), as the reference's "
                             "ConvTranspose3d(output_size=...) requires" % (self.max_disp, D))
        pairs = ((cost3, self.deconv3), (cost2, self.deconv2), (cost1, self.deconv1))
        diff = self.differentiable(raw_cost)
        if self.defer_upsample and not diff:
            return [DeferredCost(c[:, 0].contiguous(), size, "deconv", up.weight.detach()) for c, up in pairs]
        if diff:
            return [UpsampleDeconvFn.apply(c, up.weight, size) for c, up in pairs]
        return [F_.upsample_regress(c, size, "deconv", up.weight.detach())[0] for c, up in pairs]
