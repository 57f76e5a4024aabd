"""StereoNetAggregator (reference: cost_processors/aggregators/StereoNet.py:9-55)."""
import torch
import torch.nn as nn

from ...layers.basic_layers import conv3d_bn_relu, fused_plain_conv3d
from .....ops.autograd import wants_grad


class StereoNetAggregator(nn.Module):

    def __init__(self, max_disp, in_planes=32, batch_norm=True, num=4):
        super(StereoNetAggregator, self).__init__()
        self.max_disp = max_disp
        self.in_planes = in_planes
        self.batch_norm = batch_norm
        self.num = num
        self.classify = nn.ModuleList([
            conv3d_bn_relu(batch_norm, in_planes, 32, kernel_size=3, stride=1, padding=1, dilation=1, bias=True)
            for _ in range(num)
        ])
        self.lastconv = nn.Conv3d(32, 1, kernel_size=3, stride=1, padding=1, bias=True)
        self.engine = "auto"               # 'direct' | 'tc' | 'auto' (tcgen05 when the shape allows)
        self.precision = "fp16x3"

    def forward(self, raw_cost):
        if self.engine != "direct" and not self.training and not wants_grad(raw_cost):
            from . import tc_engine
            if raw_cost.is_cuda and raw_cost.dim() == 5 and raw_cost.shape[1] == self.in_planes == 32 \
                    and tc_engine.tc_available():
                return [torch.squeeze(tc_engine.run_stereonet_tc(self, raw_cost, self.precision), 1)]
            if self.engine == "tc":
                raise RuntimeError("engine='tc' requested but the tcgen05 path does not support this shape/build")
        for layer in self.classify:
            raw_cost = layer(raw_cost)
        cost = fused_plain_conv3d(self.lastconv, raw_cost)
        return [torch.squeeze(cost, 1)]
