"""GCAggregator (reference: cost_processors/aggregators/GCNet.py:7-120): 14 conv + 5 deconv
encoder/decoder on a half-resolution volume; every skip add is fused into the producing deconv's
input... the reference adds BEFORE the next layer, so the add rides as that layer's producer
epilogue here (layer33..36 outputs get their skip as a fused residual after the ReLU is not
possible -- see forward)."""
import torch
import torch.nn as nn

from ...layers.basic_layers import conv3d_bn_relu, deconv3d_bn_relu, fused_plain_conv3d
from .....ops.autograd import wants_grad


class GCAggregator(nn.Module):

    def __init__(self, max_disp, in_planes=64, batch_norm=True):
        super(GCAggregator, self).__init__()
        self.max_disp = max_disp
        self.in_planes = in_planes
        self.batch_norm = batch_norm
        self.F = F = in_planes // 2
        mk = self._make_layer
        self.layer19 = mk(in_planes, F)
        self.layer20 = mk(F, F)
        self.layer21 = mk(in_planes + F, F * 2, 2)
        self.layer22 = mk(F * 2, F * 2)
        self.layer23 = mk(F * 2, F * 2)
        self.layer24 = mk(F * 4, F * 2, 2)
        self.layer25 = mk(F * 2, F * 2)
        self.layer26 = mk(F * 2, F * 2)
        self.layer27 = mk(F * 4, F * 2, 2)
        self.layer28 = mk(F * 2, F * 2)
        self.layer29 = mk(F * 2, F * 2)
        self.layer30 = mk(F * 4, F * 4, 2)
        self.layer31 = mk(F * 4, F * 4)
        self.layer32 = mk(F * 4, F * 4)
        self.layer33 = self._make_tlayer(F * 4, F * 2)
        self.layer34 = self._make_tlayer(F * 2, F * 2)
        self.layer35 = self._make_tlayer(F * 2, F * 2)
        self.layer36 = self._make_tlayer(F * 2, F)
        self.layer37 = nn.ConvTranspose3d(F, 1, kernel_size=3, stride=2, padding=1, output_padding=1)
        # 'direct': fp32 SIMT kernels; 'tc': tcgen05 (layers 19..36); 'auto': tc when the shape allows
        self.engine = "auto"
        self.precision = "fp16x3"

    def _make_layer(self, cin, cout, stride=1):
        return conv3d_bn_relu(self.batch_norm, cin, cout, kernel_size=3, stride=stride, padding=1, dilation=1,
                              bias=False)

    def _make_tlayer(self, cin, cout):
        return deconv3d_bn_relu(self.batch_norm, cin, cout, kernel_size=3, stride=2, padding=1, output_padding=1,
                                bias=False)

    def forward(self, raw_cost):
        if self.engine != "direct" and not self.training and not wants_grad(raw_cost):
            from . import tc_engine
            if tc_engine.gc_shape_ok(self, raw_cost):
                v = tc_engine.run_gc_tc(self, raw_cost, self.precision)
                return [fused_plain_conv3d(self.layer37, v).squeeze(dim=1)]
            if self.engine == "tc":
                raise RuntimeError("engine='tc' requested but the tcgen05 path does not support this shape/build")
        v18 = raw_cost
        v19 = self.layer19(v18)
        v20 = self.layer20(v19)
        v21 = self.layer21(torch.cat([v18, v20], dim=1))
        v22 = self.layer22(v21)
        v23 = self.layer23(v22)
        v24 = self.layer24(torch.cat([v21, v23], dim=1))
        v25 = self.layer25(v24)
        v26 = self.layer26(v25)
        v27 = self.layer27(torch.cat([v24, v26], dim=1))
        v28 = self.layer28(v27)
        v29 = self.layer29(v28)
        v30 = self.layer30(torch.cat([v27, v29], dim=1))
        v31 = self.layer31(v30)
        v32 = self.layer32(v31)
        # skip adds happen AFTER the deconv's ReLU (GCNet.py:108-116): plain tensor adds
        v33 = self.layer33(v32)
        v34 = self.layer34(v33 + v29)
        v35 = self.layer35(v34 + v26)
        v36 = self.layer36(v35 + v23)
        v37 = fused_plain_conv3d(self.layer37, v36 + v20)
        return [v37.squeeze(dim=1)]
