"""PSMAggregator (reference: cost_processors/aggregators/PSMNet.py:9-95), same constructor,
parameters and state-dict keys; forward = 25 fused conv launches + the upsampling kernels."""
import torch.nn as nn

from ...layers.basic_layers import conv3d_bn, conv3d_bn_relu, fused_plain_conv3d
from ..utils.hourglass import Hourglass
from .deferred import DeferredCost
from .....ops import functional as F_
from .....ops.autograd import UpsampleTrilinearFn, wants_grad


class PSMTrunk(nn.Module):
    """dres0..dres4 + classif1..3 shared by PSMAggregator and AcfAggregator.  `bias` mirrors the
    factory defaults each reference class ends up with (PSMNet: bias=False everywhere,
    aggregators/PSMNet.py:30-53; AcfNet: bias=True on the trunk convs, aggregators/AcfNet.py:30-53)."""

    def __init__(self, max_disp, in_planes=64, batch_norm=True, bias=False):
        super(PSMTrunk, self).__init__()
        self.max_disp = max_disp
        self.in_planes = in_planes
        self.batch_norm = batch_norm
        bn = batch_norm
        self.dres0 = nn.Sequential(
            conv3d_bn_relu(bn, self.in_planes, 32, 3, 1, 1, bias=bias),
            conv3d_bn_relu(bn, 32, 32, 3, 1, 1, bias=bias),
        )
        self.dres1 = nn.Sequential(
            conv3d_bn_relu(bn, 32, 32, 3, 1, 1, bias=bias),
            conv3d_bn(bn, 32, 32, 3, 1, 1, bias=bias),
        )
        self.dres2 = Hourglass(in_planes=32, batch_norm=bn)
        self.dres3 = Hourglass(in_planes=32, batch_norm=bn)
        self.dres4 = Hourglass(in_planes=32, batch_norm=bn)
        for name in ("classif1", "classif2", "classif3"):
            setattr(self, name, nn.Sequential(
                conv3d_bn_relu(bn, 32, 32, 3, 1, 1, bias=bias),
                nn.Conv3d(32, 1, kernel_size=3, stride=1, padding=1, bias=False),
            ))
        # 'direct': fp32 SIMT kernels; 'tc': tcgen05 trunk (csrc/conv3d_tc.cu); 'auto': tc when available
        self.engine = "auto"
        # tc engine arithmetic: 'fp16x3' (split fp16, fp32-equivalent), 'bf16x3', 'fp16', 'bf16'
        self.precision = "fp16x3"

    def trunk(self, raw_cost):
        """raw [B,C,D,H,W] -> (cost1, cost2, cost3), each [B,1,D,H,W] (PSMNet.py:58-72)."""
        if self._use_tc(raw_cost):
            from .tc_engine import run_trunk_tc
            return run_trunk_tc(self, raw_cost)
        cost0 = self.dres0[1](self.dres0[0](raw_cost))
        cost0 = self.dres1[1](self.dres1[0](cost0), residual=cost0)
        out1, pre1, post1 = self.dres2(cost0, None, None, out_residual=cost0)
        out2, pre2, post2 = self.dres3(out1, pre1, post1, out_residual=cost0)
        out3, pre3, post3 = self.dres4(out2, pre2, post2, out_residual=cost0)
        cost1 = fused_plain_conv3d(self.classif1[1], self.classif1[0](out1))
        cost2 = fused_plain_conv3d(self.classif2[1], self.classif2[0](out2), residual=cost1)
        cost3 = fused_plain_conv3d(self.classif3[1], self.classif3[0](out3), residual=cost2)
        return cost1, cost2, cost3

    def blocked_cat_volume(self, ref_fms, tgt_fms, max_disp=192, start_disp=0, dilation=1):
        """Fast path used by CatCostProcessor: when the trunk will run on tcgen05, build the
        concatenation volume directly in the trunk's blocked 16-bit layout.  Returns None when
        the trunk is not going to use the tensor-core engine for this shape."""
        if self.engine == "direct" or self.training or not ref_fms.is_cuda or wants_grad(ref_fms, tgt_fms):
            return None                                   # (the blocked volume and the tcgen05 trunk have no backward)
        from . import tc_engine
        D = (max_disp + dilation - 1) // dilation
        dhw = (D, ref_fms.shape[2], ref_fms.shape[3])
        if ref_fms.shape[1] % 8 or not tc_engine.tc_shape_ok(self, 2 * ref_fms.shape[1], dhw):
            return None
        return tc_engine.cat_volume_blocked(ref_fms, tgt_fms, max_disp, start_disp, dilation, self.precision)

    def differentiable(self, raw_cost):
        """True when the forward must build an autograd graph: training mode, or an eval-mode call whose input
        carries a gradient (inference without torch.no_grad() on top of a trainable backbone)."""
        return self.training or (not hasattr(raw_cost, "dims") and wants_grad(raw_cost))

    def _use_tc(self, raw_cost):
        if self.engine == "direct" or self.differentiable(raw_cost):      # training runs the autograd path
            return False
        from .tc_engine import tc_supported
        ok = tc_supported(self, raw_cost)
        if self.engine == "tc" and not ok:
            raise RuntimeError("engine='tc' requested but the tcgen05 trunk does not support this shape/build")
        return ok


class PSMAggregator(PSMTrunk):
    """Inputs: raw_cost [B,in_planes,D/4,H/4,W/4].  Outputs: [cost3, cost2, cost1], each
    [B,max_disp,H,W] (a DeferredCost in eval mode when `defer_upsample` is set)."""

    def __init__(self, max_disp, in_planes=64, batch_norm=True):
        super(PSMAggregator, self).__init__(max_disp, in_planes, batch_norm, bias=False)
        self.defer_upsample = True

    def forward(self, raw_cost):
        D, H, W = raw_cost.dims if hasattr(raw_cost, "dims") else raw_cost.shape[2:]
        cost1, cost2, cost3 = self.trunk(raw_cost)
        size = (self.max_disp, H * 4, W * 4)
        diff = self.differentiable(raw_cost)
        if self.defer_upsample and not diff:
            return [DeferredCost(c[:, 0].contiguous(), size, "trilinear") for c in (cost3, cost2, cost1)]
        if diff:
            return [UpsampleTrilinearFn.apply(c, size) for c in (cost3, cost2, cost1)]
        return [F_.upsample_regress(c, size, "trilinear")[0] for c in (cost3, cost2, cost1)]
