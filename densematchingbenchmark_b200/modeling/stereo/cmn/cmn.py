"""Confidence measurement network (reference: dmb/modeling/stereo/cmn/cmn.py:10-93), SURVEY.md section 8f row 1.

`ConfHead` = nn.Sequential(conv_bn_relu(in_planes, in_planes // 3, 3x3), Conv2d(in_planes // 3, 1, 1x1, bias=False))
on a full-resolution cost volume [B, in_planes = max_disp, H, W] -- at 544x960 a 58 GMAC convolution that reads
401 MB per head, the largest remaining consumer of the raw cost volumes in AcfNet-adaptive.  Same constructor,
sub-module names and state-dict keys (`conf_heads.<i>.conf_net.0.0.weight`, `.0.1.{weight,bias,running_*}`,
`.1.weight`) as the reference, so its checkpoints load.  The 3x3 convolution runs on the tcgen05 stride-1 kernel as a
one-plane ("flat") 3-D convolution:
  * eval: BatchNorm folded, `dmb_b200_conv2d_tc` (only the 9 in-plane taps are issued), then the 1x1 convolution as
    `dmb_b200_blocked_dot`; split IEEE-half arithmetic like the trunk (fp32-grade);
  * train (or an input that requires grad): the 2-D weight is embedded into a [Cout, Cin, 3, 3, 3] tensor by a
    differentiable zero pad and the unit goes through the library's conv+BatchNorm autograd Function (batch statistics,
    running-stat update of the BatchNorm2d module, tcgen05 forward / input-gradient / weight-gradient kernels); the
    1x1 convolution (64 -> 1, 0.03 % of the head's MACs) is a torch op there.
The NLL confidence loss and the evaluator wrapper (dmb/modeling/stereo/cmn/loss.py, losses/conf_nll_loss.py) are
element-wise consumers of [B,1,H,W] maps, outside the kernel path: restated here in plain torch so that `Cmn` is
self-contained."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import _cabi as C
from ....ops import functional as F_
from ....ops.autograd import ConvUnitFn, wants_grad


class ConfHead(nn.Module):

    def __init__(self, in_planes, batch_norm=True):
        super(ConfHead, self).__init__()
        self.in_planes = in_planes
        self.sec_in_planes = max(int(in_planes // 3), 1)
        unit = [nn.Conv2d(in_planes, self.sec_in_planes, 3, 1, 1, bias=False)]
        if batch_norm:
            unit.append(nn.BatchNorm2d(self.sec_in_planes))
        unit.append(nn.ReLU(inplace=True))
        self.conf_net = nn.Sequential(nn.Sequential(*unit), nn.Conv2d(self.sec_in_planes, 1, 1, 1, 0, bias=False))
        self.precision = "fp16x3"
        self._cache = None

    @property
    def _conv(self):
        return self.conf_net[0][0]

    @property
    def _bn(self):
        return self.conf_net[0][1] if isinstance(self.conf_net[0][1], nn.BatchNorm2d) else None

    def _tc_ok(self, cost):
        from ..cost_processors.aggregators import tc_engine as T
        return cost.is_cuda and self.in_planes % 32 == 0 and self.sec_in_planes % 32 == 0 and T.tc_available()

    # -- eval: folded BatchNorm, flat tcgen05 conv, 1x1 conv as a blocked dot product -----------------
    def _packed(self, device):
        from ..cost_processors.aggregators import tc_engine as T
        conv, bn = self._conv, self._bn
        tensors = [conv.weight, self.conf_net[1].weight] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var] if bn is not None else [])
        key = tuple((t.data_ptr(), t._version, str(t.device)) for t in tensors if t is not None) + (self.precision,)
        if self._cache is not None and self._cache[0] == key:
            return self._cache[1]
        split, fp16 = T.PRECISIONS[self.precision]
        with torch.no_grad():
            w = conv.weight.detach().float()
            b = None
            if bn is not None:
                scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
                w = w * scale.view(-1, 1, 1, 1)
                b = (bn.bias.detach().float() - bn.running_mean.float() * scale).contiguous()
            w3 = F.pad(w.unsqueeze(2), (0, 0, 0, 0, 1, 1))                      # [Cout, Cin, 3, 3, 3], kd = 0 / 2 zero
            packed = F_.pack_conv_weight(w3.contiguous(), False)                # [27, Cin, Cout]
            blob, Cin, Cout, wscale = T.pack_blob(packed, 3, split, fp16)
            w1 = self.conf_net[1].weight.detach().float().reshape(-1).contiguous()
        val = (blob, b, Cin, Cout, wscale, w1)
        self._cache = (key, val)
        return val

    def _forward_eval(self, cost):
        from ..cost_processors.aggregators import tc_engine as T
        cost = C.f32(cost)
        B, Cc, H, W = cost.shape
        split, fp16 = T.PRECISIONS[self.precision]
        blob, bias, Cin, Cout, wscale, w1 = self._packed(cost.device)
        xb = T.Blocked.from_ncdhw(cost.view(B, Cc, 1, H, W), split, fp16)
        y = T.Blocked.empty(B, Cout, (1, H, W), split, fp16, cost.device)
        C.call("dmb_b200_conv2d_tc", C.ptr(xb.hi), C.ptr(xb.lo), Cin, C.ptr(blob), float(wscale), C.ptr(bias), None, None,
               C.ptr(y.hi), C.ptr(y.lo), Cout, B, H, W, 1, 1 if fp16 else 0, C.stream(cost.device))
        T.check_finite(y, "ConfHead conv2d_tc")
        conf = torch.empty(B, 1, H, W, dtype=torch.float32, device=cost.device)
        C.call("dmb_b200_blocked_dot", C.ptr(y.hi), C.ptr(y.lo), C.ptr(w1), C.ptr(conf), B, Cout, H * W, 1 if fp16 else 0,
               C.stream(cost.device))
        return conf

    # -- train: the library's conv + BatchNorm autograd Function on the embedded 3-D weight ----------
    def _forward_train(self, cost):
        conv, bn = self._conv, self._bn
        B, Cc, H, W = cost.shape
        w3 = F.pad(conv.weight.unsqueeze(2), (0, 0, 0, 0, 1, 1))               # differentiable embedding
        batch_stats = bn is not None and bn.training
        if bn is not None and not batch_stats:                                   # frozen BatchNorm: fold differentiably
            scale = torch.rsqrt(bn.running_var.float() + bn.eps) * (bn.weight if bn.weight is not None else 1.0)
            w3 = w3 * scale.view(-1, 1, 1, 1, 1)
            bias = -bn.running_mean.float() * scale + (bn.bias if bn.bias is not None else 0.0)
        else:
            bias = None
        cfg = dict(transposed=False, ksize=(3, 3, 3), stride=1, pad=1, opad=0, relu=True,
                   bn=bn if batch_stats else None, sync_group=getattr(self, "sync_group", None))
        y = ConvUnitFn.apply(cost.reshape(B, Cc, 1, H, W), w3, bias, bn.weight if batch_stats else None,
                             bn.bias if batch_stats else None, None, cfg)
        return F.conv2d(y.reshape(B, -1, H, W), self.conf_net[1].weight)

    def forward(self, cost):
        if cost.dim() != 4 or cost.shape[1] != self.in_planes:
            raise ValueError("ConfHead expects a [B,%d,H,W] cost volume, got %s" % (self.in_planes, tuple(cost.shape)))
        if not self._tc_ok(cost):
            raise C.DmbB200Error("ConfHead runs on the tcgen05 kernels only: CUDA tensors on an sm_100 device, in_planes and "
                                 "in_planes // 3 multiples of 32 (got %d, %d); there is no fallback" % (self.in_planes, self.sec_in_planes))
        if self.training or wants_grad(cost):
            return self._forward_train(cost)
        return self._forward_eval(cost)


class ConfidenceNllLoss(object):
    """losses/conf_nll_loss.py:6-91: mean of -log sigmoid(confidence cost) over the pixels with a valid ground truth."""

    def __init__(self, max_disp, start_disp=0, weights=None, sparse=False):
        self.max_disp, self.start_disp, self.weights, self.sparse = max_disp, start_disp, weights, sparse
        self.scale_func = F.adaptive_max_pool2d if sparse else F.adaptive_avg_pool2d

    def loss_per_level(self, estConf, gtDisp):
        H, W = estConf.shape[-2:]
        gt, scale = gtDisp, 1.0
        if gtDisp.shape[-2] != H or gtDisp.shape[-1] != W:
            scale = gtDisp.shape[-1] / (W * 1.0)
            gt = self.scale_func(gtDisp / scale, (H, W))
        mask = ((gt > self.start_disp) & (gt < (self.max_disp / scale))).detach().type_as(gtDisp)
        valid = torch.clamp(mask.float().sum(), min=1.0)
        return (-1.0 * F.logsigmoid(estConf) * mask).sum() / valid

    def __call__(self, estConf, gtDisp):
        confs = list(estConf) if isinstance(estConf, (list, tuple)) else [estConf]
        if self.weights is None:
            self.weights = [1.0] * len(confs)
        return {"conf_loss_lvl%d" % i: self.weights[i] * self.loss_per_level(c, gtDisp) for i, c in enumerate(confs)}


def make_cmn_loss_evaluator(cfg):
    """cmn/loss.py:37-50: only the NLL loss exists; every level's loss is scaled by the configured weight."""
    losses = cfg.model.cmn.losses
    if "nll_loss" not in losses:
        return lambda confs, target: {}
    args = dict(losses.nll_loss)
    weight = args.pop("weight")
    args.update(sparse=cfg.data.sparse)
    nll = ConfidenceNllLoss(**args)
    return lambda confs, target: {k: v * weight for k, v in nll(confs, target).items()}


class Cmn(nn.Module):
    """Same constructor and forward contract as the reference (cmn.py:40-82): returns (cost variances, losses) in
    training and (cost variances, confidences) in eval."""

    def __init__(self, cfg, in_planes, num, alpha, beta):
        super(Cmn, self).__init__()
        self.cfg = cfg.copy()
        self.conf_heads = nn.ModuleList([ConfHead(in_planes, self.cfg.model.batch_norm) for _ in range(num)])
        self.loss_evaluator = make_cmn_loss_evaluator(cfg)
        self.alpha = alpha
        self.beta = beta

    def get_confidence(self, costs):
        assert len(self.conf_heads) == len(costs), \
            "NUM of confidence heads({}) must be equal to NUM of cost volumes({})".format(len(self.conf_heads), len(costs))
        conf_costs = [head(cost) for cost, head in zip(costs, self.conf_heads)]
        confs = [torch.sigmoid(c) for c in conf_costs]
        cost_vars = [self.alpha * (1 - conf) + self.beta for conf in confs]
        return confs, cost_vars, conf_costs

    def get_loss(self, confs, target=None):
        return self.loss_evaluator(confs, target)

    def forward(self, costs, target=None):
        confs, cost_vars, conf_costs = self.get_confidence(costs)
        if self.training:
            return cost_vars, self.get_loss(conf_costs, target)
        return cost_vars, confs


def build_cmn(cfg):
    c = cfg.model.cmn
    return Cmn(cfg, c.in_planes, c.num, c.alpha, c.beta)
