"""StereoFocalLoss (reference: dmb/modeling/stereo/losses/stereo_focal_loss.py:9-140), same constructor, call
signature and result dict; the per-level loss and its gradients (w.r.t. the cost volume and, for AcfNet's
confidence-modulated variance, w.r.t. the variance map) are two CUDA kernels (csrc/focal_loss.cu) instead of
autograd over five materialised [B,D,H,W] volumes.  Only the Laplace ground-truth distribution the reference
hard-codes (`LaplaceDisp2Prob`, :87-90) is implemented."""
import torch
import torch.nn.functional as F

from .... import _cabi as C


class _FocalLossFn(torch.autograd.Function):

    @staticmethod
    def forward(ctx, cost, gt, var_map, var_scalar, disp_values, disp_sample, lower, upper, inner_end, coefficient):
        cost = C.f32(cost)
        B, D, H, W = cost.shape
        dev = cost.device
        sums = torch.zeros(2, dtype=torch.float64, device=dev)
        stats = torch.empty(B, 2, H, W, dtype=torch.float32, device=dev)
        C.call("dmb_b200_focal_loss_forward", C.ptr(cost), C.ptr(gt), C.ptr(var_map), float(var_scalar),
               C.ptr(disp_values), C.ptr(disp_sample), B, D, H, W, float(lower), float(upper), float(inner_end),
               float(coefficient), C.ptr(sums), C.ptr(stats), C.stream(dev))
        valid = sums[1].clamp_min(1.0)
        ctx.save_for_backward(cost, gt, var_map, disp_values, disp_sample, stats, valid)
        ctx.args = (float(var_scalar), float(lower), float(upper), float(inner_end), float(coefficient))
        return (sums[0] / valid).float()

    @staticmethod
    def backward(ctx, gout):
        cost, gt, var_map, disp_values, disp_sample, stats, valid = ctx.saved_tensors
        var_scalar, lower, upper, inner_end, coefficient = ctx.args
        B, D, H, W = cost.shape
        dev = cost.device
        gscale = (gout.double() / valid).float().reshape(1).contiguous()       # stays on the device: no host sync
        dcost = torch.empty_like(cost) if ctx.needs_input_grad[0] else None
        need_var = var_map is not None and ctx.needs_input_grad[2]
        dvar = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev) if need_var else None
        if dcost is None and dvar is None:
            return (None,) * 10
        scratch = dcost if dcost is not None else torch.empty_like(cost)
        C.call("dmb_b200_focal_loss_backward", C.ptr(cost), C.ptr(gt), C.ptr(var_map), var_scalar, C.ptr(disp_values),
               C.ptr(disp_sample), C.ptr(stats), C.ptr(gscale), B, D, H, W, lower, upper, inner_end, coefficient,
               C.ptr(scratch), C.ptr(dvar), C.stream(dev))
        return dcost, None, dvar, None, None, None, None, None, None, None


class StereoFocalLoss(object):
    """Inputs / outputs as the reference: estCost (Tensor or list) [B,D,H,W], gtDisp [B,1,H,W], variance (number,
    Tensor [B,1,H,W] or list), optional disp_sample [B,D,H,W]; returns {"stereo_focal_loss_lvl{i}": weight_i * loss_i}."""

    def __init__(self, max_disp, start_disp=0, dilation=1, weights=None, focal_coefficient=0.0, sparse=False):
        self.max_disp = max_disp
        self.start_disp = start_disp
        self.end_disp = self.max_disp + self.start_disp - 1
        self.dilation = dilation
        self.weights = weights
        self.focal_coefficient = focal_coefficient
        self.sparse = sparse
        self._disp_values = {}                     # disparity samples per (configuration, device), uploaded once
        # sparse ground truth (KITTI) -> max pooling, dense -> average pooling (stereo_focal_loss.py:55-61)
        self.scale_func = F.adaptive_max_pool2d if sparse else F.adaptive_avg_pool2d

    def loss_per_level(self, estCost, gtDisp, variance, dilation, disp_sample):
        if not estCost.is_cuda:
            raise C.DmbB200Error("StereoFocalLoss needs CUDA tensors; there is no CPU path")
        B, D, H, W = estCost.shape
        gt = gtDisp
        scale = 1.0
        if gtDisp.shape[-2] != H or gtDisp.shape[-1] != W:
            scale = gtDisp.shape[-1] / (W * 1.0)
            gt = self.scale_func(gtDisp / scale, (H, W))
        gt = C.f32(gt.detach())
        lower = self.start_disp
        max_disp = int(self.max_disp / scale)
        upper = lower + max_disp
        inner_end = self.start_disp + max_disp - 1
        disp_values = None
        if disp_sample is None:
            n = (max_disp + dilation - 1) // dilation
            if n != D:
                raise ValueError("cost volume has %d disparity samples, the loss configuration implies %d" % (D, n))
            # built on the host like the reference's (losses/utils/disp2prob.py), uploaded once per configuration and
            # device: a per-call host-to-device copy is a stream synchronisation and cannot be captured in a CUDA graph
            key = (self.start_disp, inner_end, n, estCost.device)
            disp_values = self._disp_values.get(key)
            if disp_values is None:
                disp_values = self._disp_values[key] = torch.linspace(self.start_disp, inner_end, n).to(estCost.device)
        else:
            disp_sample = C.f32(disp_sample.to(estCost.device).detach())
            assert (disp_sample.shape[0], disp_sample.shape[2], disp_sample.shape[3]) == (B, H, W), \
                'The (B, H, W) should be same between ground truth disparity map and disparity index!'
            if disp_sample.shape[1] != D:
                raise ValueError("disp_sample has %d samples, the cost volume %d" % (disp_sample.shape[1], D))
        var_map, var_scalar = None, 1.0
        if torch.is_tensor(variance):
            var_map = C.f32(variance.to(estCost.device).expand(B, 1, H, W))
        else:
            var_scalar = float(variance)
        return _FocalLossFn.apply(estCost, gt, var_map, var_scalar, disp_values, disp_sample, lower, upper, inner_end,
                                  self.focal_coefficient)

    @staticmethod
    def _per_level(value, levels):
        """A per-level list stays as it is; anything else is shared by all levels."""
        return list(value) if isinstance(value, (list, tuple)) else [value] * levels

    def __call__(self, estCost, gtDisp, variance, disp_sample=None):
        volumes = self._per_level(estCost, 1) if not isinstance(estCost, (list, tuple)) else list(estCost)
        levels = len(volumes)
        # like the reference, the broadcast settings are written back to the evaluator on first use (:104-113)
        self.weights = self._per_level(1.0 if self.weights is None else self.weights, levels)
        self.dilation = self._per_level(self.dilation, levels)
        result = {}
        for lvl, (volume, var, step, samples) in enumerate(zip(volumes, self._per_level(variance, levels), self.dilation,
                                                               self._per_level(disp_sample, levels))):
            result["stereo_focal_loss_lvl%d" % lvl] = self.weights[lvl] * self.loss_per_level(volume, gtDisp, var, step, samples)
        return result

    def __repr__(self):
        return ("{}(max_disp={}, start_disp={}, dilation={}, weights={}, focal_coefficient={}, sparse={})"
                .format(self.__class__.__name__, self.max_disp, self.start_disp, self.dilation, self.weights,
                        self.focal_coefficient, self.sparse))

    @property
    def name(self):
        return 'StereoFocalLoss'
