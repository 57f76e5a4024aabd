"""dmb.ops.spn mirror: GateRecurrent2dnoind (module) and its autograd Function
(dmb/ops/spn/modules/gaterecurrent2dnoind.py:4-12, functions/gaterecurrent2dnoind.py:8-39),
running the single-launch scan kernels of csrc/scans.cu through the C ABI."""
import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _cabi as C


class GateRecurrent2dnoindFunction(Function):

    @staticmethod
    def forward(ctx, X, G1, G2, G3, horizontal, reverse):
        if not X.is_cuda:
            # the reference prints "cpu version is not ready" and returns 0
            # (functions/gaterecurrent2dnoind.py:15-17); a silent wrong value helps nobody
            raise RuntimeError("GateRecurrent2dnoind: CUDA tensors required (no CPU implementation)")
        X, G1, G2, G3 = [C.f32(t) for t in (X, G1, G2, G3)]
        if not (X.shape == G1.shape == G2.shape == G3.shape) or X.dim() != 4:
            raise ValueError("X, G1, G2, G3 must be [N,C,H,W] tensors of equal shape")
        N, Ch, H, W = X.shape
        out = torch.empty_like(X)
        C.call("dmb_b200_spn_forward", C.ptr(X), C.ptr(G1), C.ptr(G2), C.ptr(G3), C.ptr(out), N, Ch, H, W,
               1 if horizontal else 0, 1 if reverse else 0, C.stream(X.device))
        ctx.save_for_backward(X, G1, G2, G3, out)
        ctx.horizontal = horizontal
        ctx.reverse = reverse
        return out

    @staticmethod
    def backward(ctx, grad_output):
        X, G1, G2, G3, out = ctx.saved_tensors
        go = C.f32(grad_output)
        N, Ch, H, W = X.shape
        gX, g1, g2, g3 = [torch.empty_like(X) for _ in range(4)]
        C.call("dmb_b200_spn_backward", C.ptr(X), C.ptr(G1), C.ptr(G2), C.ptr(G3), C.ptr(out), C.ptr(go),
               C.ptr(gX), C.ptr(g1), C.ptr(g2), C.ptr(g3), N, Ch, H, W,
               1 if ctx.horizontal else 0, 1 if ctx.reverse else 0, C.stream(X.device))
        return gX, g1, g2, g3, None, None


class GateRecurrent2dnoind(nn.Module):
    """Same constructor and call signature as the reference module."""

    def __init__(self, horizontal_, reverse_):
        super(GateRecurrent2dnoind, self).__init__()
        self.horizontal = horizontal_
        self.reverse = reverse_

    def forward(self, X, G1, G2, G3):
        return GateRecurrent2dnoindFunction.apply(X, G1, G2, G3, self.horizontal, self.reverse)
