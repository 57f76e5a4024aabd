"""DeferredCost: a full-resolution cost volume that has not been materialised yet.

The reference aggregators end with a x4 upsampling to [B,192,H,W] (401 MB fp32 per cost at
544x960; aggregators/PSMNet.py:75-88) that the soft-argmin predictor immediately reduces to a
[B,1,H,W] map.  `DeferredCost` carries the LOW-resolution cost and the upsampling recipe behind
a tensor of the reference's shape.  Our predictors recognise it and run the fused
upsample+soft-argmin kernel (8 MB of traffic); any other consumer (losses, Cmn, .cpu(), indexing,
arithmetic ...) transparently gets the dense tensor, materialised once by the CUDA upsampler."""
import torch
from torch.utils._pytree import tree_map

from .....ops import functional as F_


class DeferredCost(torch.Tensor):

    @staticmethod
    def __new__(cls, low, out_dhw, mode, up_weight=None):
        B = low.shape[0]
        D, H, W = out_dhw
        r = torch.Tensor._make_wrapper_subclass(cls, (B, D, H, W), dtype=torch.float32, device=low.device,
                                                requires_grad=False)
        r._low = low
        r._out_dhw = tuple(out_dhw)
        r._mode = mode
        r._up_weight = up_weight
        r._dense = None
        return r

    def __repr__(self):
        return "DeferredCost(shape=%s, mode=%s, materialised=%s)" % (tuple(self.shape), self._mode,
                                                                     self._dense is not None)

    def materialize(self):
        if self._dense is None:
            self._dense, _ = F_.upsample_regress(self._low, self._out_dhw, self._mode, self._up_weight,
                                                 want_cost=True, want_disp=False)
        return self._dense

    def regress(self, **kw):
        """Fused upsample + soft-argmin -> [B,1,H,W]."""
        if self._dense is not None:
            return F_.soft_argmin(self._dense, **kw)
        _, disp = F_.upsample_regress(self._low, self._out_dhw, self._mode, self._up_weight,
                                      want_cost=False, want_disp=True, **kw)
        return disp

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        def unwrap(t):
            return t.materialize() if isinstance(t, DeferredCost) else t
        return func(*tree_map(unwrap, args), **tree_map(unwrap, kwargs or {}))
