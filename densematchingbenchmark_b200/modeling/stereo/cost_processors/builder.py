"""PROCESSORS + build_cost_processor (reference: cost_processors/builder.py:21-107)."""
import torch.nn as nn

from .utils.cat_fms import CAT_FUNCS
from .utils.dif_fms import DIF_FUNCS
from .utils.gwc_fms import GWC_FUNCS
from .utils.correlation1d_cost import COR_FUNCS
from .aggregators import build_cost_aggregator


class CostProcessor(nn.Module):

    def forward(self, *input):
        raise NotImplementedError


class _VolumeThenAggregate(CostProcessor):
    """Shared body of the Cat/Dif/Gwc processors: `func(ref, tgt, disp_sample=..., **args)` then
    `self.aggregator(raw)` (builder.py:23-40)."""
    table = None

    def __init__(self, cfg):
        super(_VolumeThenAggregate, self).__init__()
        comp = cfg.model.cost_processor.cost_computation
        self.func = self.table[comp.get('type', 'default')]
        self.default_args = comp.copy()
        self.default_args.pop('type')
        self.aggregator = build_cost_aggregator(cfg)

    def forward(self, ref_fms, tgt_fms, disp_sample=None):
        raw_cost = self.func(ref_fms, tgt_fms, disp_sample=disp_sample, **self.default_args)
        return self.aggregator(raw_cost)


class CatCostProcessor(_VolumeThenAggregate):
    table = CAT_FUNCS

    @property
    def cat_func(self):
        return self.func

    def forward(self, ref_fms, tgt_fms, disp_sample=None):
        # fused route: default cat volume feeding a tensor-core trunk is written once, in the trunk's own
        # layout (the reference materialises 401 MB of fp32 and the aggregator re-reads it)
        if self.func is CAT_FUNCS['default'] and hasattr(self.aggregator, "blocked_cat_volume"):
            blk = self.aggregator.blocked_cat_volume(ref_fms, tgt_fms, **self.default_args)
            if blk is not None:
                return self.aggregator(blk)
        return super(CatCostProcessor, self).forward(ref_fms, tgt_fms, disp_sample)


class DifCostProcessor(_VolumeThenAggregate):
    table = DIF_FUNCS

    @property
    def dif_func(self):
        return self.func


class GwcCostProcessor(_VolumeThenAggregate):
    """NEW key 'GroupWiseCorrelation' (no counterpart in the reference snapshot)."""
    table = GWC_FUNCS


class CorCostProcessor(_VolumeThenAggregate):
    """'Correlation' (builder.py:67-87): correlation1d_cost, then the configured aggregator."""
    table = COR_FUNCS

    @property
    def cor_func(self):
        return self.func


PROCESSORS = {
    'Difference': DifCostProcessor,
    'Concatenation': CatCostProcessor,
    'Correlation': CorCostProcessor,
    'GroupWiseCorrelation': GwcCostProcessor,
}


def build_cost_processor(cfg):
    from ....utils.registry import lookup
    return lookup(PROCESSORS, "cost_processor", cfg.model.cost_processor.type)(cfg=cfg)
