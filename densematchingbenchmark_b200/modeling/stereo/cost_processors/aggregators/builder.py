"""Cost aggregators by config key (the names of dmb/modeling/stereo/cost_processors/aggregators/builder.py:8-15; the
staged, model-specific 'DeepPruner' / 'AnyNet' processors are outside this path's scope).  As in the reference
(:23-27) the model-level `batch_norm` switch is injected into the constructor arguments."""
from .....utils.registry import ctor_kwargs, lookup
from .AcfNet import AcfAggregator
from .GCNet import GCAggregator
from .PSMNet import PSMAggregator
from .StereoNet import StereoNetAggregator

AGGREGATORS = dict(GCNet=GCAggregator, PSMNet=PSMAggregator, AcfNet=AcfAggregator, StereoNet=StereoNetAggregator)


def build_cost_aggregator(cfg):
    section = cfg.model.cost_processor.cost_aggregator
    cls = lookup(AGGREGATORS, "cost_aggregator", section.type)
    return cls(**ctor_kwargs(section, batch_norm=cfg.model.batch_norm))
