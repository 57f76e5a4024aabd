"""AGGREGATORS + build_cost_aggregator (reference: cost_processors/aggregators/builder.py:8-29)."""
from .GCNet import GCAggregator
from .PSMNet import PSMAggregator
from .AcfNet import AcfAggregator
from .StereoNet import StereoNetAggregator

AGGREGATORS = {
    "GCNet": GCAggregator,
    "PSMNet": PSMAggregator,
    "AcfNet": AcfAggregator,
    "StereoNet": StereoNetAggregator,
    # 'DeepPruner' / 'AnyNet' (staged, model-specific processors) are outside this path's scope
}


def build_cost_aggregator(cfg):
    agg_type = cfg.model.cost_processor.cost_aggregator.type
    assert agg_type in AGGREGATORS, "cost_aggregator type not found, excepted: {}," \
                                    "but got {}".format(AGGREGATORS.keys(), agg_type)
    default_args = cfg.model.cost_processor.cost_aggregator.copy()
    default_args.pop('type')
    default_args.update(batch_norm=cfg.model.batch_norm)
    return AGGREGATORS[agg_type](**default_args)
