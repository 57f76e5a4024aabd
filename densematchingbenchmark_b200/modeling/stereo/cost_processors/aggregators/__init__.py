from .builder import AGGREGATORS, build_cost_aggregator  # noqa: F401
from .deferred import DeferredCost  # noqa: F401
