from .builder import PROCESSORS, build_cost_processor  # noqa: F401
from .aggregators import AGGREGATORS, build_cost_aggregator, DeferredCost  # noqa: F401
from .utils.cat_fms import CAT_FUNCS  # noqa: F401
from .utils.dif_fms import DIF_FUNCS  # noqa: F401
from .utils.gwc_fms import GWC_FUNCS  # noqa: F401
from .utils.correlation1d_cost import COR_FUNCS  # noqa: F401
