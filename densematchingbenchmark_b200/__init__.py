"""B200-native cost-volume hot path for DenseMatchingBenchmark (dmb).

Mirror of the reference's component API for this path only:
    build_cost_processor / build_cost_aggregator / build_disp_predictor and the dict tables
    PROCESSORS, CAT_FUNCS, DIF_FUNCS, GWC_FUNCS, COR_FUNCS, AGGREGATORS, PREDICTORS,
    dmb.ops' GateRecurrent2dnoind, plus the GANet SGA / LGA layers.
Every forward runs hand-written sm_100a CUDA through the C ABI of csrc/libdmb_b200.so
(include/dmb_b200.h); there is no CPU or PyTorch fallback.  `install_into_dmb()` swaps these
implementations into an importable reference `dmb` package."""
from .modeling.stereo.cost_processors import (  # noqa: F401
    PROCESSORS, CAT_FUNCS, DIF_FUNCS, GWC_FUNCS, COR_FUNCS, AGGREGATORS, build_cost_processor, build_cost_aggregator,
    DeferredCost,
)
from .modeling.stereo.disp_predictors import PREDICTORS, build_disp_predictor  # noqa: F401
from .modeling.stereo.losses import StereoFocalLoss  # noqa: F401
from .utils.config import ConfigDict, load_config  # noqa: F401
from .dropin import install_into_dmb  # noqa: F401

__version__ = "0.1.0"
