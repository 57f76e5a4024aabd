from .cost_processors import build_cost_processor  # noqa: F401
from .disp_predictors import build_disp_predictor  # noqa: F401
