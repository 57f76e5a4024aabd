"""Disparity predictors by config key (the names of dmb/modeling/stereo/disp_predictors/builder.py:5-9; `FASTER` is the
default there, :13)."""
from ....utils.registry import ctor_kwargs, lookup
from .faster_soft_argmin import FasterSoftArgmin
from .local_soft_argmin import LocalSoftArgmin
from .soft_argmin import SoftArgmin

PREDICTORS = dict(DEFAULT=SoftArgmin, FASTER=FasterSoftArgmin, LOCAL=LocalSoftArgmin)


def build_disp_predictor(cfg):
    section = cfg.model.disp_predictor
    cls = lookup(PREDICTORS, "disparity predictor", section.get("type", "FASTER"))
    return cls(**ctor_kwargs(section))
