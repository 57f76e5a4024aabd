"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Read-only import harness for the reference (`/root/reference/dmb`) in THIS container.

The reference cannot be imported as shipped: `dmb.modeling` pulls un-vendored third
party packages at import time:
  spatial_correlation_sampler  (dmb/modeling/stereo/cost_processors/utils/correlation1d_cost.py:5)
  mmcv / apex                  (dmb/apis/train.py:23-27, dmb/utils/dist_utils.py:9-13)
  detectron2, gaterecurrent2dnoind_cuda (dmb/ops/__init__.py:1-2,
                                dmb/ops/spn/functions/gaterecurrent2dnoind.py:3-6)
This module installs inert `sys.modules` stubs for those names, puts the read-only tree on
`sys.path` (no bytecode is written) and exposes helpers that build the reference's own
modules from the reference's own config files.  It is used by `oracle/make_golden.py`
(fixture generation) and by the `ref`-marked tests that only run where `/root/reference`
exists.  `/root/reference` is absent on the GPU box; nothing reachable from `-m gpu`
tests, `smoke()` or `bench.py` imports this file.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DMB_REFERENCE_ROOT", "/root/reference")

_STUB_NAMES = [
    "spatial_correlation_sampler",
    "mmcv", "mmcv.runner", "mmcv.runner.hooks", "mmcv.parallel", "mmcv.utils",
    "apex", "apex.amp", "apex.parallel",
    "detectron2", "detectron2.layers",
    "gaterecurrent2dnoind_cuda",
    "thop", "imageio", "matplotlib", "matplotlib.pyplot", "matplotlib.cm",
    "tensorboardX", "skimage", "skimage.io", "cv2", "png",
]


class _Anything:
    """Attribute sink: any attribute access / call returns another sink."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything


class ConfigDict(dict):
    """Dict with attribute access: the slice of `mmcv.Config` the builders use
    (`.get/.copy/.pop/in/getattr`; e.g. dmb/modeling/stereo/cost_processors/builder.py:23-31)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        for key, val in list(self.items()):
            self[key] = self._wrap(val)

    @classmethod
    def _wrap(cls, val):
        if isinstance(val, dict) and not isinstance(val, ConfigDict):
            return cls(val)
        if isinstance(val, (list, tuple)):
            return type(val)(cls._wrap(v) for v in val)
        return val

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = self._wrap(value)

    def copy(self):
        return ConfigDict({k: (v.copy() if isinstance(v, ConfigDict) else v) for k, v in self.items()})


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dmb"))


def install():
    """Make `import dmb` resolve to the read-only reference tree."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    for name in _STUB_NAMES:
        if name not in sys.modules:
            mod = _StubModule(name)
            mod.__path__ = []  # behave as a package
            sys.modules[name] = mod
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    warnings.filterwarnings("ignore", message=".*align_corners.*")
    return importlib.import_module("dmb")


def load_config(rel_path):
    """exec() a reference config file (e.g. 'configs/PSMNet/scene_flow.py') -> ConfigDict."""
    path = os.path.join(REFERENCE_ROOT, rel_path)
    scope = {"__file__": path}
    with open(path) as fh:
        exec(compile(fh.read(), path, "exec"), scope)
    cfg = {k: v for k, v in scope.items() if not k.startswith("_") and not isinstance(v, types.ModuleType)}
    return ConfigDict(cfg)


def ref_model(cfg):
    install()
    from dmb.modeling import build_model
    return build_model(cfg)


def ref_cost_processor(cfg):
    install()
    from dmb.modeling.stereo.cost_processors import build_cost_processor
    return build_cost_processor(cfg)


def ref_disp_predictor(cfg):
    install()
    from dmb.modeling.stereo.disp_predictors import build_disp_predictor
    return build_disp_predictor(cfg)


def ref_funcs():
    """The raw cost-volume builder tables of the reference."""
    install()
    from dmb.modeling.stereo.cost_processors.utils.cat_fms import CAT_FUNCS
    from dmb.modeling.stereo.cost_processors.utils.dif_fms import DIF_FUNCS
    return CAT_FUNCS, DIF_FUNCS
