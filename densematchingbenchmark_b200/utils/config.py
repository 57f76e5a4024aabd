"""Minimal attribute-dict config (the slice of `mmcv.Config` the builders touch:
`.get/.copy/.pop/in`, attribute access -- dmb/modeling/stereo/cost_processors/builder.py:23-31).
Any object with the same behaviour (a real mmcv Config) works with the builders too."""
import types


class ConfigDict(dict):

    def __init__(self, *args, **kwargs):
        super(ConfigDict, self).__init__(*args, **kwargs)
        for k in list(self.keys()):
            self[k] = _wrap(self[k])

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = _wrap(value)

    def copy(self):
        return ConfigDict({k: (v.copy() if isinstance(v, ConfigDict) else v) for k, v in self.items()})


def _wrap(v):
    if isinstance(v, dict) and not isinstance(v, ConfigDict):
        return ConfigDict(v)
    if isinstance(v, (list, tuple)):
        return type(v)(_wrap(i) for i in v)
    return v


def load_config(path):
    """Execute a dmb python config file (e.g. configs/PSMNet/scene_flow.py) -> ConfigDict."""
    scope = {"__file__": path}
    with open(path) as fh:
        exec(compile(fh.read(), path, "exec"), scope)
    return ConfigDict({k: v for k, v in scope.items()
                       if not k.startswith("_") and not isinstance(v, types.ModuleType)})
