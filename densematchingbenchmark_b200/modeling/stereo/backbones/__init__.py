from .PSMNet import PSMNetBackbone  # noqa: F401
