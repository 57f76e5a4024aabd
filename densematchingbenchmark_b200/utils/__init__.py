from .config import ConfigDict, load_config  # noqa: F401
