from .config import Config, ConfigDict, DictAction
from .registry import Registry, build_from_cfg

__all__ = ["Config", "ConfigDict", "DictAction", "Registry", "build_from_cfg"]
