"""String -> class registry: the plugin mechanism behind `type:` keys in the configs.

Behavioural mirror of the reference's vendored mmcv registry
(`yanerf/utils/registry.py:10-50, 53-305`): `Registry.build(cfg)` pops `type`, looks the
class up (KeyError if unknown), calls it with the remaining keys and re-raises
constructor errors as `type(e)(f"{cls.__name__}: {e}")`.
"""
from __future__ import annotations

import inspect
from typing import Any, Callable, Dict, Optional


def build_from_cfg(cfg: Dict[str, Any], registry: "Registry", default_args: Optional[Dict[str, Any]] = None):
    if not isinstance(cfg, dict):
        raise TypeError(f"cfg must be a dict, but got {type(cfg)}")
    if not isinstance(registry, Registry):
        raise TypeError(f"registry must be a Registry, but got {type(registry)}")
    if default_args is not None and not isinstance(default_args, dict):
        raise TypeError(f"default_args must be a dict or None, but got {type(default_args)}")
    args = dict(cfg)
    for k, v in (default_args or {}).items():
        args.setdefault(k, v)
    if "type" not in args:
        raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}\n{default_args}')
    kind = args.pop("type")
    if isinstance(kind, str):
        cls = registry.get(kind)
        if cls is None:
            raise KeyError(f"{kind} is not in the {registry.name} registry")
    elif inspect.isclass(kind):
        cls = kind
    else:
        raise TypeError(f"type must be a str or valid type, but got {type(kind)}")
    try:
        return cls(**args)
    except Exception as e:  # keep the exception type, prepend the class name
        raise type(e)(f"{cls.__name__}: {e}")


class Registry:
    def __init__(self, name: str, build_func: Optional[Callable] = None):
        self._name = name
        self._modules: Dict[str, type] = {}
        self.build_func = build_func or build_from_cfg

    name = property(lambda self: self._name)
    module_dict = property(lambda self: self._modules)

    def __len__(self):
        return len(self._modules)

    def __contains__(self, key):
        return key in self._modules

    def __repr__(self):
        return f"{type(self).__name__}(name={self._name}, items={sorted(self._modules)})"

    def get(self, key: str):
        return self._modules.get(key)

    def build(self, *args, **kwargs):
        return self.build_func(*args, **kwargs, registry=self)

    def _register(self, cls, name=None, force=False):
        if not inspect.isclass(cls):
            raise TypeError(f"module must be a class, but got {type(cls)}")
        names = [name] if isinstance(name, str) else (name or [cls.__name__])
        for n in names:
            if not force and n in self._modules:
                raise KeyError(f"{n} is already registered in {self._name}")
            self._modules[n] = cls

    def register_module(self, name=None, force: bool = False, module=None):
        """Use as `@REG.register_module()` / `@REG.register_module(name=...)` or `REG.register_module(module=cls)`."""
        if not isinstance(force, bool):
            raise TypeError(f"force must be a boolean, but got {type(force)}")
        if not (name is None or isinstance(name, str) or (isinstance(name, (list, tuple)) and all(isinstance(n, str) for n in name))):
            raise TypeError(f"name must be None, a str or a sequence of str, but got {type(name)}")
        if module is not None:
            self._register(module, name, force)
            return module

        def deco(cls):
            self._register(cls, name, force)
            return cls

        return deco
