"""Config loading for the `configs/nerf/*.yml` files and the `.py` test configs.

Own implementation of the subset of the reference's vendored mmcv `Config`
(`yanerf/utils/config.py:35-48, 173-259, 320-324, 556-693`) that `scripts/run.py` and the
pipeline need: yaml / json / python files, `_base_` inheritance, `{{ fileDirname }}`-style
template variables, `custom_imports`, `--cfg_options a.b=c` deep merges, attribute access.
"""
from __future__ import annotations

import ast
import copy
import importlib
import json
import os
import re
import types
from argparse import Action
from typing import Any, Dict, Optional

import yaml

BASE_KEY = "_base_"
DELETE_KEY = "_delete_"


class ConfigDict(dict):
    """dict with attribute access; nested dicts (also inside lists/tuples) are wrapped on the way in."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    def __delattr__(self, name):
        try:
            del self[name]
        except KeyError:
            raise AttributeError(name)

    def update(self, *args, **kwargs):
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def setdefault(self, k, default=None):
        if k not in self:
            self[k] = default
        return self[k]

    def copy(self):
        return type(self)(self)

    def __deepcopy__(self, memo):
        return type(self)({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def to_dict(self) -> Dict[str, Any]:
        def plain(v):
            if isinstance(v, dict):
                return {k: plain(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return type(v)(plain(x) for x in v)
            return v

        return plain(self)


def _substitute_templates(text: str, filename: str) -> str:
    base = os.path.basename(filename)
    values = {
        "fileDirname": os.path.dirname(filename),
        "fileBasename": base,
        "fileBasenameNoExtension": os.path.splitext(base)[0],
        "fileExtname": os.path.splitext(base)[1],
    }
    for key, val in values.items():
        text = re.sub(r"\{\{\s*" + key + r"\s*\}\}", val.replace("\\", "/"), text)
    return text


def _merge(child: Dict[str, Any], base: Dict[str, Any], allow_list_keys: bool = False) -> Dict[str, Any]:
    """Deep-merge `child` into a copy of `base` (`_delete_: true` replaces instead of merging)."""
    out = copy.deepcopy(base)
    for k, v in child.items():
        if allow_list_keys and k.isdigit() and isinstance(out, list):
            k = int(k)
            if k >= len(out):
                raise KeyError(f"Index {k} exceeds the length of list {out}")
            out[k] = _merge(v, out[k], allow_list_keys) if isinstance(v, dict) else v
        elif isinstance(v, dict):
            if k in out and not v.pop(DELETE_KEY, False):
                if not isinstance(out[k], (dict, list) if allow_list_keys else dict):
                    raise TypeError(
                        f"{k}={v} in child config cannot inherit from base because {k} is a dict in the child "
                        f"config but is of type {type(out[k])} in base config. You may set `{DELETE_KEY}=True`."
                    )
                out[k] = _merge(v, out[k], allow_list_keys)
            else:
                out[k] = ConfigDict(v) if not isinstance(out, list) else v
        else:
            out[k] = v
    return out


def _load_file(filename: str) -> Dict[str, Any]:
    filename = os.path.abspath(os.path.expanduser(filename))
    if not os.path.isfile(filename):
        raise FileNotFoundError(f'file "{filename}" does not exist')
    ext = os.path.splitext(filename)[1]
    if ext not in (".py", ".json", ".yaml", ".yml"):
        raise IOError("Only py/yml/yaml/json type are supported now!")
    with open(filename, "r", encoding="utf-8") as f:
        text = _substitute_templates(f.read(), filename)
    if ext == ".py":
        ast.parse(text, filename)  # surface syntax errors with the file name
        scope = _exec_py(text, filename)
        cfg = {
            k: v for k, v in scope.items()
            if not k.startswith("__") and not isinstance(v, (types.ModuleType, types.FunctionType)) and not isinstance(v, type)
        }
    elif ext == ".json":
        cfg = json.loads(text)
    else:
        cfg = yaml.safe_load(text) or {}
    if BASE_KEY in cfg:
        here = os.path.dirname(filename)
        bases = cfg.pop(BASE_KEY)
        bases = bases if isinstance(bases, list) else [bases]
        merged: Dict[str, Any] = {}
        for b in bases:
            loaded = _load_file(os.path.join(here, b))
            dup = merged.keys() & loaded.keys()
            if dup:
                raise KeyError(f"Duplicate key is not allowed among bases. Duplicate keys: {dup}")
            merged.update(loaded)
        cfg = _merge(cfg, merged)
    return cfg


def _exec_py(text: str, filename: str) -> Dict[str, Any]:
    scope: Dict[str, Any] = {"__file__": filename, "__name__": "__yanerf_config__"}
    exec(compile(text, filename, "exec"), scope)
    return scope


class Config:
    """`Config.fromfile(path)` -> attribute-style access to the parsed file."""

    def __init__(self, cfg_dict: Optional[Dict[str, Any]] = None, filename: Optional[str] = None):
        if cfg_dict is None:
            cfg_dict = {}
        elif not isinstance(cfg_dict, dict):
            raise TypeError(f"cfg_dict must be a dict, but got {type(cfg_dict)}")
        object.__setattr__(self, "_cfg_dict", ConfigDict(cfg_dict))
        object.__setattr__(self, "_filename", filename)

    @staticmethod
    def fromfile(filename: str, import_custom_modules: bool = True) -> "Config":
        cfg = _load_file(str(filename))
        if import_custom_modules and cfg.get("custom_imports"):
            ci = cfg["custom_imports"]
            for mod in ci.get("imports", []):
                try:
                    importlib.import_module(mod)
                except ImportError:
                    if not ci.get("allow_failed_imports", False):
                        raise
        return Config(cfg, filename=str(filename))

    filename = property(lambda self: self._filename)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = value

    __setitem__ = __setattr__

    def __contains__(self, name):
        return name in self._cfg_dict

    def __iter__(self):
        return iter(self._cfg_dict)

    def __len__(self):
        return len(self._cfg_dict)

    def __repr__(self):
        return f"Config (path: {self._filename}): {self._cfg_dict!r}"

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)

    def to_dict(self):
        return self._cfg_dict.to_dict()

    @property
    def pretty_text(self) -> str:
        return yaml.safe_dump(self.to_dict(), sort_keys=False)

    def dump(self, file: Optional[str] = None) -> Optional[str]:
        text = self.pretty_text
        if file is None:
            return text
        with open(file, "w", encoding="utf-8") as f:
            f.write(text)
        return None

    def merge_from_dict(self, options: Dict[str, Any], allow_list_keys: bool = True) -> None:
        """`{'a.b': 1, 'c.0.d': 2}` -> deep merge (list indices allowed), as `--cfg_options` does."""
        nested: Dict[str, Any] = {}
        for full_key, v in options.items():
            d = nested
            keys = full_key.split(".")
            for k in keys[:-1]:
                d = d.setdefault(k, {})
            d[keys[-1]] = v
        merged = _merge(nested, self._cfg_dict, allow_list_keys=allow_list_keys)
        object.__setattr__(self, "_cfg_dict", ConfigDict(merged))


class DictAction(Action):
    """argparse action: `--cfg_options a.b=1 c=[1,2] d=(x,y)` -> dict with parsed python-ish values."""

    @staticmethod
    def _scalar(val: str):
        for cast in (int, float):
            try:
                return cast(val)
            except ValueError:
                pass
        low = val.lower()
        if low in ("true", "false"):
            return low == "true"
        if low in ("none", "null"):
            return None
        return val

    @classmethod
    def _parse(cls, val: str):
        val = val.strip().strip("'\"")
        is_tuple = val.startswith("(") and val.endswith(")")
        if is_tuple or (val.startswith("[") and val.endswith("]")):
            inner, items, depth, cur = val[1:-1], [], 0, ""
            for ch in inner:
                if ch == "," and depth == 0:
                    items.append(cur)
                    cur = ""
                    continue
                depth += ch in "(["
                depth -= ch in ")]"
                cur += ch
            if cur.strip():
                items.append(cur)
            out = [cls._parse(x) for x in items]
            return tuple(out) if is_tuple else out
        if "," in val:
            return [cls._parse(x) for x in val.split(",")]
        return cls._scalar(val)

    def __call__(self, parser, namespace, values, option_string=None):
        options = {}
        for kv in values:
            key, val = kv.split("=", maxsplit=1)
            options[key] = self._parse(val)
        setattr(namespace, self.dest, options)
