"""A small Hydra/OmegaConf-compatible config layer (hydra-core and omegaconf are not installable in
this image).  It covers what the reference's entry point uses (SURVEY.md section 5 "Config / flags"):

  * defaults-list composition of config groups with `_self_` (default_ff_ppo.yaml:1-7),
  * `group=option` and dotted `a.b.c=value` command-line overrides, `+key=value` additions,
  * `${a.b}` interpolation, attribute access, runtime mutation (OmegaConf.set_struct(cfg, False)),
  * `_target_` instantiation (hydra.utils.instantiate, ff_ppo.py:439-444) -- `stoix.` targets resolve
    to this package so the reference's network configs work verbatim.
"""
from __future__ import annotations

import copy
import importlib
import re
from pathlib import Path
from typing import Any, Dict, Iterable, List, Optional

import yaml

CONFIG_ROOT = Path(__file__).resolve().parent / "configs"
_INTERP = re.compile(r"\$\{([^}]+)\}")


class DictConfig(dict):
    """dict with attribute access and `${}` interpolation resolved against the root on read."""

    def __init__(self, data: Optional[Dict[str, Any]] = None, _root: Optional["DictConfig"] = None):
        super().__init__()
        object.__setattr__(self, "_root", _root if _root is not None else self)
        for k, v in (data or {}).items():
            self[k] = v

    def _wrap(self, v: Any) -> Any:
        root = object.__getattribute__(self, "_root")
        if isinstance(v, DictConfig):
            _reroot(v, root)
            return v
        if isinstance(v, dict):
            return DictConfig(v, _root=root)
        if isinstance(v, (list, tuple)):
            return [self._wrap(x) for x in v]
        return v

    def __setitem__(self, k: str, v: Any) -> None:
        super().__setitem__(k, self._wrap(v))

    def __getitem__(self, k: str) -> Any:
        return self._resolve(super().__getitem__(k))

    def get(self, k: str, default: Any = None) -> Any:
        return self[k] if k in self else default

    def __getattr__(self, k: str) -> Any:
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k: str, v: Any) -> None:
        self[k] = v

    def __deepcopy__(self, memo):
        return DictConfig(copy.deepcopy(to_container(self, resolve=False), memo))

    def _resolve(self, v: Any) -> Any:
        if not isinstance(v, str) or "${" not in v:
            return v
        root = object.__getattribute__(self, "_root")
        whole = _INTERP.fullmatch(v)
        if whole:
            return select(root, whole.group(1))
        return _INTERP.sub(lambda m: str(select(root, m.group(1))), v)


def _reroot(node: DictConfig, root: DictConfig) -> None:
    object.__setattr__(node, "_root", root)
    for v in dict.values(node):
        if isinstance(v, DictConfig):
            _reroot(v, root)
        elif isinstance(v, list):
            for x in v:
                if isinstance(x, DictConfig):
                    _reroot(x, root)


def select(cfg: DictConfig, dotted: str) -> Any:
    node: Any = cfg
    for part in dotted.split("."):
        node = node[part]
    return node


def set_path(cfg: DictConfig, dotted: str, value: Any) -> None:
    parts = dotted.split(".")
    node = cfg
    for p in parts[:-1]:
        if p not in node or not isinstance(dict.__getitem__(node, p), dict):
            node[p] = {}
        node = dict.__getitem__(node, p)
    node[parts[-1]] = value


def to_container(cfg: Any, resolve: bool = True) -> Any:
    if isinstance(cfg, DictConfig):
        return {k: to_container(cfg[k] if resolve else dict.__getitem__(cfg, k), resolve) for k in cfg}
    if isinstance(cfg, (list, tuple)):
        return [to_container(x, resolve) for x in cfg]
    return cfg


def _merge(dst: Dict[str, Any], src: Dict[str, Any]) -> Dict[str, Any]:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


class _Loader(yaml.SafeLoader):
    """SafeLoader whose float resolver also accepts exponents without a dot (`3e-4`, `1e7`): PyYAML follows YAML 1.1
    and reads those as strings, OmegaConf (the reference's loader) as floats."""


_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                    |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                    |\.[0-9][0-9_]*(?:[eE][-+]?[0-9]+)?
                    |[-+]?\.(?:inf|Inf|INF)
                    |\.(?:nan|NaN|NAN))$""", re.X),
    list("-+0123456789."),
)


def _load_yaml(path: Path) -> Dict[str, Any]:
    if not path.exists():
        raise FileNotFoundError(f"config file not found: {path}")
    return yaml.load(path.read_text(), Loader=_Loader) or {}


def _parse_value(text: str) -> Any:
    try:
        return yaml.load(text, Loader=_Loader)
    except yaml.YAMLError:
        return text


def compose(config_name: str = "default_ff_ppo", overrides: Iterable[str] = (), config_dir: str = "default/anakin",
            root: Optional[Path] = None) -> DictConfig:
    """Compose `<root>/<config_dir>/<config_name>.yaml` with its defaults list and apply overrides."""
    root = Path(root) if root is not None else CONFIG_ROOT
    name = config_name[:-5] if config_name.endswith(".yaml") else config_name
    primary = _load_yaml(root / config_dir / f"{name}.yaml")
    defaults: List[Any] = primary.pop("defaults", [])
    primary.pop("hydra", None)
    overrides = list(overrides)
    groups = {list(d.keys())[0]: list(d.values())[0] for d in defaults if isinstance(d, dict)}
    value_overrides = []
    for ov in overrides:
        key, _, val = ov.partition("=")
        add = key.startswith("+")
        key = key.lstrip("+~")
        if key in groups and "." not in key and not add:
            groups[key] = val  # config-group selection, e.g. env=synthetic/box
        else:
            value_overrides.append((key, _parse_value(val)))
    merged: Dict[str, Any] = {}
    seen_self = False
    for d in defaults:
        if d == "_self_":
            _merge(merged, primary)
            seen_self = True
        elif isinstance(d, dict):
            g = list(d.keys())[0]
            opt = groups[g]
            if opt is None:
                continue
            _merge(merged, {g: _load_yaml(root / g / f"{opt}.yaml")})
    if not seen_self:
        _merge(merged, primary)
    cfg = DictConfig(merged)
    for key, val in value_overrides:
        set_path(cfg, key, val)
    return cfg


_ALIASES = (("stoix.", "stoix_b200."),)


def _locate(target: str) -> Any:
    for old, new in _ALIASES:
        if target.startswith(old):
            target = new + target[len(old):]
            break
    module, _, attr = target.rpartition(".")
    return getattr(importlib.import_module(module), attr)


def instantiate(node: Any, *args: Any, **kwargs: Any) -> Any:
    """hydra.utils.instantiate for `_target_` nodes (recursive, kwargs override config values)."""
    if not isinstance(node, dict) or "_target_" not in node:
        raise ValueError("instantiate() needs a config node with a _target_")
    params = {}
    for k in node:
        if k in ("_target_", "_partial_"):
            continue
        v = node[k]
        params[k] = instantiate(v) if isinstance(v, dict) and "_target_" in v else (to_container(v) if isinstance(v, (dict, list)) else v)
    params.update(kwargs)
    return _locate(node["_target_"])(*args, **params)
