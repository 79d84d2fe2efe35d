"""The reference's plug-in mechanism without Hydra: compose `cfg/config.yaml` + `cfg/model/*.yaml` + `cfg/exp/*.yaml`,
resolve `${...}` interpolations and build objects from `_target_` nodes -- with the reference's targets swapped for the
B200 drop-ins.  This is what "cfg/exp/* drives it unchanged" means in practice:

    cfg = compose("/path/to/RemFx/cfg", exp="5-5_full")          # scripts/train.py:9  @hydra.main + `+exp=5-5_full`
    model = instantiate(cfg["model"])                            # scripts/train.py:17 hydra.utils.instantiate(cfg.model)
    -> remfx_b200.train.RemFX(network=remfx_b200.models.DemucsModel(...))

Only the subset of Hydra / OmegaConf that RemFx's config tree uses is implemented: a `defaults` list with `_self_`,
`group: name`, `group: null` and `override /group: name`; `# @package _global_` files (every file under cfg/model and
cfg/exp is one); `${a.b}` (node or in-string), `${oc.env:VAR[,default]}` and `${now:%fmt}`; dotted `key=value` overrides.
Host logic only (PyYAML); hydra / omegaconf are not installed in this image.
"""
from __future__ import annotations

import copy
import datetime
import importlib
import os
import re
from typing import Any, Dict, Mapping, Optional

import yaml

# reference `_target_` -> drop-in.  Anything else is imported as written (and fails loudly if its package is absent).
TARGET_MAP: Dict[str, str] = {
    "remfx.models.RemFX": "remfx_b200.train.RemFX",
    "remfx.models.OpenUnmixModel": "remfx_b200.models.OpenUnmixModel",
    "remfx.models.TCNModel": "remfx_b200.models.TCNModel",
    "remfx.models.DemucsModel": "remfx_b200.models.DemucsModel",
    "remfx.classifier.Cnn14": "remfx_b200.classifier.Cnn14",
    "remfx.models.RemFXChainInference": "remfx_b200.chain.RemFXChainInference",
}
# reference targets that are knowingly not provided (asteroid is absent, no oracle)
UNSUPPORTED = {"remfx.models.DCUNetModel": "DCUNet (asteroid) has no drop-in: SURVEY.md 8(f) N1",
               "remfx.models.DPTNetModel": "DPTNet (asteroid) has no drop-in"}


class _Loader(yaml.SafeLoader):
    pass


# YAML 1.1 reads `1e-4` as a string; OmegaConf reads it as a float (cfg/model/*.yaml: `lr: 1e-4`)
_Loader.add_implicit_resolver(
    "tag:yaml.org,2002:float",
    re.compile(r"^[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?$|^[-+]?\.(?:inf|Inf|INF)$|^\.(?:nan|NaN|NAN)$"),
    list("-+0123456789."))


def _load(path: str) -> dict:
    with open(path) as fh:
        return yaml.load(fh, Loader=_Loader) or {}


def _merge(dst: dict, src: Mapping) -> dict:
    for k, v in src.items():
        if isinstance(v, Mapping) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


def compose(cfg_dir: str, exp: Optional[str] = None, groups: Optional[Mapping[str, Optional[str]]] = None,
            overrides: Optional[Mapping[str, Any]] = None, resolve_now: bool = True) -> dict:
    """Hydra-style composition of the RemFx config tree.  `groups` plays the command line's `model=demucs`; `exp` its
    `+exp=NAME`; `overrides` its dotted `a.b=c`."""
    root = _load(os.path.join(cfg_dir, "config.yaml"))
    choice: Dict[str, Optional[str]] = {}
    for d in root.pop("defaults", []):
        if isinstance(d, Mapping):
            for g, name in d.items():
                choice[g] = name
    exp_cfg = None
    if exp is not None:
        exp_cfg = _load(os.path.join(cfg_dir, "exp", exp + ".yaml"))
        for d in exp_cfg.pop("defaults", []):
            if isinstance(d, Mapping):
                for g, name in d.items():
                    choice[g.replace("override", "").strip().lstrip("/")] = name
    for g, name in (groups or {}).items():
        choice[g] = name
    cfg = dict(root)  # `_self_` comes first in cfg/config.yaml:1-5: groups then override it
    for g, name in choice.items():
        if name is None:
            continue
        path = os.path.join(cfg_dir, g, f"{name}.yaml")
        if not os.path.exists(path):
            raise FileNotFoundError(f"config group '{g}' has no option '{name}' ({path})")
        with open(path) as fh:
            is_global = "@package _global_" in fh.readline()
        node = _load(path)
        _merge(cfg, node if is_global else {g: node})
    if exp_cfg is not None:
        _merge(cfg, exp_cfg)  # every cfg/exp file is `# @package _global_`
    for key, val in (overrides or {}).items():
        cur = cfg
        parts = key.lstrip("+").split(".")
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = val
    return resolve(cfg) if resolve_now else cfg


_INTERP = re.compile(r"\$\{([^${}]+)\}")


def resolve(cfg: dict) -> dict:
    """Resolve every `${...}` in place (nodes referenced as a whole are deep-copied, as OmegaConf does on to_container)."""
    now = datetime.datetime.now()

    def lookup(expr: str, stack):
        expr = expr.strip()
        if expr.startswith("oc.env:"):
            name, _, default = expr[len("oc.env:"):].partition(",")
            if name in os.environ:
                return os.environ[name]
            if _:
                return default
            # OmegaConf resolves lazily and only fails when the node is READ (cfg/config.yaml:55 needs DATASET_ROOT for the
            # datamodule only): keep the interpolation text in place instead of failing the whole composition
            return "${" + expr + "}"
        if expr.startswith("now:"):
            return now.strftime(expr[len("now:"):])
        if expr in stack:
            raise ValueError(f"interpolation cycle through '{expr}'")
        cur: Any = cfg
        for p in expr.split("."):
            if not isinstance(cur, Mapping) or p not in cur:
                raise KeyError(f"interpolation key '{expr}' not found")
            cur = cur[p]
        return walk(copy.deepcopy(cur), stack + (expr,))

    def walk(node, stack=()):
        if isinstance(node, dict):
            for k in list(node):
                node[k] = walk(node[k], stack)
            return node
        if isinstance(node, list):
            return [walk(v, stack) for v in node]
        if isinstance(node, str) and "${" in node:
            m = _INTERP.fullmatch(node)
            if m:
                return lookup(m.group(1), stack)
            for _ in range(8):  # nested in-string interpolations; unresolved ${oc.env:...} text stays as it is
                new = _INTERP.sub(lambda mm: str(lookup(mm.group(1), stack)), node)
                if new == node:
                    break
                node = new
            return node
        return node

    return walk(cfg)


def _locate(path: str):
    mod, _, name = path.rpartition(".")
    return getattr(importlib.import_module(mod), name)


def instantiate(node: Any, target_map: Optional[Mapping[str, str]] = None, **kwargs):
    """hydra.utils.instantiate for resolved plain-dict configs: builds `_target_` nodes recursively (children first)."""
    tmap = TARGET_MAP if target_map is None else target_map
    if isinstance(node, list):
        return [instantiate(v, tmap) for v in node]
    if not isinstance(node, Mapping):
        return node
    if "_target_" not in node:
        return {k: instantiate(v, tmap) for k, v in node.items()}
    target = node["_target_"]
    if target in UNSUPPORTED and target not in tmap:
        raise NotImplementedError(f"_target_ {target}: {UNSUPPORTED[target]}")
    cls = _locate(tmap.get(target, target))
    args = {k: instantiate(v, tmap) for k, v in node.items() if k != "_target_"}
    args.update(kwargs)
    return cls(**args)
