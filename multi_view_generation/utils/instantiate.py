"""A small `_target_` instantiator for environments without Hydra / OmegaConf (this image has neither): loads a YAML model config of
the reference's shape (configs/model/stage_2.yaml: nested `_target_` class paths, `${...}` interpolations against the root config) and
builds the objects the way `hydra.utils.instantiate` does - children first, `_target_` dicts become constructor calls with the remaining
keys as keyword arguments.  With Hydra installed the reference's own launcher works unchanged; this exists so the drop-in claim can be
tested offline (tests/test_hydra_targets_*.py)."""
import importlib
import re
from typing import Any, Mapping

import yaml

_INTERP = re.compile(r"\$\{([^}]+)\}")


def _lookup(root: Mapping, dotted: str):
    node: Any = root
    for part in dotted.split("."):
        node = node[int(part)] if isinstance(node, list) else node[part]
    return node


def resolve(node: Any, root: Mapping) -> Any:
    """OmegaConf-style `${a.b}` interpolation (whole-value references keep their type, embedded ones are formatted into the string)."""
    if isinstance(node, dict):
        return {k: resolve(v, root) for k, v in node.items()}
    if isinstance(node, list):
        return [resolve(v, root) for v in node]
    if isinstance(node, str):
        m = _INTERP.fullmatch(node)
        if m:
            return resolve(_lookup(root, m.group(1)), root)
        return _INTERP.sub(lambda mm: str(resolve(_lookup(root, mm.group(1)), root)), node)
    return node


def locate(path: str):
    module, _, name = path.rpartition(".")
    return getattr(importlib.import_module(module), name)


def instantiate(cfg: Any, **overrides) -> Any:
    """hydra.utils.instantiate for plain dicts: recursive, `_target_` -> call; keys starting with `_` other than `_target_` are ignored."""
    if isinstance(cfg, list):
        return [instantiate(v) for v in cfg]
    if not isinstance(cfg, dict):
        return cfg
    kwargs = {k: instantiate(v) for k, v in cfg.items() if not k.startswith("_")}
    kwargs.update(overrides)
    if "_target_" not in cfg:
        return kwargs
    return locate(cfg["_target_"])(**kwargs)


def load_yaml(path, **root_overrides) -> dict:
    with open(path) as fh:
        root = yaml.safe_load(fh)
    root.update(root_overrides)
    return resolve(root, root)
