"""``_target_`` instantiation of the reference's YAML trees without hydra.

The reference builds its dataset/generator objects with ``hydra.utils.instantiate`` over
``configs/**.yaml`` (``fetalsyngen/test.py:8-12``, ``configs/dataset/generator/default.yaml``).
hydra/omegaconf are not in the target image, so this module provides the three things those
configs use: recursive ``_target_`` construction, ``defaults:`` composition of one nested
group, and relative ``${..key}`` interpolation.  ``aliases=True`` routes every
``fetalsyngen.*`` target to its ``fetalsyngen_b200.*`` counterpart, which is how the reference's
own YAML text drives the B200 path.
"""
from __future__ import annotations

import importlib
from pathlib import Path

import yaml

ALIAS_PREFIXES = {
    "fetalsyngen.generator.model.": "fetalsyngen_b200.generator.model.",
    "fetalsyngen.generator.intensity.rand_gmm.": "fetalsyngen_b200.generator.intensity.rand_gmm.",
    "fetalsyngen.generator.deformation.affine_nonrigid.": "fetalsyngen_b200.generator.deformation.affine_nonrigid.",
    "fetalsyngen.generator.augmentation.synthseg.": "fetalsyngen_b200.generator.augmentation.synthseg.",
    "fetalsyngen.generator.augmentation.artifacts.": "fetalsyngen_b200.generator.augmentation.artifacts.",
    "fetalsyngen.generator.artifacts.utils.": "fetalsyngen_b200.generator.artifacts.utils.",
    "fetalsyngen.data.datasets.": "fetalsyngen_b200.data.datasets.",
}


def resolve_target(target: str, aliases: bool = True):
    if aliases:
        for ref, ours in ALIAS_PREFIXES.items():
            if target.startswith(ref):
                target = ours + target[len(ref):]
                break
    mod, _, name = target.rpartition(".")
    try:
        return getattr(importlib.import_module(mod), name)
    except (ImportError, AttributeError) as e:
        raise ImportError(f"cannot resolve _target_ {target!r}: {e}") from e


def _interp(value, root, path):
    """``${..key}``: one dot = the current node, every further dot one level up."""
    if isinstance(value, str) and value.startswith("${") and value.endswith("}"):
        ref = value[2:-1]
        key = ref.lstrip(".")
        up = len(ref) - len(key)
        node = root
        for k in (path[: len(path) - (up - 1)] if up > 0 else []):
            node = node[k]
        for part in key.split("."):
            node = node[part]
        return _interp(node, root, path)
    return value


def instantiate(cfg, aliases: bool = True, _root=None, _path=None, **overrides):
    """Recursive ``_target_`` construction (the subset of hydra.utils.instantiate the reference uses)."""
    root = cfg if _root is None else _root
    path = [] if _path is None else _path
    if isinstance(cfg, dict):
        built = {}
        for k, v in cfg.items():
            if k in ("_target_", "defaults"):
                continue
            v = _interp(v, root, path)
            built[k] = instantiate(v, aliases, root, path + [k]) if isinstance(v, (dict, list)) else v
        built.update(overrides)
        if "_target_" in cfg:
            return resolve_target(cfg["_target_"], aliases)(**built)
        return built
    if isinstance(cfg, list):
        out = []
        for i, v in enumerate(cfg):
            v = _interp(v, root, path)
            out.append(instantiate(v, aliases, root, path + [i]) if isinstance(v, (dict, list)) else v)
        return out
    return cfg


def load_yaml(path, config_root=None) -> dict:
    """Load a YAML file and merge its ``defaults:`` list the way the reference's configs use it:
    ``- generator/default`` mounts ``<dir>/generator/default.yaml`` under key ``generator``."""
    path = Path(path)
    cfg = yaml.safe_load(path.read_text()) or {}
    base = Path(config_root) if config_root is not None else path.parent
    for item in cfg.pop("defaults", None) or []:
        if isinstance(item, str) and item != "_self_":
            group, _, _name = item.rpartition("/")
            sub = load_yaml(base / f"{item}.yaml", base)
            if group:
                node = cfg
                for part in group.split("/")[:-1]:
                    node = node.setdefault(part, {})
                node.setdefault(group.split("/")[-1], {})
                merged = dict(sub)
                merged.update(node[group.split("/")[-1]] or {})
                node[group.split("/")[-1]] = merged
            else:
                merged = dict(sub)
                merged.update(cfg)
                cfg = merged
        elif isinstance(item, dict):
            for group, name in item.items():
                if name is None:
                    continue
                sub = load_yaml(base / group / f"{name}.yaml", base)
                merged = dict(sub)
                merged.update(cfg.get(group) or {})
                cfg[group] = merged
    return cfg
