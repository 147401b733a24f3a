"""Minimal loader for the reference's mmcv-style backbone config files (``visbackbone/swin_*.py``).

Only ``model.backbone.{patch_size, embed_dim, depths, num_heads, window_size, patch_norm}`` is read by
``get_vidswin_model`` (reference video_swin.py:624-637), so this implements exactly what that needs:
python-file configs, ``_base_`` inheritance (string or list, relative to the including file) and
recursive dict merging.  The reference's own engine (visbackbone/config.py, a vendored mmcv Config)
needs ``addict``/``yapf`` and is out of scope.

When the file is not on disk (the reference's ``violet`` branch points at a directory that does not
exist, video_swin.py:598) the built-in preset of the same base name is used.
"""
from __future__ import annotations

import os
from typing import Any, Dict

_TINY = dict(patch_size=(4, 4, 4), embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
             window_size=(8, 7, 7), mlp_ratio=4., qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
             drop_path_rate=0.2, patch_norm=True)

# values of visbackbone/swin_{tiny,small,base,large,violet}.py merged with the *_patch244_* overrides
PRESETS: Dict[str, Dict[str, Any]] = {
    "swin_tiny_patch244_window877_kinetics400_1k": {**_TINY, "patch_size": (2, 4, 4)},
    "swin_small_patch244_window877_kinetics400_1k": {**_TINY, "patch_size": (2, 4, 4), "depths": [2, 2, 18, 2]},
    "swin_base_patch244_window877_kinetics400_1k": {**_TINY, "patch_size": (2, 4, 4), "depths": [2, 2, 18, 2],
                                                    "embed_dim": 128, "num_heads": [4, 8, 16, 32]},
    "swin_base_patch244_window877_kinetics400_22k": {**_TINY, "patch_size": (2, 4, 4), "depths": [2, 2, 18, 2],
                                                     "embed_dim": 128, "num_heads": [4, 8, 16, 32]},
    "swin_base_patch244_window877_kinetics600_22k": {**_TINY, "patch_size": (2, 4, 4), "depths": [2, 2, 18, 2],
                                                     "embed_dim": 128, "num_heads": [4, 8, 16, 32]},
    "swin_large_patch244_window877_kinetics400_22k": {**_TINY, "patch_size": (2, 4, 4), "depths": [2, 2, 18, 2],
                                                      "embed_dim": 192, "num_heads": [6, 12, 24, 48]},
    "swin_large_384_patch244_window81212_kinetics400_22k": {**_TINY, "patch_size": (2, 4, 4),
                                                            "depths": [2, 2, 18, 2], "embed_dim": 192,
                                                            "num_heads": [6, 12, 24, 48], "window_size": (8, 12, 12)},
    "swin_large_384_patch244_window81212_kinetics600_22k": {**_TINY, "patch_size": (2, 4, 4),
                                                            "depths": [2, 2, 18, 2], "embed_dim": 192,
                                                            "num_heads": [6, 12, 24, 48], "window_size": (8, 12, 12)},
    "swin_violet_patch244_window877": {**_TINY, "patch_size": (2, 4, 4), "depths": [2, 2, 18, 2]},
}


def _merge(base: dict, over: dict) -> dict:
    out = dict(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.pop("_delete_", False):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def _load_file(path: str) -> dict:
    scope: Dict[str, Any] = {}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), scope)  # config files are plain python assignments
    cfg = {k: v for k, v in scope.items() if not k.startswith("__") and not callable(v) and k != "_base_"}
    bases = scope.get("_base_", [])
    if isinstance(bases, str):
        bases = [bases]
    merged: dict = {}
    for b in bases:
        merged = _merge(merged, _load_file(os.path.join(os.path.dirname(path), b)))
    return _merge(merged, cfg)


def load_backbone_cfg(config_path: str) -> Dict[str, Any]:
    """-> the ``model.backbone`` dict of a swin_* config (file on disk, else built-in preset)."""
    candidates = [config_path, os.path.join("visbackbone", os.path.basename(config_path))]
    for p in candidates:
        if os.path.isfile(p):
            try:
                return dict(_load_file(p)["model"]["backbone"])
            except (OSError, KeyError):
                break  # e.g. swin_small_* inherits from a path that does not exist upstream
    name = os.path.splitext(os.path.basename(config_path))[0]
    if name in PRESETS:
        return dict(PRESETS[name])
    raise FileNotFoundError(f"backbone config {config_path!r} not found and no built-in preset named {name!r}")
