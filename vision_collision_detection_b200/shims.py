"""Import shims for the reference's own scripts.

``nexar_inference.InferenceEngine.predict`` builds its Dataset with
``from nexar_video_aug import create_video_transforms`` and ``from nexar_data import NvidiaDashcamDataset``
(nexar_inference.py:203-204).  ``nexar_data`` does not exist in the reference tree (the class lives in
``nexar_videos.py``), so that call path fails there with the "Required package import error" message.
``install()`` registers a ``nexar_data`` module whose ``NvidiaDashcamDataset`` is the GPU Dataset mirror, and — on
request — replaces ``nexar_video_aug``'s factories with the GPU ones, so the unmodified ``predict`` (and the trainers'
``from nexar_video_aug import ...`` lines) run on this path.  Nothing is installed implicitly.
"""
from __future__ import annotations

import sys
import types
from typing import Dict


def install(replace_video_aug: bool = False, deferred: bool = False) -> Dict[str, types.ModuleType]:
    """Register the shim modules in ``sys.modules``; returns them by name.  Idempotent.  ``deferred=True`` makes the
    shimmed ``nexar_video_aug.create_video_transforms`` build deferred transforms (deferred.py), which is what the
    unmodified trainers need: their Datasets run in forked DataLoader workers."""
    from . import VideoAugmentation, create_video_transform, create_video_transforms
    from .videos import GpuDashcamDataset, GpuVideoDataset

    out: Dict[str, types.ModuleType] = {}
    data = sys.modules.get("nexar_data")
    if data is None or not getattr(data, "__nexar_b200_shim__", False):
        data = types.ModuleType("nexar_data")
        data.__doc__ = "shim: nexar_inference.py:204 imports NvidiaDashcamDataset from here"
        data.__nexar_b200_shim__ = True
        sys.modules["nexar_data"] = data
    data.NvidiaDashcamDataset = GpuDashcamDataset          # nexar_videos.py:28
    data.VideoDataset = GpuVideoDataset                    # nexar_complete_with_validation.py:57
    out["nexar_data"] = data
    if replace_video_aug:
        aug = types.ModuleType("nexar_video_aug")
        aug.__doc__ = "shim: the GPU factories under the reference's module name (nexar_video_aug.py:318,636)"
        aug.__nexar_b200_shim__ = True
        if deferred:
            import functools
            aug.create_video_transforms = functools.partial(create_video_transforms, deferred=True)
            functools.update_wrapper(aug.create_video_transforms, create_video_transforms)
        else:
            aug.create_video_transforms = create_video_transforms
        aug.create_video_transform = create_video_transform
        aug.VideoAugmentation = VideoAugmentation
        sys.modules["nexar_video_aug"] = aug
        out["nexar_video_aug"] = aug
    return out


def uninstall() -> None:
    for name in ("nexar_data", "nexar_video_aug"):
        m = sys.modules.get(name)
        if m is not None and getattr(m, "__nexar_b200_shim__", False):
            del sys.modules[name]
