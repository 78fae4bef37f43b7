"""B200-native clip transform for the dashcam collision classifier's input path.

Public surface (mirrors the reference's names for this path):
  create_video_transforms, create_video_transform, VideoAugmentation  (nexar_video_aug.py)
  GpuVideoTransform.forward / forward_batch
  clip sampler + Dataset/loader mirror in ``videos`` (nexar_videos.py), windowing in ``inference``.
The CUDA library is loaded lazily on first use and there is no CPU fallback.
"""
from .params import VideoAugmentation  # noqa: F401
from .video_aug import GpuVideoTransform, create_video_transform, create_video_transforms  # noqa: F401

__all__ = ["create_video_transforms", "create_video_transform", "VideoAugmentation", "GpuVideoTransform"]
