"""Integer restatement of the reference's temporal window selection (R1).
TEST INFRASTRUCTURE ONLY.

Follows the reference Datasets line by line over a frame count.  Integer-only, hence
exact.  Parity pinning: tests/test_dataset_golden.py holds it to what the UNMODIFIED
``NvidiaDashcamDataset`` / ``VideoDataset`` did on a 36-video grid (972 cases, indices and
``random`` consumption), frozen by tests/golden/make_dataset_golden.py, which imports the
reference modules with stubs for their missing plotting / decode imports and fake
``decord.VideoReader`` / ``cv2.VideoCapture`` objects.
"""
import random as _random
from typing import List, Optional

import numpy as np


def start_frame(num_frames: int, need: int, strategy: str, rng=_random,
                timestamp_sec: Optional[float] = None, video_fps: float = 0.0) -> int:
    """nexar_videos.py:364-415 / nexar_complete_with_validation.py:126-155."""
    if strategy in ("metadata_time", "metadata_center"):
        if timestamp_sec is not None and video_fps > 0:
            half = need // 2
            center = int(timestamp_sec * video_fps)
            start = max(0, center - half)
            if start + need > num_frames:
                start = max(0, num_frames - need)
            if strategy == "metadata_time":
                start = max(0, min(start, num_frames - 1))  # nexar_videos.py:387
        elif strategy == "metadata_time":
            start = rng.randint(0, max(0, num_frames - need))  # nexar_videos.py:389,391
        else:
            start = rng.randint(0, num_frames - need) if num_frames > need else 0  # ncwv:198-202
    elif strategy == "center":
        if num_frames > need:
            start = max(0, num_frames // 2 - need // 2)
            if start + need > num_frames:
                start = max(0, num_frames - need)
        else:
            start = 0
    else:  # 'random' (and anything unknown, nexar_videos.py:57-58)
        start = rng.randint(0, num_frames - need) if num_frames > need else 0
    return max(0, min(start, num_frames - 1))  # nexar_videos.py:415


def window_indices(num_frames: int, need: int, start: int) -> List[int]:
    """nexar_videos.py:416-435: range(start, min(start+need, N)), short windows
    padded by repeating the last frame, long ones truncated."""
    end = min(start + need, num_frames)
    idx = list(range(start, end))
    if len(idx) < need and idx:
        idx = idx + [idx[-1]] * (need - len(idx))
    return idx[:need]


def uniform_indices(total_frames: int, num_frames: int) -> List[int]:
    """inference.ipynb cell 0 ``_load_video_frames``: wrap-pad if short, else
    linspace(0, N-1, num, dtype=int) (truncation)."""
    if total_frames < num_frames:
        idx = np.pad(np.arange(total_frames), (0, num_frames - total_frames), mode="wrap")
    else:
        idx = np.linspace(0, total_frames - 1, num_frames, dtype=int)
    return [int(i) for i in idx]


def model_subsample(num_frames: int) -> List[int]:
    """nexar_arch.py:411-415 — the model keeps even frames when T > 10."""
    return list(range(0, num_frames, 2)) if num_frames > 10 else list(range(num_frames))


def sliding_window_starts(num_frames: int, window: int, stride: int) -> List[int]:
    """cfg4 extension: starts k*stride, k = 0..floor((N-window)/stride);
    the reference itself has no sliding-window code."""
    if num_frames <= window:
        return [0]
    return [k * stride for k in range((num_frames - window) // stride + 1)]
