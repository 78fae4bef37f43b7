"""IMU / accelerometer rows aligned to video frames.

The reference does this per item with pandas (``nexar_videos.py:302-346``): read the accelerometer CSV, make the
timestamps relative to the first sample, re-index onto the union of sample times and frame times, interpolate
linearly along the index and pick the frame times; ``__getitem__`` then cuts the clip's window out of it and pads
with the last row (``nexar_videos.py:453-477``).  pandas' ``interpolate('index')`` is ``numpy.interp`` on the valid
samples with leading NaNs preserved, so the same numbers come out of the few vectorised numpy calls below (float64,
bit for bit; checked against the pandas expression in tests/test_host_logic.py) without building three DataFrames
per clip."""
from __future__ import annotations

import csv
import os
from typing import Optional, Sequence

import numpy as np

SENSOR_FILE = "Dashcam-Accelerometer_Acceleration.csv"          # nexar_videos.py:33
SENSOR_COLUMNS = ("accel_x_G", "accel_y_G", "accel_z_G", "accel_total_G")   # nexar_videos.py:340


def find_sensor_path(video_path: Optional[str], sensor_subdir: str = "signals") -> Optional[str]:
    """nexar_videos.py:33-34: the CSV lives in ``<video dir>/<sensor_subdir>/``; None when it is missing."""
    if not video_path:
        return None
    p = os.path.join(os.path.dirname(video_path), sensor_subdir, SENSOR_FILE)
    return p if os.path.exists(p) else None


def read_sensor_csv(path: str):
    """-> (time_sec [N] float64, accel [N,4] float64).  Same columns ``pd.read_csv(path, index_col=0)`` would
    expose by name; empty cells become NaN as in pandas."""
    try:   # the reference's own call (nexar_videos.py:313): pandas' default float parser is NOT round-trip exact (it can be
        # one ulp off Python's float()), so bit-equality with the reference needs the same parser
        import pandas as pd
        df = pd.read_csv(path, index_col=0)
        return (df["time_sec"].to_numpy(dtype=np.float64),
                df[list(SENSOR_COLUMNS)].to_numpy(dtype=np.float64))
    except ImportError:
        pass
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    if len(rows) < 2:
        raise ValueError("empty sensor file")
    header = [h.strip() for h in rows[0]]
    cols = [header.index(c) for c in ("time_sec",) + SENSOR_COLUMNS]

    def num(s):
        s = s.strip()
        return float(s) if s else float("nan")

    data = np.array([[num(r[c]) for c in cols] for r in rows[1:] if r], dtype=np.float64)
    return data[:, 0], data[:, 1:]


def sync_sensor_to_frames(time_sec: Sequence[float], accel, frame_count: int, fps: float) -> np.ndarray:
    """nexar_videos.py:318-341 -> float64 ``[frame_count, 4]``: the accelerometer interpolated at ``i / fps``.
    Raises ValueError where pandas raises (duplicate timestamps); the caller maps every failure to zeros."""
    time_sec = np.asarray(time_sec, dtype=np.float64)
    accel = np.asarray(accel, dtype=np.float64)
    if time_sec.ndim != 1 or accel.shape != (time_sec.shape[0], 4) or time_sec.shape[0] == 0:
        raise ValueError("expected time_sec [N] and accel [N,4]")
    rel = time_sec - time_sec[0]                                     # :324-326
    if np.isnan(rel).any():
        raise ValueError("NaN timestamps")                           # union/reindex would not place them
    order = np.argsort(rel, kind="stable")                           # reindex onto the (sorted) union
    rel, accel = rel[order], accel[order]
    if (np.diff(rel) == 0).any():
        raise ValueError("cannot reindex on an axis with duplicate labels")   # what pandas raises at :332
    vt = np.array([i / fps for i in range(frame_count)], dtype=np.float64)    # :329 (python float division)
    out = np.empty((frame_count, 4), dtype=np.float64)
    for c in range(4):
        y = accel[:, c]
        ok = ~np.isnan(y)
        if not ok.any():
            out[:, c] = np.nan
            continue
        col = np.interp(vt, rel[ok], y[ok])                          # 'index' interpolation; right side clamps
        col[vt < rel[ok][0]] = np.nan                                # leading NaNs stay NaN (limit_direction='forward')
        out[:, c] = col
    return out


def load_and_sync_sensor(sensor_path: Optional[str], frame_count: int, fps: float, need: int) -> np.ndarray:
    """``_load_and_sync_sensor_data`` (nexar_videos.py:301-346): zeros ``[need, 4]`` float32 on any problem."""
    empty = np.zeros((need, 4), dtype=np.float32)
    if sensor_path is None or not os.path.exists(sensor_path):
        return empty
    try:
        if frame_count == 0 or fps == 0:
            return empty
        t, a = read_sensor_csv(sensor_path)
        return sync_sensor_to_frames(t, a, frame_count, fps)
    except Exception:
        return empty


def window_sensor(sensor: np.ndarray, num_frames: int, start: int, end: int, need: int) -> np.ndarray:
    """nexar_videos.py:456-474: the clip's rows, padded with the last row / trimmed to ``need``; zeros when the
    sensor table is shorter than the video.  Returns float32 ``[need, 4]``."""
    if isinstance(sensor, np.ndarray) and len(sensor) > 0 and len(sensor) >= num_frames:
        cut = sensor[start:end]
        if len(cut) < need:
            last = cut[-1] if len(cut) > 0 else np.zeros(4, dtype=np.float32)
            cut = np.concatenate([cut, np.repeat(last[np.newaxis, :], need - len(cut), axis=0)], axis=0)
        elif len(cut) > need:
            cut = cut[:need]
        return np.asarray(cut, dtype=np.float32)
    return np.zeros((need, 4), dtype=np.float32)
