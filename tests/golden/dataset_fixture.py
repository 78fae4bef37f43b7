"""The synthetic video / sensor grid behind tests/golden/dataset_golden.json: shared by the generator
(make_dataset_golden.py, which runs the unmodified reference on it) and by tests/test_dataset_golden.py (which runs the
product on the same grid).  Deterministic: numpy RandomState streams are stable across numpy versions."""
import os

import numpy as np

SENSOR_FILE = "Dashcam-Accelerometer_Acceleration.csv"
FRAME_COUNTS = (1, 7, 49, 50, 51, 99, 100, 101, 150, 299, 900, 1200)
VIDEO_FPS = (10.0, 29.97, 30.0)


def sensor_rows(kind, n_rows, seed, t0=1000.0, rate=50.0):
    """-> dict of columns (time_sec + the four accelerometer channels)."""
    rs = np.random.RandomState(seed)
    t = t0 + np.arange(n_rows) / rate + (rs.uniform(-0.004, 0.004, n_rows) if kind != "regular" else 0.0)
    a = rs.normal(0, 1, (n_rows, 4))
    if kind == "nan":
        a[rs.uniform(size=a.shape) < 0.15] = np.nan
    if kind == "short":
        t = t[: max(2, n_rows // 4)]
        a = a[: len(t)]
    return {"time_sec": t, "accel_x_G": a[:, 0], "accel_y_G": a[:, 1], "accel_z_G": a[:, 2], "accel_total_G": a[:, 3]}


def video_grid():
    grid, vid_no = [], 0
    for n in FRAME_COUNTS:
        for fps in VIDEO_FPS:
            vid_no += 1
            vid = f"v{vid_no:03d}"
            grid.append(dict(
                id=vid, video_type=("Normal", "Near Collision", "Collision")[vid_no % 3], n=n, vfps=fps,
                sensor_kind=("regular", "jitter", "nan", "short", None)[vid_no % 5], sensor_seed=vid_no,
                sensor_rows=int(n / fps * 50) + 60,
                fname=(f"{vid}.mp4", f"anonymized_{vid}.mp4", f"{vid}.mov")[vid_no % 3],
                event_time=[0.0, 0.4, n / fps / 2, n / fps, n / fps + 3.0, float("nan")][vid_no % 6]))
    return grid


def build_tree(base, grid=None):
    """Create <base>/<id>/<fname> (empty files) and the sensor CSVs (pandas ``to_csv`` layout: an unnamed index column
    first, as the reference's ``pd.read_csv(path, index_col=0)`` expects).  -> {video path: grid row}."""
    import pandas as pd
    registry = {}
    for g in grid or video_grid():
        d = os.path.join(base, g["id"])
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, g["fname"])
        open(path, "wb").close()
        registry[path] = g
        if g["sensor_kind"]:
            os.makedirs(os.path.join(d, "signals"), exist_ok=True)
            pd.DataFrame(sensor_rows(g["sensor_kind"], g["sensor_rows"], g["sensor_seed"])).to_csv(
                os.path.join(d, "signals", SENSOR_FILE))
    return registry
