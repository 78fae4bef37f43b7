"""Rows R1 / F4 pinned to the REAL reference: tests/golden/dataset_golden.json + dataset_sensor.npz were produced by
running the unmodified ``NvidiaDashcamDataset`` / ``VideoDataset`` of /root/reference over a synthetic video grid
(tests/golden/make_dataset_golden.py, fake decoder + fake cv2.VideoCapture).  Here the oracle restatement and the
product's Dataset mirrors run over the same grid and must reproduce, case by case: the frame indices of the clip, the
indices asked of the decoder, the amount of ``random`` consumed, the sensor rows, file discovery and the fallbacks."""
import json
import os
import random
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

from dataset_fixture import build_tree, video_grid  # noqa: E402
from oracle import clip_sampler_oracle as S  # noqa: E402

with open(os.path.join(HERE, "golden", "dataset_golden.json")) as _f:
    GOLD = json.load(_f)
SENSOR = np.load(os.path.join(HERE, "golden", "dataset_sensor.npz"))
GRID = {g["id"]: g for g in video_grid()}


def _nv_cases(strategy=None):
    return [c for c in GOLD["cases"] if c["dataset"] == "NvidiaDashcamDataset" and (strategy is None or c["strategy"] == strategy)]


def _vd_cases():
    return [c for c in GOLD["cases"] if c["dataset"] == "VideoDataset"]


# ---------------------------------------------------------------------------------------------------------------------
# the oracle restatement against the reference's own behaviour
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_window(c, timestamp):
    need = c["fps"] * c["duration"]
    strategy = c.get("effective_strategy") or c["strategy"]
    random.seed(c["seed"])
    start = S.start_frame(c["n"], need, strategy, random, timestamp, c["video_fps"])
    return S.window_indices(c["n"], need, start), random.random().hex()


def test_oracle_sampler_matches_the_reference_datasets():
    n_checked = 0
    for c in _nv_cases():
        if not c["asked"]:                    # the reference fell into its zero-clip fallback (NaN timestamp)
            continue
        ts = c["event_time"] if c["effective_strategy"] == "metadata_time" else None
        idx, nxt = _oracle_window(c, ts)
        assert idx == c["indices"] and nxt == c["next_random"], c["key"]
        n_checked += 1
    for c in _vd_cases():
        idx, nxt = _oracle_window(c, c["center"] if c["strategy"] == "metadata_center" else None)
        assert idx == c["indices"] and nxt == c["next_random"], c["key"]
        n_checked += 1
    assert n_checked > 900


def test_uniform_sampler_matches_the_notebook_rule():
    from vision_collision_detection_b200 import videos as V
    for key, want in GOLD["uniform"].items():
        total, num = (int(v) for v in key.split("/"))
        assert S.uniform_indices(total, num) == want and V.uniform_indices(total, num) == want, key


# ---------------------------------------------------------------------------------------------------------------------
# the product's Dataset mirrors on the same grid
# ---------------------------------------------------------------------------------------------------------------------
class _Reader:
    """Index-encoding fake decoder (the same convention the golden generator used)."""
    asked = []

    def __init__(self, meta):
        self.n, self.fps = meta["n"], meta["vfps"]

    def __len__(self):
        return self.n

    def get_avg_fps(self):
        return self.fps

    def get_batch(self, indices):
        _Reader.asked = [int(i) for i in indices]
        out = np.zeros((len(indices), 4, 6, 3), np.uint8)
        for k, i in enumerate(indices):
            out[k, 0, 0, 0], out[k, 0, 0, 1] = int(i) & 255, int(i) >> 8
        return out


def _decode(frames):
    a = (frames[:, 0, 0, :2].numpy() * 255.0).round().astype(np.int64)
    return [int(lo + 256 * hi) for lo, hi in a]


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    base = str(tmp_path_factory.mktemp("nexar_grid"))
    registry = build_tree(base)

    def decoder(path):
        if path not in registry:
            raise RuntimeError(f"cannot open {path}")
        return _Reader(registry[path])

    return base, registry, decoder


@pytest.mark.parametrize("strategy,time_column,fps,duration", [
    ("random", None, 10, 5), ("center", None, 10, 5), ("metadata_time", "event_time", 10, 5), ("uniform", None, 10, 5),
    ("random", None, 8, 2), ("center", None, 16, 1), ("metadata_time", "missing_column", 10, 5)])
def test_dashcam_dataset_matches_the_reference(tree, strategy, time_column, fps, duration):
    from vision_collision_detection_b200 import videos as V
    base, registry, decoder = tree
    rows = [dict(id=g["id"], video_type=g["video_type"], event_time=g["event_time"]) for g in GRID.values()]
    ds = V.GpuDashcamDataset(rows, [base], fps=fps, duration=duration, transform=None, sample_strategy=strategy,
                             time_column=time_column, decoder=decoder, video_fps_lookup=lambda p: registry[p]["vfps"])
    cases = {c["key"]: c for c in _nv_cases(strategy) if c["time_column"] == time_column and c["fps"] == fps and c["duration"] == duration}
    assert len(ds) == len(GRID) and len(cases) == 3 * len(GRID)
    assert ds.sample_strategy == next(iter(cases.values()))["effective_strategy"]
    ids = [r["id"] for r in ds.rows]
    for key, c in cases.items():
        i = ids.index(c["id"])
        random.seed(c["seed"])
        _Reader.asked = []
        item = ds[i]
        nxt = random.random().hex()
        assert list(item["frames"].shape) == c["frames_shape"], key
        assert item["target"] == c["target"] and item["id"] == c["id"]
        if c["asked"]:
            assert _decode(item["frames"]) == c["indices"] and _Reader.asked == c["asked"], key
        else:                                       # reference fallback: all-zeros clip of the standard size
            assert float(item["frames"].abs().sum()) == 0.0, key
        assert nxt == c["next_random"], key
        want = SENSOR[key]
        got = item["sensor"].numpy()
        assert got.shape == want.shape and np.array_equal(got, want, equal_nan=True), key


def test_video_discovery_and_missing_files_match_the_reference(tree):
    from vision_collision_detection_b200 import videos as V
    base, _, decoder = tree
    rows = [dict(id="v001", video_type="Normal"), dict(id="nosuch", video_type="Collision"), dict(id="v002", video_type="Normal")]
    d = GOLD["discovery"]
    skip = V.GpuDashcamDataset(rows, [base], skip_missing=True, transform=None, decoder=decoder)
    keep = V.GpuDashcamDataset(rows, [base], skip_missing=False, transform=None, decoder=decoder)
    assert len(skip) == d["skip_len"] and len(keep) == d["keep_len"]
    assert [os.path.relpath(p, base) for p in skip.video_paths] == d["skip_paths"]     # anonymized_*.mp4 and *.mov are found
    assert [os.path.relpath(p, base) for p in keep.video_paths] == d["keep_paths"]
    broken = keep[1]
    assert list(broken["frames"].shape) == d["broken_frames_shape"] and float(broken["frames"].sum()) == d["broken_frames_sum"]
    assert list(broken["sensor"].shape) == d["broken_sensor_shape"] and broken["target"] == d["broken_target"]


@pytest.mark.parametrize("strategy", ["random", "center", "metadata_center"])
def test_video_dataset_matches_the_reference(tree, strategy):
    from vision_collision_detection_b200 import videos as V
    base, registry, decoder = tree
    grid = list(GRID.values())
    paths = [os.path.join(base, g["id"], g["fname"]) for g in grid]
    meta = [dict(id=g["id"], center=g["event_time"]) for g in grid if int(g["id"][1:]) % 7]     # some ids are not in the metadata
    ds = V.GpuVideoDataset(paths, [g["video_type"] for g in grid], [g["id"] for g in grid], fps=10, duration=5, transform=None,
                           sample_strategy=strategy, center_time_column="center" if strategy == "metadata_center" else None,
                           metadata_df=meta if strategy == "metadata_center" else None, decoder=decoder,
                           video_fps_lookup=lambda p: registry[p]["vfps"])
    cases = [c for c in _vd_cases() if c["strategy"] == strategy]
    assert len(cases) == 2 * len(grid)
    for c in cases:
        i = [g["id"] for g in grid].index(c["id"])
        random.seed(c["seed"])
        _Reader.asked = []
        item = ds[i]
        nxt = random.random().hex()
        assert list(item["frames"].shape) == c["frames_shape"], c["key"]
        assert _decode(item["frames"]) == c["indices"] and _Reader.asked == c["asked"], c["key"]
        assert nxt == c["next_random"] and item["target"] == c["target"], c["key"]


def test_transform_errors_are_not_swallowed_and_workers_are_refused():
    """a failure of the GPU transform must not be turned into a black clip."""
    from vision_collision_detection_b200 import videos as V

    class Boom(torch.nn.Module):
        crop_size = 8

        def forward(self, x):
            raise RuntimeError("CUDA error: simulated")

    rows = [dict(id="a", video_type="Normal", path="a.mp4")]
    ds = V.GpuDashcamDataset(rows, fps=2, duration=2, transform=Boom(), sample_strategy="center",
                             decoder=lambda p: _Reader(dict(n=10, vfps=10.0)))
    with pytest.raises(RuntimeError, match="simulated"):
        ds[0]
    vd = V.GpuVideoDataset(["a.mp4"], [0], fps=2, duration=2, transform=Boom(), sample_strategy="center",
                           decoder=lambda p: _Reader(dict(n=10, vfps=10.0)))
    with pytest.raises(RuntimeError, match="simulated"):
        vd[0]
    # decode failures still are (the reference's contract)
    bad = V.GpuDashcamDataset(rows, fps=2, duration=2, transform=Boom(), sample_strategy="center",
                              decoder=lambda p: (_ for _ in ()).throw(IOError("no such file")))
    assert tuple(bad[0]["frames"].shape) == (4, 224, 224, 3) and float(bad[0]["frames"].sum()) == 0.0
    # defer=False inside a DataLoader worker process is refused with a clear message
    from torch.utils.data import DataLoader
    with pytest.raises(RuntimeError, match="DataLoader worker"):
        next(iter(DataLoader(ds, batch_size=1, num_workers=1)))
