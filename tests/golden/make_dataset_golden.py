"""Freeze what the UNMODIFIED reference Datasets do around the transform (rows R1 / F4 of SURVEY.md section 8).

Run in the build container only (needs /root/reference):
    python tests/golden/make_dataset_golden.py
Writes tests/golden/dataset_golden.json (+ dataset_sensor.npz).

The reference modules import decord / matplotlib / seaborn / imageio / IPython, which are not installed; they are
stubbed with empty modules (none of them is used by the code under test) and the two I/O classes the Datasets call are
replaced by fakes fed from a registry:
  decord.VideoReader(path, ctx)  -> len(), [0].shape, get_batch(indices) (frames whose first pixel encodes the index)
  cv2.VideoCapture(path)         -> .get(CAP_PROP_FRAME_COUNT / CAP_PROP_FPS), .release()
Everything else — `NvidiaDashcamDataset.__getitem__` (nexar_videos.py:348-496), `_load_and_sync_sensor_data`
(:302-346), `_find_video_and_sensor_paths` (:17-36), `VideoDataset.__getitem__` (nexar_complete_with_validation.py
:117-234) and the notebook's `_load_video_frames` index rule (inference.ipynb cell 0) — is the reference's own code,
executed as it is.  Recorded per case: the frame indices of the returned clip, the indices asked of the decoder (empty
when the reference fell into its swallow-everything fallback and returned the all-zeros [T,720,1280,3] clip, e.g. for a
NaN timestamp), the sensor rows, and the next `random.random()` after the call (pins how many draws the call consumed).
"""
import json
import os
import random
import sys
import tempfile
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = os.environ.get("NEXAR_REFERENCE_DIR", "/root/reference")

REGISTRY = {}          # path -> dict(n=frames, fps=float, h=int, w=int)
ASKED = []             # indices of the last get_batch call


class _FakeFrames:
    def __init__(self, arr):
        self._a = arr

    def asnumpy(self):
        return self._a

    def __len__(self):
        return len(self._a)


class FakeVideoReader:
    def __init__(self, path, ctx=None):
        if path not in REGISTRY:
            raise RuntimeError(f"cannot open {path}")
        self.meta = REGISTRY[path]

    def __len__(self):
        return self.meta["n"]

    def _frame(self, i):
        f = np.zeros((self.meta["h"], self.meta["w"], 3), np.uint8)
        f[0, 0, 0], f[0, 0, 1] = i & 255, i >> 8
        return f

    def __getitem__(self, i):
        return self._frame(i)

    def get_batch(self, indices):
        ASKED[:] = [int(i) for i in indices]
        if len(indices) == 0:
            return _FakeFrames(np.zeros((0, self.meta["h"], self.meta["w"], 3), np.uint8))
        return _FakeFrames(np.stack([self._frame(int(i)) for i in indices]))


class FakeCapture:
    def __init__(self, path):
        self.meta = REGISTRY.get(path)

    def get(self, prop):
        import cv2
        if self.meta is None:
            return 0.0
        if prop == cv2.CAP_PROP_FRAME_COUNT:
            return float(self.meta["n"])
        if prop == cv2.CAP_PROP_FPS:
            return float(self.meta["fps"])
        return 0.0

    def release(self):
        pass


def import_reference_datasets():
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    for name in ("pytorchvideo", "matplotlib", "seaborn", "imageio", "IPython"):
        if name not in sys.modules:
            stub(name)
    stub("pytorchvideo.transforms", create_video_transform=None)
    mpl = sys.modules["matplotlib"]
    mpl.use = lambda *a, **k: None
    for sub in ("pyplot", "gridspec", "patches", "animation", "colors", "cm", "ticker", "dates"):
        setattr(mpl, sub, stub(f"matplotlib.{sub}", GridSpec=object, FuncAnimation=object, LinearSegmentedColormap=object))
    stub("IPython.display", display=lambda *a, **k: None, HTML=object, clear_output=lambda *a, **k: None, Video=object,
         Javascript=object)
    stub("decord", VideoReader=FakeVideoReader, cpu=lambda i=0: None, bridge=types.SimpleNamespace(set_bridge=lambda *_: None))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import cv2
    import nexar_videos
    nexar_videos.cv2.VideoCapture = FakeCapture            # the module looks `cv2.VideoCapture` up at call time
    import torch  # noqa: F401  (nexar_complete_with_validation uses `torch` before importing it at :51 unless present)
    import builtins
    builtins.torch = torch                                  # ncwv:51 references `torch` without importing it first
    try:
        import nexar_complete_with_validation as ncwv
    finally:
        del builtins.torch
    assert cv2.VideoCapture is FakeCapture
    return nexar_videos, ncwv


def decode_indices(frames):
    """frames: float [T,H,W,C] in [0,1] (transform=None path divides by 255) -> the source frame index of each."""
    a = (frames[:, 0, 0, :2].numpy() * 255.0).round().astype(np.int64)
    return [int(lo + 256 * hi) for lo, hi in a]


def main():
    nv, ncwv = import_reference_datasets()
    tmp = tempfile.mkdtemp(prefix="nexar_golden_")
    cases = []
    sensor_arrays = {}

    from dataset_fixture import build_tree, video_grid
    grid = video_grid()
    for path, g in build_tree(tmp, grid).items():
        REGISTRY[path] = dict(n=g["n"], fps=g["vfps"], h=4, w=6)
    meta = pd.DataFrame([{k: g[k] for k in ("id", "video_type", "event_time")} for g in grid])
    for strategy, time_column, fps_d in (("random", None, (10, 5)), ("center", None, (10, 5)), ("metadata_time", "event_time", (10, 5)),
                                         ("uniform", None, (10, 5)), ("random", None, (8, 2)), ("center", None, (16, 1)),
                                         ("metadata_time", "missing_column", (10, 5))):
        ds = nv.NvidiaDashcamDataset(meta, [tmp], fps=fps_d[0], duration=fps_d[1], is_train=True, skip_missing=True,
                                     transform=None, sample_strategy=strategy, time_column=time_column)
        assert len(ds) == len(grid)
        for seed in (0, 1, 2):
            for i, g in enumerate(grid):
                random.seed(1000 * seed + i)
                ASKED[:] = []
                item = ds[i]
                after = random.random()
                idx = decode_indices(item["frames"])
                key = f"nv/{strategy}/{time_column}/{fps_d[0]}x{fps_d[1]}/{seed}/{g['id']}"
                sensor_arrays[key] = item["sensor"].numpy()
                cases.append(dict(key=key, dataset="NvidiaDashcamDataset", strategy=strategy, effective_strategy=ds.sample_strategy,
                                  time_column=time_column, fps=fps_d[0], duration=fps_d[1], seed=1000 * seed + i, id=g["id"],
                                  n=g["n"], video_fps=g["vfps"], event_time=None if g["event_time"] != g["event_time"] else g["event_time"],
                                  sensor_kind=g["sensor_kind"], fname=g["fname"], target=item["target"],
                                  indices=idx, asked=list(ASKED), next_random=after.hex(),
                                  frames_shape=list(item["frames"].shape)))
    # discovery rules (_find_video_and_sensor_paths, skip_missing) and the zero-clip fallback
    meta2 = pd.DataFrame([dict(id="v001", video_type="Normal"), dict(id="nosuch", video_type="Collision"),
                          dict(id="v002", video_type="Normal")])
    ds_skip = nv.NvidiaDashcamDataset(meta2, [tmp], skip_missing=True, transform=None)
    ds_keep = nv.NvidiaDashcamDataset(meta2, [tmp], skip_missing=False, transform=None)
    broken = ds_keep[1]
    discovery = dict(skip_len=len(ds_skip), keep_len=len(ds_keep),
                     skip_paths=[os.path.relpath(p, tmp) for p in ds_skip.video_paths],
                     keep_paths=[os.path.relpath(p, tmp) for p in ds_keep.video_paths],
                     broken_frames_shape=list(broken["frames"].shape), broken_frames_sum=float(broken["frames"].sum()),
                     broken_sensor_shape=list(broken["sensor"].shape), broken_target=broken["target"])

    # ---- VideoDataset (nexar_complete_with_validation.py) --------------------------------------------------------
    paths = [os.path.join(tmp, g["id"], g["fname"]) for g in grid]
    labels = [g["video_type"] for g in grid]
    ids = [g["id"] for g in grid]
    meta3 = pd.DataFrame([dict(id=g["id"], center=g["event_time"]) for g in grid if int(g["id"][1:]) % 7])  # some ids missing
    for strategy in ("random", "center", "metadata_center"):
        ds = ncwv.VideoDataset(paths, labels, ids, fps=10, duration=5, transform=None, sample_strategy=strategy,
                               center_time_column="center" if strategy == "metadata_center" else None,
                               metadata_df=meta3 if strategy == "metadata_center" else None)
        for seed in (0, 1):
            for i, g in enumerate(grid):
                random.seed(7000 * seed + i)
                ASKED[:] = []
                item = ds[i]
                after = random.random()
                in_meta = bool(int(g["id"][1:]) % 7)
                cases.append(dict(key=f"vd/{strategy}/{seed}/{g['id']}", dataset="VideoDataset", strategy=strategy, fps=10, duration=5,
                                  seed=7000 * seed + i, id=g["id"], n=g["n"], video_fps=g["vfps"],
                                  center=(None if (not in_meta or g["event_time"] != g["event_time"]) else g["event_time"]),
                                  target=item["target"], indices=decode_indices(item["frames"]), asked=list(ASKED),
                                  next_random=after.hex(), frames_shape=list(item["frames"].shape)))

    # ---- notebook uniform sampler (inference.ipynb cell 0: the only linspace rule in the reference) -----------------
    nb = json.load(open(os.path.join(REF, "inference.ipynb")))
    src = "".join(nb["cells"][0]["source"])
    assert "np.linspace(0, total_frames - 1, self.num_frames, dtype=int)" in src and "mode='wrap'" in src
    uniform = {}
    for total in (1, 5, 15, 16, 17, 31, 100, 299, 1200):
        for num in (16, 50):
            if total < num:     # the notebook's two lines, evaluated by numpy itself
                idx = np.pad(np.arange(total), (0, num - total), mode='wrap')
            else:
                idx = np.linspace(0, total - 1, num, dtype=int)
            uniform[f"{total}/{num}"] = [int(i) for i in idx]

    out = dict(versions=dict(numpy=np.__version__, pandas=pd.__version__), cases=cases, discovery=discovery, uniform=uniform)
    with open(os.path.join(HERE, "dataset_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    np.savez_compressed(os.path.join(HERE, "dataset_sensor.npz"), **sensor_arrays)
    print(f"{len(cases)} cases, {len(sensor_arrays)} sensor tables -> tests/golden/dataset_golden.json, dataset_sensor.npz")


if __name__ == "__main__":
    main()
