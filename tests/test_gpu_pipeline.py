"""GPU tests of the callers either side of the kernel: Dataset mirror + GpuAugLoader (drop-in consumer line),
inference windowing (cfg4 semantics), host-buffer pipeline (the e2e path of bench.py)."""
import random

import numpy as np
import pytest
import torch

from oracle import clip_sampler_oracle as S
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
TOL_AFTER = 1e-3


class _FakeReader:
    def __init__(self, n, seed, h=96, w=160):
        from vision_collision_detection_b200.synth import make_clip_np
        self.frames = make_clip_np(n, h, w, seed, "dashcam")

    def __len__(self):
        return len(self.frames)

    def get_batch(self, idx):
        return self.frames[list(idx)]


def _rows():
    return [{"id": f"v{i}", "video_type": "Normal", "path": f"{n}_{i}"} for i, n in enumerate([80, 30, 0, 64])]


def _decoder(path):
    n, seed = map(int, path.split("_"))
    if n == 0:
        raise IOError("broken video")
    return _FakeReader(n, seed)


def test_dataset_loader_drop_in_matches_oracle():
    from torch.utils.data import DataLoader
    from vision_collision_detection_b200 import create_video_transforms
    from vision_collision_detection_b200.videos import GpuAugLoader, GpuDashcamDataset, deferred_collate
    kw = dict(mode="train", crop_size=56, enable_custom_augmentation=True, brightness_range=(0.9, 1.1),
              contrast_range=(0.9, 1.1), saturation_range=(0.9, 1.1), rotation_range=(-5, 5))
    tf = create_video_transforms(**kw)
    ds = GpuDashcamDataset(_rows(), fps=4, duration=3, transform=tf, sample_strategy="random", decoder=_decoder, defer=True)
    random.seed(99)
    loader = GpuAugLoader(DataLoader(ds, batch_size=4, shuffle=False, num_workers=0, collate_fn=deferred_collate), tf)
    batch = next(iter(loader))
    frames = batch["frames"]
    assert tuple(frames.shape) == (4, 12, 56, 56, 3) and frames.is_cuda
    dev = frames.device
    x = batch["frames"].permute(0, 4, 1, 2, 3).float().to(dev)       # the trainers' consumer line (dvc:708)
    assert x.is_contiguous() and x.data_ptr() == frames.data_ptr()     # pure view chain, no copy
    assert torch.all(x[2] == 0)                                        # failed item -> all-zeros clip
    # same draws on the oracle
    cfg = O.TransformConfig(mode="train", crop_size=56, enable_custom_augmentation=True,
                            aug=O.AugConfig(rotation_range=(-5, 5)))
    random.seed(99)
    for i, row in enumerate(_rows()):
        n, seed = map(int, row["path"].split("_"))
        if n == 0:
            continue
        start = S.start_frame(n, 12, "random", random)
        idx = S.window_indices(n, 12, start)
        clip = _FakeReader(n, seed).frames[idx]
        want = O.clip_transform(clip.transpose(3, 0, 1, 2), cfg, random)
        assert np.abs(x[i].cpu().numpy() - want).max() <= TOL_AFTER, i


def test_dataset_with_transform_is_reference_shaped():
    from vision_collision_detection_b200 import create_video_transforms
    from vision_collision_detection_b200.videos import GpuDashcamDataset
    tf = create_video_transforms(mode="val", crop_size=56)
    ds = GpuDashcamDataset(_rows(), fps=4, duration=3, is_train=False, transform=tf, sample_strategy="center",
                           decoder=_decoder)
    item = ds[0]
    assert tuple(item["frames"].shape) == (12, 56, 56, 3) and item["frames"].dtype == torch.float32
    assert tuple(item["sensor"].shape) == (12, 4) and item["target"] == "Normal" and item["id"] == "v0"
    clip = _FakeReader(80, 0).frames[S.window_indices(80, 12, S.start_frame(80, 12, "center"))]
    want = O.apply_clip_transform(clip.transpose(3, 0, 1, 2), O.TransformConfig(mode="val", crop_size=56),
                                  {"flip": False, "aug": None})
    assert np.abs(item["frames"].permute(3, 0, 1, 2).cpu().numpy() - want).max() <= TOL_AFTER
    bad = ds[2]                     # decode failure -> zeros [T,224,224,3] like nexar_videos.py:479-489
    assert tuple(bad["frames"].shape) == (12, 224, 224, 3) and float(bad["frames"].abs().sum()) == 0.0


def test_sliding_windows_view_and_materialised():
    from vision_collision_detection_b200.inference import SlidingWindowTransform, center_window, sliding_window_starts
    from vision_collision_detection_b200.synth import make_clip_torch
    video = make_clip_torch(1200, 48, 80, 5, "dashcam")              # 40 s x 30 fps (small frames)
    for stride, k in ((16, 75), (8, 149)):
        sw = SlidingWindowTransform(window=16, stride=stride, out_dtype=torch.float32)
        view = sw.windows(video)
        shape, nbytes = tuple(view.shape), view.untyped_storage().nbytes()
        assert shape == (k, 3, 16, 224, 224)
        assert nbytes == 1200 * 3 * 224 * 224 * 4          # zero copy: the storage is the per-frame result
        starts = sliding_window_starts(1200, 16, stride)
        for w in (0, k // 2, k - 1):
            direct = sw.tf.forward_batch(video[starts[w]:starts[w] + 16].unsqueeze(0), out_dtype=torch.float32)[0]
            same = bool(torch.equal(view[w], direct))
            assert same, (stride, w)
    sw = SlidingWindowTransform(window=16, stride=8, out_dtype=torch.float32)
    same = bool(torch.equal(sw.windows(video[:200], materialize=True), sw.windows(video[:200]).contiguous()))
    assert same
    short = sw.windows(video[:10])                                      # shorter than a window: last frame repeated
    shape, same = tuple(short.shape), bool(torch.equal(short[0, :, 9], short[0, :, 15]))
    assert shape == (1, 3, 16, 224, 224) and same
    cw = center_window(video, fps=10, duration=5)
    assert tuple(cw.shape) == (1, 3, 50, 224, 224)
    idx = S.window_indices(1200, 50, S.start_frame(1200, 50, "center"))
    want = O.apply_clip_transform(video[idx[:2]].cpu().numpy().transpose(3, 0, 1, 2), O.TransformConfig(mode="val"),
                                  {"flip": False, "aug": None})
    assert np.abs(cw[0, :, :2].cpu().numpy() - want).max() <= TOL_AFTER


def test_host_pipeline_equals_device_path():
    from vision_collision_detection_b200 import create_video_transforms
    from vision_collision_detection_b200.host_pipeline import HostClipPipeline
    from vision_collision_detection_b200.synth import make_clip_np
    kw = dict(mode="train", crop_size=56, enable_custom_augmentation=True)
    tf = create_video_transforms(**kw, out_dtype=torch.bfloat16)
    clips = np.stack([make_clip_np(4, 96, 160, s, "dashcam") for s in range(7)])
    random.seed(5)
    params = tf.sample_params(7, 96, 160)
    pipe = HostClipPipeline(tf, n_clips=7, frames=4, height=96, width=160, clips_per_chunk=2, n_streams=3)
    host_in = pipe.pinned_input()
    host_in.copy_(torch.from_numpy(clips))
    got = pipe.run(host_in, params=params).clone()
    want = tf.forward_batch(torch.from_numpy(clips).cuda(), params=params).cpu()
    same = bool(torch.equal(got, want))
    assert got.dtype == torch.bfloat16 and same
    same = bool(torch.equal(pipe.run(host_in, params=params), want))  # buffers are reusable
    assert same
    # two batches in flight (the steady state bench.py's e2e measures)
    host_in2 = pipe.pinned_input()
    host_in2.copy_(torch.from_numpy(clips[::-1].copy()))
    want2 = tf.forward_batch(torch.from_numpy(clips[::-1].copy()).cuda(), params=params).cpu()
    t1 = pipe.submit(host_in, params=params)
    t2 = pipe.submit(host_in2, params=params)
    r1 = pipe.wait(t1).clone()
    r2 = pipe.wait(t2).clone()
    same = bool(torch.equal(r1, want)) and bool(torch.equal(r2, want2))
    assert same


def test_integration_md_ctypes_example_runs():
    """INTEGRATION.md section B is the binding a reference maintainer would write against the C ABI: execute it as
    written (with a smaller batch) and compare with the Python host on the same frames."""
    import os
    import re
    from vision_collision_detection_b200 import create_video_transforms
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    section = text[text.index("## B."):]
    code = re.search(r"```python\n(.*?)```", section, re.S).group(1)
    code = code.replace("B, T = 32, 16", "B, T = 2, 3")
    code = code.replace('C.CDLL("vision_collision_detection_b200/libnexar_clip_b200.so")',
                        'C.CDLL(%r)' % os.path.join(root, "vision_collision_detection_b200", "libnexar_clip_b200.so"))
    code = code.replace('frames = torch.empty((B, T, 720, 1280, 3), dtype=torch.uint8, device="cuda")',
                        'frames = FRAMES')
    from vision_collision_detection_b200.synth import make_clip_torch
    frames = torch.stack([make_clip_torch(3, 720, 1280, seed=s, kind="dashcam", device="cuda") for s in (1, 2)])
    ns = {"FRAMES": frames}
    exec(compile(code, "INTEGRATION.md#B", "exec"), ns)
    torch.cuda.synchronize()
    got = ns["out"]
    tf = create_video_transforms(mode="val", out_dtype=torch.bfloat16)
    want = tf.forward_batch(frames)
    assert tuple(got.shape) == (2, 3, 3, 224, 224) and got.dtype == torch.bfloat16
    same = bool(torch.equal(got, want))
    assert same


def _decoder_mixed(path):
    n, seed = map(int, path.split("_"))
    if n == 0:
        raise IOError("broken video")
    h, w = ((96, 160), (72, 128), (160, 96))[seed % 3]        # three source resolutions in one dataset
    return _FakeReader(n, seed, h, w)


def test_loader_with_forked_workers_and_mixed_resolutions():
    """The trainers' real setting: DataLoader worker PROCESSES (num_workers=2, nexar_train_distributed.py:97) decode and
    draw the random decisions, nothing in them touches CUDA; batches that mix source resolutions are transformed per
    resolution.  Every clip must equal the oracle on its own frames with its own draws."""
    from torch.utils.data import DataLoader
    from vision_collision_detection_b200 import create_video_transforms
    from vision_collision_detection_b200.videos import GpuAugLoader, GpuDashcamDataset, deferred_collate
    kw = dict(mode="train", crop_size=56, enable_custom_augmentation=True, brightness_range=(0.9, 1.1),
              contrast_range=(0.9, 1.1), saturation_range=(0.9, 1.1), rotation_range=(-5, 5))
    tf = create_video_transforms(**kw)
    rows = [{"id": f"v{i}", "video_type": "Normal", "path": f"{n}_{i}"} for i, n in enumerate([40, 30, 0, 64, 25, 33])]
    ds = GpuDashcamDataset(rows, fps=4, duration=3, transform=tf, sample_strategy="center", decoder=_decoder_mixed, defer=True)
    dl = DataLoader(ds, batch_size=3, shuffle=False, num_workers=2, collate_fn=deferred_collate, pin_memory=True,
                    multiprocessing_context="fork")
    # capture what the workers drew: the deferred batches carry the parameter records
    seen = []

    class Tap:
        def __iter__(self):
            for b in dl:
                seen.append(b)
                yield b

        def __len__(self):
            return len(dl)

    out = [b["frames"] for b in GpuAugLoader(Tap(), tf)]
    assert len(out) == 2 and all(tuple(o.shape) == (3, 12, 56, 56, 3) and o.is_cuda for o in out)
    cfg = O.TransformConfig(mode="train", crop_size=56, enable_custom_augmentation=True, aug=O.AugConfig(rotation_range=(-5, 5)))
    checked = 0
    for bi, b in enumerate(seen):
        assert sum(len(g["index"]) for g in b["groups"]) == sum(b["valid"]) and len(b["groups"]) >= 2
        for g in b["groups"]:
            for k, pos in enumerate(g["index"]):
                clip = g["frames_u8"][k].numpy()
                rec = g["params"][k]
                want = O.apply_clip_transform(clip.transpose(3, 0, 1, 2), cfg, {"flip": rec["flip"], "aug": rec["aug"]})
                got = out[bi][pos].permute(3, 0, 1, 2).float().cpu().numpy()
                assert float(np.abs(got - want).max()) <= TOL_AFTER
                checked += 1
    assert checked == 5 and torch.all(out[0][2] == 0)          # the broken video is the all-zeros clip


def test_cfg4_windows_of_a_720p_video_against_the_oracle():
    """BASELINE configs[3] at the real frame size (the fixed-point fast kernel, not the general one): a 720p video,
    16-frame windows at stride 8, val chain; windows 0 and the last one against the numpy oracle on the same frames."""
    from vision_collision_detection_b200.inference import SlidingWindowTransform, sliding_window_starts
    from vision_collision_detection_b200.synth import make_clip_torch
    video = make_clip_torch(44, 720, 1280, 17, "dashcam")
    sw = SlidingWindowTransform(window=16, stride=8, out_dtype=torch.float32)
    view = sw.windows(video)
    starts = sliding_window_starts(44, 16, 8)
    assert starts == [0, 8, 16, 24] and tuple(view.shape) == (4, 3, 16, 224, 224)
    for k in (0, 3):
        frames = video[starts[k]:starts[k] + 16:5].cpu().numpy()       # frames 0, 5, 10, 15 of the window
        want = O.apply_clip_transform(frames.transpose(3, 0, 1, 2), O.TransformConfig(mode="val"), {"flip": False, "aug": None})
        got = view[k][:, ::5].cpu().numpy()
        assert np.abs(got - want).max() <= TOL_AFTER
    bf = SlidingWindowTransform(window=16, stride=8).windows(video, materialize=True)     # bf16, materialised
    assert bf.is_contiguous() and bf.dtype == torch.bfloat16
    assert (bf.float() - view).abs().max().item() <= 2.0 ** -6           # one bf16 ulp at |v| < 4


def test_window_predictor_is_predict_shaped():
    """nexar_inference.py:211-331 per window: softmax / argmax, class names, probabilities dict, plus the window position."""
    from vision_collision_detection_b200.inference import CLASS_MAP, SlidingWindowTransform, WindowPredictor, classify_outputs
    from vision_collision_detection_b200.synth import make_clip_torch
    video = make_clip_torch(70, 96, 160, 23, "dashcam")
    w_mat = torch.tensor([[1.0, -2.0, 0.5], [-0.3, 0.8, 1.1], [0.2, 0.1, -0.9]], device="cuda")

    def model(x):                                   # [b,3,T,cs,cs] -> logits [b,3]: a fixed linear read-out of channel means
        assert x.is_contiguous() and x.dim() == 5 and x.shape[1] == 3 and x.shape[2] == 16
        return x.float().mean(dim=(2, 3, 4)) @ w_mat

    wp = WindowPredictor(model, window=16, stride=8, batch_size=3)
    res = wp.predict(video, video_path="synthetic.mp4", fps=30.0)
    assert len(res) == 7 and [r["window_start"] for r in res] == [0, 8, 16, 24, 32, 40, 48]
    view = SlidingWindowTransform(window=16, stride=8, out_dtype=torch.float32).windows(video)
    logits = view.mean(dim=(2, 3, 4)) @ w_mat
    probs = torch.softmax(logits, dim=1).cpu().numpy()
    for k, r in enumerate(res):
        assert set(r) >= {"predicted_class", "predicted_class_name", "probabilities", "video_path", "window_start", "center_frame"}
        assert r["predicted_class"] == int(probs[k].argmax()) and r["predicted_class_name"] == CLASS_MAP[r["predicted_class"]]
        assert list(r["probabilities"]) == ["Normal", "Near Collision", "Collision"]
        assert np.allclose(list(r["probabilities"].values()), probs[k], atol=1e-5)
        assert r["center_frame"] == r["window_start"] + 8 and r["video_path"] == "synthetic.mp4"
    p2, c2 = classify_outputs(torch.tensor([[-2.0], [3.0]]), num_classes=2)            # binary head (:251-259)
    assert p2.shape == (2, 1) and c2.tolist() == [0, 1]


class _ReferenceShapedDataset(torch.utils.data.Dataset):
    """The body of nexar_videos.NvidiaDashcamDataset.__getitem__ (:348-496) that matters here, line for line: decode,
    CTHW view, ``self.transform(frames)``, THWC view, dict; any failure -> all-zeros clip (:479-489)."""

    def __init__(self, specs, transform, need=6):
        self.specs, self.transform, self.need = specs, transform, need

    def __len__(self):
        return len(self.specs)

    def __getitem__(self, idx):
        from vision_collision_detection_b200.synth import make_clip_np
        try:
            h, w, seed = self.specs[idx]
            if h == 0:
                raise IOError("broken video")
            frames = torch.from_numpy(make_clip_np(self.need, h, w, seed, "dashcam"))
            frames = frames.permute(3, 0, 1, 2)                  # :441
            frames = self.transform(frames)                      # :445
            frames = frames.permute(1, 2, 3, 0)                  # :451
            return {"frames": frames, "sensor": torch.zeros(self.need, 4), "target": "Normal", "id": f"v{idx}"}
        except Exception:
            return {"frames": torch.zeros(self.need, 224, 224, 3), "sensor": torch.zeros(self.need, 4), "target": "Normal",
                    "id": f"v{idx}"}


@pytest.mark.parametrize("workers", [0, 2])
def test_zero_line_deferred_drop_in_through_an_unmodified_loader(workers):
    """The reference's Dataset lines, DataLoader(num_workers, pin_memory=True) with the DEFAULT collate, and the
    trainers' consumer line, all unchanged; only the factory is the deferred GPU one.  Mixed resolutions and a failed
    item in the batch; every clip against the oracle with the parameters its worker drew."""
    from torch.utils.data import DataLoader
    from vision_collision_detection_b200 import create_video_transforms
    from vision_collision_detection_b200.deferred import DeferredBatch
    from vision_collision_detection_b200.synth import make_clip_np
    kw = dict(mode="train", crop_size=56, enable_custom_augmentation=True, brightness_range=(0.9, 1.1),
              contrast_range=(0.9, 1.1), saturation_range=(0.9, 1.1), rotation_range=(-5, 5))
    tf = create_video_transforms(**kw, deferred=True)
    specs = [(96, 160, 1), (96, 160, 2), (0, 0, 0), (120, 90, 3), (96, 160, 4), (120, 90, 5)]
    loader = DataLoader(_ReferenceShapedDataset(specs, tf), batch_size=3, shuffle=False, num_workers=workers, pin_memory=True)
    cfg = O.TransformConfig(mode="train", crop_size=56, enable_custom_augmentation=True, aug=O.AugConfig(rotation_range=(-5, 5)))
    seen = 0
    for bi, batch in enumerate(loader):
        fb = batch["frames"]
        assert isinstance(fb, DeferredBatch) and fb.is_pinned() and tuple(fb.shape) == (3, 6, 56, 56, 3)
        x = batch["frames"].permute(0, 4, 1, 2, 3).float().to(torch.device("cuda"))      # nexar_train.py:1139 / dvc:708
        assert x.is_cuda and x.dtype == torch.float32 and tuple(x.shape) == (3, 3, 6, 56, 56)
        got = x.cpu().numpy()
        for g in fb.groups:
            for pos, params in zip(g["index"], g["params"]):
                h, w, seed = specs[bi * 3 + pos]
                clip = make_clip_np(6, h, w, seed, "dashcam")
                want = O.apply_clip_transform(clip.transpose(3, 0, 1, 2), cfg, {"flip": params["flip"], "aug": params["aug"]})
                assert np.abs(got[pos] - want).max() <= TOL_AFTER
                seen += 1
        if bi == 0:
            assert np.all(got[2] == 0)                                                   # the failed item stays zeros
    assert seen == 5
