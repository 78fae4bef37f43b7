"""NV12 decoder-surface source: the stated BT.601 conversion, its numpy oracle, the synthetic
surface generator, and - on the GPU - that the NV12 path is the RGB path bit for bit once the surfaces are converted."""
import random

import numpy as np
import pytest
import torch

from oracle import nv12_oracle as N
from vision_collision_detection_b200.synth import make_clip_np, rgb_to_nv12


def test_oracle_known_answers():
    """Points of the BT.601 limited-range cube: black, white, the primaries (tables of ITU-R BT.601 / the usual 8-bit form)."""
    def px(y, u, v):
        nv = np.array([[y, y], [y, y], [u, v]], np.uint8)          # one 2 x 2 block
        return tuple(int(c) for c in N.nv12_to_rgb(nv)[0, 0])
    assert px(16, 128, 128) == (0, 0, 0)
    assert px(235, 128, 128) == (255, 255, 255)
    assert px(126, 128, 128) == (128, 128, 128)
    assert px(81, 90, 240) == (255, 0, 0)
    assert px(145, 54, 34) == (0, 255, 1)     # B = (298 * 129 - 516 * 74 + 128) >> 8 = 386 >> 8 = 1: the 8-bit form is not exact
    assert px(41, 240, 110) == (0, 0, 255)
    assert px(0, 128, 128) == (0, 0, 0) and px(255, 128, 128) == (255, 255, 255)      # out-of-range luma is clipped
    nv = np.array([[16, 235], [126, 81], [128, 128]], np.uint8)   # chroma is shared by the 2 x 2 block
    assert N.nv12_to_rgb(nv)[..., 0].tolist() == [[0, 255], [128, 76]]


def test_synthetic_surfaces_numpy_equals_torch_and_round_trip():
    clip = make_clip_np(2, 48, 64, 5, "dashcam")
    a = rgb_to_nv12(clip)
    b = rgb_to_nv12(torch.from_numpy(clip)).numpy()
    assert a.shape == (2, 72, 64) and a.dtype == np.uint8 and np.array_equal(a, b)
    smooth = np.broadcast_to(np.linspace(20, 230, 64).astype(np.uint8)[None, None, :, None], (1, 48, 64, 3)).copy()
    back = N.nv12_to_rgb(rgb_to_nv12(smooth))
    assert np.abs(back.astype(int) - smooth.astype(int)).max() <= 3      # gray ramp: rounding only
    with pytest.raises(ValueError):
        rgb_to_nv12(np.zeros((3, 5, 3), np.uint8))


KW_CUSTOM = dict(mode="train", enable_custom_augmentation=True, brightness_range=(0.9, 1.1), contrast_range=(0.9, 1.1),
                 saturation_range=(0.9, 1.1), rotation_range=(-5, 5))


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,cs,kind", [(720, 1280, 224, "dashcam"), (720, 1280, 224, "noise"), (90, 150, 64, "noise"), (360, 648, 112, "dashcam")])
@pytest.mark.parametrize("kw", [dict(mode="val"), KW_CUSTOM], ids=["val", "custom"])
def test_nv12_path_equals_rgb_path_on_converted_frames(h, w, cs, kind, kw):
    """16-byte-aligned widths take the vector conversion kernel, the others the 2 x 2 one; both must produce exactly the
    bytes of the numpy oracle, i.e. the NV12 result equals the RGB result on oracle-converted frames (fp32, bit for bit)."""
    from vision_collision_detection_b200 import create_video_transforms
    nv = rgb_to_nv12(np.stack([make_clip_np(3, h, w, 11 + i, kind) for i in range(2)]))
    if kind == "noise":                      # exercise the clipping branches: arbitrary luma / chroma bytes
        nv = np.random.RandomState(3).randint(0, 256, nv.shape).astype(np.uint8)
    rgb = N.nv12_to_rgb(nv)
    tf = create_video_transforms(**kw, crop_size=cs)
    random.seed(5)
    params = tf.sample_params(2, h, w)
    got = tf.forward_batch(torch.from_numpy(nv).cuda(), params=params, pixel_format="nv12").cpu().numpy()
    want = tf.forward_batch(torch.from_numpy(rgb).cuda(), params=params).cpu().numpy()
    assert got.shape == (2, 3, 3, cs, cs)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_nv12_frame_gather_max_rule_and_errors():
    from vision_collision_detection_b200 import create_video_transforms
    h, w, cs = 96, 160, 64
    nv = rgb_to_nv12(np.stack([make_clip_np(4, h, w, 70 + i, "dashcam") for i in range(2)]))
    nv[1, :, :h] = 16                         # second clip: black luma, neutral chroma -> RGB 0: the "max <= 1" rule
    nv[1, :, h:] = 128
    rgb = N.nv12_to_rgb(nv)
    assert rgb[1].max() == 0
    tf = create_video_transforms(mode="val", crop_size=cs)
    index = torch.tensor([[3, 1, 1], [4, 7, 5]])
    got = tf.forward_batch(torch.from_numpy(nv).cuda(), frame_index=index, pixel_format="nv12").cpu().numpy()
    want = tf.forward_batch(torch.from_numpy(rgb).cuda(), frame_index=index).cpu().numpy()
    assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        tf.forward_batch(torch.zeros((1, 2, 100, 160), dtype=torch.uint8).cuda(), pixel_format="nv12")   # 100 is not H * 3 / 2
    with pytest.raises(ValueError):
        tf.forward_batch(torch.from_numpy(nv).cuda(), pixel_format="yuv444")


@pytest.mark.gpu
def test_host_pipeline_nv12():
    from vision_collision_detection_b200 import create_video_transforms
    from vision_collision_detection_b200.host_pipeline import HostClipPipeline
    h, w, cs = 180, 320, 112
    tf = create_video_transforms(**KW_CUSTOM, crop_size=cs, out_dtype=torch.bfloat16)
    nv = torch.from_numpy(rgb_to_nv12(np.stack([make_clip_np(4, h, w, 90 + i, "dashcam") for i in range(5)])))
    pipe = HostClipPipeline(tf, n_clips=5, frames=4, height=h, width=w, clips_per_chunk=2, pixel_format="nv12")
    host = pipe.pinned_input()
    assert tuple(host.shape) == (5, 4, h * 3 // 2, w)
    host.copy_(nv)
    random.seed(9)
    params = tf.sample_params(5, h, w)
    out = pipe.run(host, params=params).float().numpy().copy()
    want = tf.forward_batch(torch.from_numpy(N.nv12_to_rgb(nv.numpy())).cuda(), params=params).float().cpu().numpy()
    assert np.array_equal(out, want)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,pad", [(720, 1280, 64), (96, 160, 16), (90, 150, 7)])
def test_nv12_padded_pitch_through_the_engine(h, w, pad):
    """Decoder surfaces have a pitch: the C ABI takes it as src_row_stride (Y and UV planes share it).  16-byte pitches
    of 16-pixel-multiple widths take the vector conversion kernel, everything else the 2 x 2 one; the padding bytes must
    never be read (they are 255 here) and the result equals the packed surfaces' bit for bit."""
    from vision_collision_detection_b200 import create_video_transforms
    from vision_collision_detection_b200.engine import get_engine, _alloc_out
    from vision_collision_detection_b200.params import pack_clip_params
    t, cs = 2, 64
    nv = torch.from_numpy(rgb_to_nv12(make_clip_np(t, h, w, 31, "dashcam"))).cuda()      # [t, h*3/2, w]
    tf = create_video_transforms(mode="val", crop_size=cs)
    want = tf.forward_batch(nv.unsqueeze(0), pixel_format="nv12").cpu().numpy()
    eng = get_engine(nv.device)
    plan = tf._plan(eng, h, w, "nv12")
    pitch = w + pad
    rows = h * 3 // 2
    buf = torch.full((t, rows, pitch), 255, dtype=torch.uint8, device="cuda")
    buf[:, :, :w] = nv
    offsets = torch.arange(t, dtype=torch.int64, device="cuda") * (rows * pitch)
    packed, any_flags = pack_clip_params(tf.last_params, cs, tf.video_aug)
    out, strides = _alloc_out("BCTHW", 1, t, cs, torch.float32, nv.device)
    eng.run(plan, buf, offsets, 1, t, eng.upload_params(packed), any_flags, out, strides,
            tf.normalize, tf.video_mean, tf.video_std, src_row_stride=pitch)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)
