"""Generate the committed golden vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/golden_*.npz + golden_meta.json.  Inputs are NOT stored:
they are regenerated from (t, h, w, seed, kind) by
vision_collision_detection_b200.synth.make_clip_np, which is integer-only and
therefore identical everywhere.  Outputs are the reference's own tensors.
"""
import json
import os
import random
import sys

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference_aug  # noqa: E402
from vision_collision_detection_b200.synth import make_clip_np  # noqa: E402

ref = import_reference_aug()

# kwargs of the live call sites (reference file:line in the comment)
KW_VAL = dict(mode="val")                                            # nexar_videos.py:947, nexar_inference.py:218
KW_TRAIN = dict(mode="train")                                        # nexar_videos.py:936
KW_CUSTOM = dict(mode="train", enable_custom_augmentation=True,      # nexar_videos.py:2003-2010
                 brightness_range=(0.9, 1.1), contrast_range=(0.9, 1.1),
                 saturation_range=(0.9, 1.1), rotation_range=(-5, 5))
KW_NCWV = dict(mode="train", enable_custom_augmentation=True,        # nexar_complete_with_validation.py:1208-1225
               brightness_range=(0.9, 1.1), contrast_range=(0.9, 1.1), saturation_range=(0.9, 1.1),
               hue_range=(-0.05, 0.05), rotation_range=(-7, 7), scale_range=(0.95, 1.1),
               translate_range=(0.0, 0.07), grayscale_prob=0.02, blur_sigma=0.5, cutout_prob=0.1,
               cutout_count=(1, 2), cutout_size_range=(0.1, 0.15), horizontal_flip_prob=0.5,
               aug_probability=0.9)
KW_ALLFX = dict(mode="train", enable_custom_augmentation=True, hue_range=(-0.3, 0.3),
                rotation_range=(-25, 25), scale_range=(0.7, 1.3), shear_range=(-10, 10),
                translate_range=(0.0, 0.2), grayscale_prob=0.4, blur_sigma=0.8, cutout_prob=0.7,
                color_inversion_prob=0.3, solarization_prob=0.4, posterization_prob=0.0)
KW_POSTER = dict(mode="train", enable_custom_augmentation=True, posterization_prob=1.0,
                 rotation_range=(0, 0), scale_range=(1, 1), shear_range=(0, 0), translate_range=(0, 0))

CASES = [
    # name, (t,h,w,seed,kind), kwargs, random seed
    ("val_720p_noise", (1, 720, 1280, 0, "noise"), KW_VAL, 0),
    ("train_720p_dashcam", (1, 720, 1280, 1, "dashcam"), KW_TRAIN, 3),
    ("custom_720p_dashcam", (2, 720, 1280, 2, "dashcam"), KW_CUSTOM, 1234),
    ("val_1080p_noise", (1, 1080, 1920, 3, "noise"), KW_VAL, 0),
    ("val_720p_320", (1, 720, 1280, 4, "dashcam"), dict(mode="val", crop_size=320), 0),
    ("val_small", (3, 96, 160, 5, "noise"), dict(mode="val", crop_size=56), 0),
    ("val_portrait", (2, 160, 96, 6, "dashcam"), dict(mode="val", crop_size=56), 0),
    ("train_portrait_flip", (2, 160, 96, 6, "dashcam"), dict(mode="train", crop_size=56, horizontal_flip_prob=1.0), 0),
    ("val_upscale", (2, 40, 60, 7, "noise"), dict(mode="val", crop_size=56), 0),
    ("val_odd", (2, 97, 131, 8, "noise"), dict(mode="val", crop_size=64), 0),
    ("val_nonorm", (2, 96, 160, 9, "dashcam"), dict(mode="val", crop_size=56, normalize=False), 0),
    ("val_meanstd", (2, 96, 160, 9, "dashcam"), dict(mode="val", crop_size=56, video_mean=(0.485, 0.456, 0.406),
                                                       video_std=(0.229, 0.224, 0.225)), 0),
] + [
    (f"custom_small_s{s}", (3, 96, 160, 10 + s, "dashcam"), dict(KW_CUSTOM, crop_size=56), s) for s in range(4)
] + [
    (f"ncwv_small_s{s}", (3, 96, 160, 20 + s, "dashcam"), dict(KW_NCWV, crop_size=56), s) for s in range(6)
] + [
    (f"allfx_small_s{s}", (2, 96, 160, 30 + s, "noise" if s % 2 else "dashcam"), dict(KW_ALLFX, crop_size=64), s)
    for s in range(8)
] + [
    (f"poster_small_s{s}", (2, 96, 160, 40 + s, "dashcam"), dict(KW_POSTER, crop_size=56), s) for s in range(2)
]


def _jsonable(o):
    if isinstance(o, float):
        return {"f": o.hex()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(x) for x in o]
    if isinstance(o, dict):
        return {k: _jsonable(v) for k, v in o.items()}
    return o


def trace_params(kwargs, seed):
    """Record what the reference draws: flip decision + the aug params dict."""
    tf = ref.create_video_transforms(**kwargs)
    rec = {"flip": None, "aug": None}
    cs = kwargs.get("crop_size", 224)
    random.seed(seed)
    # replay the draws in the order VideoTransform.forward makes them
    if kwargs.get("mode", "train") == "train" and kwargs.get("horizontal_flip_prob", 0.5) > 0:
        rec["flip"] = random.random() < kwargs.get("horizontal_flip_prob", 0.5)
    for t in tf.transforms:
        if isinstance(t, ref.VideoAugmentation):
            rec["aug"] = t._sample_augmentation_parameters((3, 1, cs, cs))
    return rec


def main():
    meta = {"torch": torch.__version__, "torchvision": torchvision.__version__, "numpy": np.__version__,
            "cases": {}}
    for name, (t, h, w, seed, kind), kwargs, rseed in CASES:
        clip = make_clip_np(t, h, w, seed, kind)
        video = torch.from_numpy(clip).permute(3, 0, 1, 2)
        tf = ref.create_video_transforms(**kwargs)
        random.seed(rseed)
        torch.manual_seed(rseed)
        out = tf(video).numpy()
        params = trace_params(kwargs, rseed)
        np.savez_compressed(os.path.join(HERE, f"golden_{name}.npz"), out=out)
        meta["cases"][name] = {"input": [t, h, w, seed, kind], "kwargs": _jsonable(kwargs), "random_seed": rseed,
                               "params": _jsonable(params), "out_shape": list(out.shape)}
        print(name, out.shape, float(out.min()), float(out.max()))

    # special inputs ---------------------------------------------------------
    special = {}
    tfv = ref.create_video_transforms(mode="val", crop_size=56)
    b01 = (make_clip_np(2, 96, 160, 50, "noise") & 1).astype(np.uint8)          # clip max == 1: NOT rescaled
    special["max1_u8"] = tfv(torch.from_numpy(b01).permute(3, 0, 1, 2)).numpy()
    zeros = np.zeros((2, 96, 160, 3), np.uint8)
    special["zeros_u8"] = tfv(torch.from_numpy(zeros).permute(3, 0, 1, 2)).numpy()
    f01 = make_clip_np(2, 96, 160, 51, "dashcam").astype(np.float32) / 255.0     # float input in [0,1]
    special["float01"] = tfv(torch.from_numpy(f01).permute(3, 0, 1, 2)).numpy()
    f255 = make_clip_np(2, 96, 160, 52, "dashcam").astype(np.float32)            # float input in [0,255]
    special["float255"] = tfv(torch.from_numpy(f255).permute(3, 0, 1, 2)).numpy()
    # dead-code factory R11: resize short side then centre / random crop
    clip = make_clip_np(2, 96, 160, 53, "dashcam")
    video = torch.from_numpy(clip).permute(3, 0, 1, 2)
    tfs = ref.create_video_transform(mode="val", min_size=56, crop_size=56, normalize=False)
    x = video.float() / 255.0
    resized = tfs.transforms[0](x)
    special["r11_resized"] = resized.numpy()
    special["r11_center"] = tfs.transforms[1](resized, use_letterbox=False).numpy()
    tft = ref.create_video_transform(mode="train", min_size=56, max_size=None, crop_size=56, normalize=False)
    random.seed(11)
    special["r11_random"] = tft.transforms[1](resized, use_letterbox=False).numpy()
    random.seed(11)
    meta["r11_random_left"] = random.randint(0, resized.shape[-1] - 56)  # h == cs -> no 'top' draw
    # the never-called factory's default forward: short-side resize, then a SECOND antialiased resize (letterbox)
    tfd = ref.create_video_transform(mode="val", min_size=40, crop_size=56)
    special["singular_val_40_56"] = tfd(video).numpy()
    tfd = ref.create_video_transform(mode="train", min_size=40, max_size=None, crop_size=56, horizontal_flip_prob=1.0)
    random.seed(5)
    special["singular_train_flip_40_56"] = tfd(video).numpy()
    np.savez_compressed(os.path.join(HERE, "golden_special.npz"), **special)

    # letterbox geometry table: measured from the reference on all-ones inputs ---
    geo = {}
    for (h, w, cs) in [(720, 1280, 224), (1080, 1920, 224), (720, 1280, 320), (1280, 720, 224), (96, 160, 56),
                       (160, 96, 56), (40, 60, 56), (97, 131, 64), (480, 640, 224), (224, 224, 224), (1, 7, 8)]:
        tf1 = ref.create_video_transforms(mode="val", crop_size=cs, normalize=False)
        o = tf1(torch.full((3, 1, h, w), 255, dtype=torch.uint8))[0, 0].numpy()
        rows = np.where(o.max(axis=1) > 0)[0]
        cols = np.where(o.max(axis=0) > 0)[0]
        geo[f"{h}x{w}->{cs}"] = [int(len(rows)), int(len(cols)), int(rows[0]) if len(rows) else 0,
                                 int(cols[0]) if len(cols) else 0]
    meta["geometry"] = geo

    # RNG draw trace under random.seed(7) -----------------
    random.seed(7)
    meta["rng_seed7_first"] = [random.random().hex(), random.random().hex(), random.uniform(0.9, 1.1).hex()]

    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", len(meta["cases"]), "cases")


if __name__ == "__main__":
    main()
