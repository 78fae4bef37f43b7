"""BASELINE.json configs[0] frozen from the UNMODIFIED reference: `create_video_transforms` in its three live
configurations on ONE synthetic 16-frame 1280x720 uint8 clip -> 16 x 224 x 224, CPU, seeded.

Run in the build container only (needs /root/reference):
    python tests/golden/make_cfg1_golden.py
Writes tests/golden/golden_cfg1_16f.npz.  The full outputs are 9.6 MB each, so the fixture keeps, per configuration,
the output sampled every 5th row / column of every frame ([3,16,45,45]) plus the mean of every frame and channel over
ALL pixels ([3,16], float64); the input clip is regenerated from its seed.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference_aug  # noqa: E402
from vision_collision_detection_b200.synth import make_clip_np  # noqa: E402

CLIP = (16, 720, 1280, 101, "dashcam")
SEED = 2024
KW = {
    "val": dict(mode="val"),
    "train": dict(mode="train"),
    "custom": dict(mode="train", enable_custom_augmentation=True, brightness_range=(0.9, 1.1),
                   contrast_range=(0.9, 1.1), saturation_range=(0.9, 1.1), rotation_range=(-5, 5)),
}
STRIDE = 5


def main():
    ref = import_reference_aug()
    video = torch.from_numpy(make_clip_np(*CLIP)).permute(3, 0, 1, 2)
    out = {}
    for name, kw in KW.items():
        tf = ref.create_video_transforms(**kw)
        random.seed(SEED)
        torch.manual_seed(SEED)
        o = tf(video).numpy()
        assert o.shape == (3, 16, 224, 224)
        out[f"{name}_sub"] = o[:, :, ::STRIDE, ::STRIDE].copy()
        out[f"{name}_mean"] = o.astype(np.float64).mean(axis=(2, 3))
        print(name, o.shape, float(o.min()), float(o.max()))
    np.savez_compressed(os.path.join(HERE, "golden_cfg1_16f.npz"), **out)


if __name__ == "__main__":
    main()
