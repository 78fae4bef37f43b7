"""Helpers to read the committed golden vectors (tests/golden/)."""
import json
import os

import numpy as np

from oracle import np_oracle as O
from vision_collision_detection_b200.synth import make_clip_np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

with open(os.path.join(GOLDEN_DIR, "golden_meta.json")) as _f:
    META = json.load(_f)


def decode(o):
    if isinstance(o, dict):
        if set(o.keys()) == {"f"}:
            return float.fromhex(o["f"])
        return {k: decode(v) for k, v in o.items()}
    if isinstance(o, list):
        return [decode(x) for x in o]
    return o


def case_names(prefix=""):
    return sorted(n for n in META["cases"] if n.startswith(prefix))


_AUG_KEYS = {f.name for f in O.AugConfig.__dataclass_fields__.values()}
# kwargs the reference factory accepts but never forwards (nexar_video_aug.py:762-788)
_DROPPED = {"aug_probability", "cutout_count", "cutout_size_range", "posterization_bits_range",
            "solarization_threshold"}


def oracle_config(kwargs) -> O.TransformConfig:
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in decode(kwargs).items()}
    aug = O.AugConfig(**{k: v for k, v in kw.items() if k in _AUG_KEYS and k not in _DROPPED})
    top = {k: v for k, v in kw.items() if k in ("mode", "crop_size", "normalize", "video_mean", "video_std",
                                                "horizontal_flip_prob", "enable_custom_augmentation")}
    return O.TransformConfig(aug=aug, **top)


def load_case(name):
    """-> dict(clip u8 [T,H,W,3], kwargs, cfg, params, out [C,T,cs,cs], random_seed)"""
    c = META["cases"][name]
    t, h, w, seed, kind = c["input"]
    params = decode(c["params"])
    if params["flip"] is None:
        params["flip"] = False
    if params["aug"] is not None and "cutout_boxes" in params["aug"]:
        params["aug"]["cutout_boxes"] = [tuple(b) for b in params["aug"]["cutout_boxes"]]
    kwargs = {k: (tuple(v) if isinstance(v, list) else v) for k, v in decode(c["kwargs"]).items()}
    return {
        "clip": make_clip_np(t, h, w, seed, kind),
        "kwargs": kwargs,
        "cfg": oracle_config(c["kwargs"]),
        "params": params,
        "out": np.load(os.path.join(GOLDEN_DIR, f"golden_{name}.npz"))["out"],
        "random_seed": c["random_seed"],
    }


def special():
    return np.load(os.path.join(GOLDEN_DIR, "golden_special.npz"))
