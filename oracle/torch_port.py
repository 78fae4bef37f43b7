"""CPU port of the reference clip transform on torch/torchvision ops.
TEST INFRASTRUCTURE ONLY — this is the timed ``cpu_baseline`` ("port") arm of
bench.py and a second checker next to ``np_oracle``.

It issues the same torch / torchvision calls, in the same per-frame Python
loops and with the same whole-clip float passes, as the reference
(nexar_video_aug.py:705-739 letterbox, :746-755 flip, :200-315 augmentation,
:794-799 normalise, :809-821 driver), so that its CPU time is representative of
the reference's own CPU path.  The random decisions come from
``np_oracle.sample_clip_params`` (same ``random`` draw order).
"""
from __future__ import annotations

import random as _random
from typing import Any, Dict, Optional

import torch
import torchvision.transforms.functional as TF

from . import np_oracle as O


def _per_frame(video: torch.Tensor, fn) -> torch.Tensor:
    return torch.stack([fn(video[:, i]) for i in range(video.shape[1])], dim=1)


def _letterbox(video: torch.Tensor, cs: int) -> torch.Tensor:
    c, _, h, w = video.shape
    new_h, new_w, pad_h, pad_w = O.letterbox_geometry(h, w, cs)

    def one(frame):
        canvas = torch.zeros(c, cs, cs, device=video.device)
        canvas[:, pad_h:pad_h + new_h, pad_w:pad_w + new_w] = TF.resize(frame, [new_h, new_w], antialias=True)
        return canvas

    return _per_frame(video, one)


def _augment(frame: torch.Tensor, p: Dict[str, Any], aug: O.AugConfig) -> torch.Tensor:
    if p.get("skip_augmentation", False):
        return frame
    frame = TF.adjust_hue(
        TF.adjust_saturation(
            TF.adjust_contrast(TF.adjust_brightness(frame, p["brightness"]), p["contrast"]),
            p["saturation"]),
        p["hue"])
    if p["apply_affine"]:
        frame = TF.affine(frame, angle=p["rotation"], translate=[p["translate_x"], p["translate_y"]],
                          scale=p["scale"], shear=p["shear"],
                          interpolation=TF.InterpolationMode.BILINEAR, fill=0)
    if p["apply_grayscale"]:
        frame = TF.rgb_to_grayscale(frame, num_output_channels=3)
    if p["apply_noise"]:
        frame = torch.clamp(frame + torch.randn_like(frame) * aug.noise_level, 0, 1)
    if p["apply_blur"]:
        frame = TF.gaussian_blur(frame.unsqueeze(0), kernel_size=int(aug.blur_sigma * 4) * 2 + 1,
                                 sigma=aug.blur_sigma).squeeze(0)
    if p["apply_posterization"]:
        frame = TF.posterize((frame * 255).byte(), p["posterization_bits"]).float() / 255.0
    if p["apply_solarization"]:
        frame = TF.solarize(frame, aug.solarization_threshold)
    if p["apply_color_inversion"]:
        frame = 1.0 - frame
    if p["apply_cutout"]:
        for top, left, ch, cw in p["cutout_boxes"]:
            frame[:, top:top + ch, left:left + cw] = 0
    return frame


def apply_clip_transform(video: torch.Tensor, cfg: O.TransformConfig, params: Dict[str, Any]) -> torch.Tensor:
    """[C,T,H,W] uint8/float CPU tensor -> [C,T,cs,cs] float32."""
    if video.dtype != torch.float32:
        video = video.float()
    if video.max() > 1.0:
        video = video / 255.0
    video = _letterbox(video, cfg.crop_size)
    if params["flip"]:
        video = _per_frame(video, TF.hflip)
    if params["aug"] is not None:
        video = _per_frame(video, lambda f: _augment(f, params["aug"], cfg.aug))
    if cfg.normalize:
        c = video.shape[0]
        mean = torch.tensor(cfg.video_mean).view(c, 1, 1, 1)
        std = torch.tensor(cfg.video_std).view(c, 1, 1, 1)
        video = (video - mean) / std
    return video


def clip_transform(video: torch.Tensor, cfg: O.TransformConfig, rng=_random) -> torch.Tensor:
    return apply_clip_transform(video, cfg, O.sample_clip_params(cfg, rng))
