"""Numpy restatement of the reference clip transform.  TEST INFRASTRUCTURE ONLY.

Each function restates, in plain numpy float32 arithmetic, one stage of
``/root/reference/nexar_video_aug.py`` (the reference) or of the third-party
code that stage calls: torchvision 0.26.0 ``transforms/_functional_tensor.py`` /
``transforms/functional.py`` (cited as ``tv:``) and ATen (torch 2.11.0)
``_upsample_bilinear2d_aa`` / ``grid_sampler_2d`` (cited as ``aten:``).  Those
libraries are not vendored by the reference and are unpinned there; the
versions above are the ones the golden vectors were produced with.

Parity pinning: checked against outputs of the unmodified reference in
``tests/test_oracle_vs_golden.py`` (committed fixtures; ``tests/golden/make_golden.py``
regenerates them where ``/root/reference`` exists).

Frames are handled as float32 ``[C, H, W]`` planes like the reference does.
"""
from __future__ import annotations

import math
import random as _random
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------
# R3: letterbox geometry + antialiased (triangle filter) resize
# ----------------------------------------------------------------------------
def letterbox_geometry(h: int, w: int, cs: int) -> Tuple[int, int, int, int]:
    """nexar_video_aug.py:713-719 — python float64 truncation, floor-div pads."""
    scale = min(cs / h, cs / w)
    new_h = int(h * scale)
    new_w = int(w * scale)
    return new_h, new_w, (cs - new_h) // 2, (cs - new_w) // 2


def short_side_geometry(h: int, w: int, size: int) -> Tuple[int, int]:
    """nexar_video_aug.py:411-415 (dead factory's resize_tensor)."""
    if h > w:
        return size * h // w, size
    return size, size * w // h


def aa_taps(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """aten: UpSampleKernel.cpp ``_compute_indices_min_size_weights_aa`` for
    bilinear (interp_size 2), align_corners=False, scale = in/out, all in float32
    (opmath of float is float; verified against F.resize: max diff 1.2e-7)."""
    scale = F32(in_size) / F32(out_size)
    if scale >= 1.0:
        support = scale
        invscale = F32(1.0) / scale
    else:
        support = F32(1.0)
        invscale = F32(1.0)
    kmax = int(math.ceil(float(support))) * 2 + 1
    xmin = np.zeros(out_size, np.int64)
    xsize = np.zeros(out_size, np.int64)
    wts = np.zeros((out_size, kmax), F32)
    for i in range(out_size):
        center = scale * F32(i + 0.5)
        lo = max(int(center - support + F32(0.5)), 0)
        hi = min(int(center + support + F32(0.5)), in_size)
        n = min(max(hi - lo, 0), kmax)
        total = F32(0.0)
        w = np.zeros(n, F32)
        for j in range(n):
            x = (F32(j + lo) - center + F32(0.5)) * invscale
            ax = abs(x)
            w[j] = F32(1.0) - ax if ax < 1.0 else F32(0.0)
            total = F32(total + w[j])
        if total != 0.0:
            w = (w / total).astype(F32)
        xmin[i], xsize[i] = lo, n
        wts[i, :n] = w
    return xmin, xsize, wts


def _resample_last_axis(x: np.ndarray, out_size: int) -> np.ndarray:
    """One separable pass over the last axis; t = s0*w0, t += sj*wj (aten:
    ``interpolate_aa_single_dim``)."""
    in_size = x.shape[-1]
    if in_size == out_size:
        return x.copy()
    xmin, xsize, wts = aa_taps(in_size, out_size)
    k = wts.shape[1]
    out = None
    for j in range(k):
        idx = np.minimum(xmin + j, in_size - 1)
        term = x[..., idx] * wts[:, j]
        out = term if out is None else out + term
    return out.astype(F32)


def aa_resize(frame: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """tv:_functional_tensor.py:441-475 resize(antialias=True) on [C,H,W] f32:
    horizontal pass first, then vertical (aten: separable_upsample_generic_Nd)."""
    hor = _resample_last_axis(frame.astype(F32), new_w)
    ver = _resample_last_axis(np.ascontiguousarray(hor.transpose(0, 2, 1)), new_h)
    return np.ascontiguousarray(ver.transpose(0, 2, 1))


def prologue(video: np.ndarray) -> np.ndarray:
    """nexar_video_aug.py:809-815 — to float32; divide by 255 iff clip max > 1."""
    v = video.astype(F32)
    if v.size and v.max() > 1.0:
        v = v / F32(255.0)
    return v


def letterbox_resize(video: np.ndarray, cs: int) -> np.ndarray:
    """nexar_video_aug.py:705-739; video is [C,T,H,W] float32."""
    c, t, h, w = video.shape
    new_h, new_w, pad_h, pad_w = letterbox_geometry(h, w, cs)
    out = np.zeros((c, t, cs, cs), F32)
    for i in range(t):
        out[:, i, pad_h:pad_h + new_h, pad_w:pad_w + new_w] = aa_resize(video[:, i], new_h, new_w)
    return out


def resize_short_side(video: np.ndarray, size: int) -> np.ndarray:
    """nexar_video_aug.py:407-424."""
    c, t, h, w = video.shape
    new_h, new_w = short_side_geometry(h, w, size)
    return np.stack([aa_resize(video[:, i], new_h, new_w) for i in range(t)], axis=1)


def crop_offsets(h: int, w: int, cs: int, random_crop: bool, rng=_random) -> Tuple[int, int]:
    """nexar_video_aug.py:466-473 — center, or randint top then left."""
    if not random_crop:
        return (h - cs) // 2, (w - cs) // 2
    top = rng.randint(0, h - cs) if h > cs else 0
    left = rng.randint(0, w - cs) if w > cs else 0
    return top, left


# ----------------------------------------------------------------------------
# R6: colour ops (tv:_functional_tensor.py:171-321)
# ----------------------------------------------------------------------------
def _blend(a: np.ndarray, b, ratio: float) -> np.ndarray:
    """tv:_functional_tensor.py:258-261."""
    r = F32(float(ratio))
    q = F32(1.0 - float(ratio))
    return np.clip((r * a).astype(F32) + (q * np.asarray(b, F32)).astype(F32), F32(0), F32(1)).astype(F32)


def rgb_to_gray(img: np.ndarray) -> np.ndarray:
    """tv:_functional_tensor.py:147-156."""
    r, g, b = img[0], img[1], img[2]
    return ((F32(0.2989) * r).astype(F32) + (F32(0.587) * g).astype(F32) + (F32(0.114) * b).astype(F32)).astype(F32)


def adjust_brightness(img, f):
    return _blend(img, np.zeros_like(img), f)


def adjust_contrast(img, f):
    """tv:_functional_tensor.py:181-195 — mean of the gray frame (pads included)."""
    mean = F32(rgb_to_gray(img).mean(dtype=np.float64))
    return _blend(img, mean, f)


def adjust_saturation(img, f):
    return _blend(img, rgb_to_gray(img)[None], f)


def _rgb2hsv(img):
    """tv:_functional_tensor.py:264-300."""
    r, g, b = img[0], img[1], img[2]
    maxc = img.max(axis=0)
    minc = img.min(axis=0)
    eqc = maxc == minc
    cr = maxc - minc
    ones = np.ones_like(maxc)
    s = cr / np.where(eqc, ones, maxc)
    div = np.where(eqc, ones, cr)
    rc = (maxc - r) / div
    gc = (maxc - g) / div
    bc = (maxc - b) / div
    hr = (maxc == r) * (bc - gc)
    hg = ((maxc == g) & (maxc != r)) * (F32(2.0) + rc - bc)
    hb = ((maxc != g) & (maxc != r)) * (F32(4.0) + gc - rc)
    h = (hr + hg + hb).astype(F32)
    h = np.fmod((h / F32(6.0) + F32(1.0)).astype(F32), F32(1.0)).astype(F32)
    return h, s.astype(F32), maxc


def _hsv2rgb(h, s, v):
    """tv:_functional_tensor.py:303-321."""
    h6 = (h * F32(6.0)).astype(F32)
    i = np.floor(h6)
    f = (h6 - i).astype(F32)
    i = i.astype(np.int32) % 6
    one = F32(1.0)
    p = np.clip(v * (one - s), 0, 1).astype(F32)
    q = np.clip(v * (one - (s * f).astype(F32)), 0, 1).astype(F32)
    t = np.clip(v * (one - (s * (one - f)).astype(F32)), 0, 1).astype(F32)
    r = np.choose(i, [v, q, p, p, t, v])
    g = np.choose(i, [t, v, v, q, p, p])
    b = np.choose(i, [p, p, t, v, v, q])
    return np.stack([r, g, b]).astype(F32)


def adjust_hue(img, f):
    """tv:_functional_tensor.py:198-221."""
    if not (-0.5 <= f <= 0.5):
        raise ValueError(f"hue_factor ({f}) is not in [-0.5, 0.5].")
    h, s, v = _rgb2hsv(img)
    h = np.mod((h + F32(f)).astype(F32), F32(1.0)).astype(F32)
    return _hsv2rgb(h, s, v)


# ----------------------------------------------------------------------------
# R7: affine (tv:functional.py:1006-1064,1135-1243; tv:_functional_tensor.py:545-618)
# ----------------------------------------------------------------------------
def inverse_affine_matrix(angle: float, translate: Sequence[float], scale: float, shear: Sequence[float]) -> List[float]:
    """tv:functional.py:1006-1064 with center=(0,0), inverted=True (float64)."""
    rot = math.radians(angle)
    sx = math.radians(shear[0])
    sy = math.radians(shear[1])
    tx, ty = translate
    a = math.cos(rot - sy) / math.cos(sy)
    b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c = math.sin(rot - sy) / math.cos(sy)
    d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m = [x / scale for x in m]
    m[2] += m[0] * (-tx) + m[1] * (-ty)
    m[5] += m[3] * (-tx) + m[4] * (-ty)
    return m


def affine_bilinear_fill0(img: np.ndarray, matrix: Sequence[float]) -> np.ndarray:
    """tv F_t.affine(bilinear, fill=0): fp32 grid = base @ (theta^T / [w/2,h/2]);
    grid_sample(zeros, align_corners=False) on [img | ones]; img*mask."""
    c, h, w = img.shape
    theta = np.asarray(matrix, F32).reshape(2, 3)
    xs = (np.arange(w, dtype=np.float64) - w * 0.5 + 0.5).astype(F32)
    ys = (np.arange(h, dtype=np.float64) - h * 0.5 + 0.5).astype(F32)
    rt = np.stack([theta[0] / F32(0.5 * w), theta[1] / F32(0.5 * h)], axis=1).astype(F32)  # [3,2]
    gx = (xs[None, :] * rt[0, 0] + ys[:, None] * rt[1, 0] + rt[2, 0]).astype(F32)
    gy = (xs[None, :] * rt[0, 1] + ys[:, None] * rt[1, 1] + rt[2, 1]).astype(F32)
    ix = (((gx + F32(1)) * F32(w) - F32(1)) / F32(2)).astype(F32)
    iy = (((gy + F32(1)) * F32(h) - F32(1)) / F32(2)).astype(F32)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    wx1 = (ix - x0).astype(F32)
    wy1 = (iy - y0).astype(F32)
    wx0 = (F32(1) - wx1).astype(F32)
    wy0 = (F32(1) - wy1).astype(F32)
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    src = np.concatenate([img.astype(F32), np.ones((1, h, w), F32)], axis=0)
    out = np.zeros((c + 1, h, w), F32)
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xx = x0 + dx
            yy = y0 + dy
            ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            val = src[:, np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)] * ok
            out = out + val * (wx * wy).astype(F32)
    mask = out[-1:]
    return (out[:-1] * mask + (F32(1.0) - mask) * F32(0.0)).astype(F32)


# ----------------------------------------------------------------------------
# R8: effects (nexar_video_aug.py:239-272)
# ----------------------------------------------------------------------------
def gaussian_kernel1d(ksize: int, sigma: float) -> np.ndarray:
    """tv:_functional_tensor.py:727-735."""
    half = (ksize - 1) * 0.5
    x = np.linspace(-half, half, ksize).astype(F32)
    pdf = np.exp(F32(-0.5) * (x / F32(sigma)) ** 2).astype(F32)
    return (pdf / pdf.sum(dtype=F32)).astype(F32)


def gaussian_blur(img: np.ndarray, sigma: float) -> np.ndarray:
    """nexar_video_aug.py:251-255 -> tv:_functional_tensor.py:748-771 (reflect pad,
    depthwise conv with the outer-product kernel)."""
    ksize = int(sigma * 4) * 2 + 1
    k1 = gaussian_kernel1d(ksize, sigma)
    k2 = np.outer(k1, k1).astype(F32)
    p = ksize // 2
    padded = np.pad(img, ((0, 0), (p, p), (p, p)), mode="reflect")
    c, h, w = img.shape
    out = np.zeros_like(img, dtype=F32)
    for dy in range(ksize):
        for dx in range(ksize):
            out = out + padded[:, dy:dy + h, dx:dx + w] * k2[dy, dx]
    return out.astype(F32)


def posterize_float(img: np.ndarray, bits: int) -> np.ndarray:
    """nexar_video_aug.py:260-262 + tv:_functional_tensor.py:779-790."""
    b = (img * F32(255)).astype(F32).astype(np.uint8)  # C truncation, values in [0,255]
    mask = np.uint8((-int(2 ** (8 - bits))) & 0xFF)
    return ((b & mask).astype(F32) / F32(255.0)).astype(F32)


def solarize(img: np.ndarray, thr: float) -> np.ndarray:
    """tv:_functional_tensor.py:793-806: where(img >= thr, 1 - img, img)."""
    return np.where(img >= F32(thr), F32(1.0) - img, img).astype(F32)


# ----------------------------------------------------------------------------
# R0: RNG draw order; R5: per-frame chain
# ----------------------------------------------------------------------------
@dataclass
class AugConfig:
    """The kwargs create_video_transforms forwards to VideoAugmentation
    (nexar_video_aug.py:762-788); everything else is dropped by the reference."""
    brightness_range: Tuple[float, float] = (0.9, 1.1)
    contrast_range: Tuple[float, float] = (0.9, 1.1)
    saturation_range: Tuple[float, float] = (0.9, 1.1)
    hue_range: Tuple[float, float] = (-0.05, 0.05)
    rotation_range: Tuple[float, float] = (-5, 5)
    scale_range: Tuple[float, float] = (0.95, 1.05)
    shear_range: Tuple[float, float] = (-2, 2)
    translate_range: Tuple[float, float] = (0.0, 0.05)
    noise_level: float = 0.0
    blur_sigma: float = 0.0
    grayscale_prob: float = 0.0
    cutout_prob: float = 0.0
    color_inversion_prob: float = 0.0
    solarization_prob: float = 0.0
    posterization_prob: float = 0.0
    # VideoAugmentation defaults that the plural factory never overrides (:44-56)
    cutout_count_range: Tuple[int, int] = (1, 3)
    cutout_size_range: Tuple[float, float] = (0.1, 0.2)
    solarization_threshold: float = 0.5
    posterization_bits_range: Tuple[int, int] = (3, 6)
    aug_probability: float = 1.0


def sample_aug_params(cfg: AugConfig, h: int, w: int, rng=_random) -> Dict[str, Any]:
    """nexar_video_aug.py:97-182 — identical draw order from ``rng``."""
    p: Dict[str, Any] = {}
    if rng.random() > cfg.aug_probability:
        p["skip_augmentation"] = True
        return p
    p["brightness"] = rng.uniform(*cfg.brightness_range)
    p["contrast"] = rng.uniform(*cfg.contrast_range)
    p["saturation"] = rng.uniform(*cfg.saturation_range)
    p["hue"] = rng.uniform(*cfg.hue_range)
    p["rotation"] = rng.uniform(*cfg.rotation_range)
    p["scale"] = rng.uniform(*cfg.scale_range)
    p["shear"] = rng.uniform(*cfg.shear_range)
    p["translate_x"] = rng.uniform(-cfg.translate_range[1], cfg.translate_range[1]) * w
    p["translate_y"] = rng.uniform(-cfg.translate_range[1], cfg.translate_range[1]) * h
    p["apply_affine"] = (p["rotation"] != 0 or p["scale"] != 1 or p["shear"] != 0
                         or p["translate_x"] != 0 or p["translate_y"] != 0)
    p["apply_grayscale"] = rng.random() < cfg.grayscale_prob
    p["apply_noise"] = cfg.noise_level > 0
    p["apply_blur"] = cfg.blur_sigma > 0
    p["apply_cutout"] = rng.random() < cfg.cutout_prob
    if p["apply_cutout"]:
        p["cutout_count"] = rng.randint(*cfg.cutout_count_range)
        p["cutout_boxes"] = []
        for _ in range(p["cutout_count"]):
            sf = rng.uniform(*cfg.cutout_size_range)
            ch, cw = int(h * sf), int(w * sf)
            max_top, max_left = max(0, h - ch - 1), max(0, w - cw - 1)
            if max_top > 0 and max_left > 0:
                top = rng.randint(0, max_top)
                left = rng.randint(0, max_left)
                p["cutout_boxes"].append((top, left, ch, cw))
    p["apply_color_inversion"] = rng.random() < cfg.color_inversion_prob
    p["apply_solarization"] = rng.random() < cfg.solarization_prob
    p["apply_posterization"] = rng.random() < cfg.posterization_prob
    if p["apply_posterization"]:
        p["posterization_bits"] = rng.randint(*cfg.posterization_bits_range)
    return p


def augment_frame(frame: np.ndarray, p: Dict[str, Any], cfg: AugConfig, noise: Optional[np.ndarray] = None) -> np.ndarray:
    """nexar_video_aug.py:200-274, fixed op order.  ``noise`` (standard normal,
    same shape) replaces torch.randn_like so that runs are reproducible."""
    if p.get("skip_augmentation", False):
        return frame
    frame = adjust_brightness(frame, p["brightness"])
    frame = adjust_contrast(frame, p["contrast"])
    frame = adjust_saturation(frame, p["saturation"])
    frame = adjust_hue(frame, p["hue"])
    if p["apply_affine"]:
        m = inverse_affine_matrix(p["rotation"], [p["translate_x"], p["translate_y"]], p["scale"], [p["shear"], 0.0])
        frame = affine_bilinear_fill0(frame, m)
    if p["apply_grayscale"]:
        frame = np.broadcast_to(rgb_to_gray(frame)[None], frame.shape).astype(F32)
    if p["apply_noise"]:
        n = np.zeros_like(frame) if noise is None else noise
        frame = np.clip(frame + n * F32(cfg.noise_level), 0, 1).astype(F32)
    if p["apply_blur"]:
        frame = gaussian_blur(frame, cfg.blur_sigma)
    if p["apply_posterization"]:
        frame = posterize_float(frame, p["posterization_bits"])
    if p["apply_solarization"]:
        frame = solarize(frame, cfg.solarization_threshold)
    if p["apply_color_inversion"]:
        frame = (F32(1.0) - frame).astype(F32)
    if p["apply_cutout"]:
        frame = frame.copy()
        for top, left, ch, cw in p["cutout_boxes"]:
            frame[:, top:top + ch, left:left + cw] = 0
    return frame


def normalize(video: np.ndarray, mean: Sequence[float], std: Sequence[float]) -> np.ndarray:
    """nexar_video_aug.py:794-799 (true division)."""
    m = np.asarray(mean, F32).reshape(-1, 1, 1, 1)
    s = np.asarray(std, F32).reshape(-1, 1, 1, 1)
    return ((video - m) / s).astype(F32)


@dataclass
class TransformConfig:
    """create_video_transforms kwargs that have an effect (nexar_video_aug.py:636-696)."""
    mode: str = "train"
    crop_size: int = 224
    normalize: bool = True
    video_mean: Tuple[float, float, float] = (0.45, 0.45, 0.45)
    video_std: Tuple[float, float, float] = (0.225, 0.225, 0.225)
    horizontal_flip_prob: float = 0.5
    enable_custom_augmentation: bool = False
    aug: AugConfig = field(default_factory=AugConfig)


def sample_clip_params(cfg: TransformConfig, rng=_random) -> Dict[str, Any]:
    """Draws of one VideoTransform.forward call, in order: flip
    (nexar_video_aug.py:748), then the augmentation block (:290)."""
    out: Dict[str, Any] = {"flip": False, "aug": None}
    if cfg.mode == "train" and cfg.horizontal_flip_prob > 0:
        out["flip"] = rng.random() < cfg.horizontal_flip_prob
    if cfg.mode == "train" and cfg.enable_custom_augmentation:
        out["aug"] = sample_aug_params(cfg.aug, cfg.crop_size, cfg.crop_size, rng)
    return out


def apply_clip_transform(video: np.ndarray, cfg: TransformConfig, params: Dict[str, Any],
                         noise: Optional[np.ndarray] = None, stage: str = "final") -> np.ndarray:
    """VideoTransform.forward (nexar_video_aug.py:809-821) with the random
    decisions supplied in ``params``.  ``video`` is [C,T,H,W] uint8 or float.
    ``stage``: 'letterbox' | 'aug' (pre-normalisation, in [0,1]) | 'final'."""
    v = prologue(video)
    v = letterbox_resize(v, cfg.crop_size)
    if stage == "letterbox":
        return v
    if params["flip"]:
        v = v[..., ::-1].copy()
    if params["aug"] is not None:
        t = v.shape[1]
        frames = []
        for i in range(t):
            nz = None if noise is None else noise[:, i]
            frames.append(augment_frame(v[:, i], params["aug"], cfg.aug, nz))
        v = np.stack(frames, axis=1)
    if stage == "aug" or not cfg.normalize:
        return v
    return normalize(v, cfg.video_mean, cfg.video_std)


def clip_transform(video: np.ndarray, cfg: TransformConfig, rng=_random) -> np.ndarray:
    return apply_clip_transform(video, cfg, sample_clip_params(cfg, rng))


def resize_crop_transform(video: np.ndarray, size: int, cs: int, top: int, left: int) -> np.ndarray:
    """Dead-code variant R11 (nexar_video_aug.py:407-424, 464-482): short side to
    ``size`` then a cs x cs crop at (top, left); [C,T,H,W] in, values in [0,1]."""
    v = resize_short_side(prologue(video), size)
    return np.ascontiguousarray(v[:, :, top:top + cs, left:left + cs])
