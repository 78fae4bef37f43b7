"""Deterministic synthetic dashcam clips (integer arithmetic only, so numpy on
the host and torch on the device produce identical bytes).

Two distributions: ``noise`` — i.i.d. uniform bytes, the
worst case for antialiased-resize parity; ``dashcam`` — sky-to-road vertical
gradient, lane/horizon block structure drifting with time, and +-16 noise.
"""
from __future__ import annotations

import numpy as np

_M32 = 0xFFFFFFFF


def _hash32(idx, seed: int):
    """32-bit integer hash evaluated in 64-bit integers (works for numpy int64
    arrays and torch int64 tensors alike; only the low 32 bits are kept)."""
    x = (idx * 2654435761 + (seed * 40503 + 0x9E3779B9)) & _M32
    x = x ^ (x >> 15)
    x = (x * 2246822519) & _M32
    x = x ^ (x >> 13)
    x = (x * 3266489917) & _M32
    x = x ^ (x >> 16)
    return x


def _pattern(t, y, x, c, h: int, seed: int, kind: str, lin):
    noise = _hash32(lin, seed)
    if kind == "noise":
        return noise & 0xFF
    if kind != "dashcam":
        raise ValueError(f"unknown synthetic kind {kind!r}")
    sky = 215 - (y * 150) // max(h, 1)
    blocks = (((x + 3 * t) // 32 + y // 24) % 2) * 28
    lane = ((x + y // 2 + 5 * t) % 160 < 6) * 40
    tint = c * 9
    v = sky + blocks + lane - tint + (noise & 31) - 16
    return v.clip(0, 255) if hasattr(v, "clip") else v.clamp(0, 255)


def make_clip_np(t: int, h: int, w: int, seed: int = 0, kind: str = "noise") -> np.ndarray:
    """uint8 [T,H,W,3] on the host."""
    tt, yy, xx, cc = np.meshgrid(np.arange(t, dtype=np.int64), np.arange(h, dtype=np.int64),
                                 np.arange(w, dtype=np.int64), np.arange(3, dtype=np.int64), indexing="ij")
    lin = ((tt * h + yy) * w + xx) * 3 + cc
    return _pattern(tt, yy, xx, cc, h, seed, kind, lin).astype(np.uint8)


def make_clip_torch(t: int, h: int, w: int, seed: int = 0, kind: str = "noise", device="cuda"):
    """uint8 [T,H,W,3] generated on ``device`` (same bytes as make_clip_np)."""
    import torch

    ar = lambda n: torch.arange(n, dtype=torch.int64, device=device)
    tt = ar(t).view(t, 1, 1, 1)
    yy = ar(h).view(1, h, 1, 1)
    xx = ar(w).view(1, 1, w, 1)
    cc = ar(3).view(1, 1, 1, 3)
    lin = ((tt * h + yy) * w + xx) * 3 + cc
    return _pattern(tt, yy, xx, cc, h, seed, kind, lin).to(torch.uint8)


def rgb_to_nv12(clip):
    """Synthetic decoder surfaces from RGB frames: uint8 ``[...,H,W,3]`` -> uint8 ``[...,H*3/2,W]`` (Y plane, then the
    interleaved UV plane of the 2 x 2 block means), BT.601 limited range in the usual 8-bit integer form
    (Y = ((66R + 129G + 25B + 128) >> 8) + 16, U = ((-38R - 74G + 112B + 128) >> 8) + 128, V = ((112R - 94G - 18B + 128) >> 8) + 128).
    Works on numpy arrays and torch tensors (integer arithmetic only, identical bytes)."""
    is_np = isinstance(clip, np.ndarray)
    x = clip.astype(np.int32) if is_np else clip.int()
    r, g, b = x[..., 0], x[..., 1], x[..., 2]
    y = ((66 * r + 129 * g + 25 * b + 128) >> 8) + 16
    h, w = y.shape[-2], y.shape[-1]
    if h % 2 or w % 2:
        raise ValueError("nv12 needs an even height and width")

    def blocks(v):  # mean of each 2 x 2 block, rounded
        v = v.reshape(v.shape[:-2] + (h // 2, 2, w // 2, 2))
        return (v.sum(axis=-1).sum(axis=-2) + 2) >> 2 if is_np else (v.sum(dim=-1).sum(dim=-2) + 2) >> 2

    rm, gm, bm = blocks(r), blocks(g), blocks(b)
    u = ((-38 * rm - 74 * gm + 112 * bm + 128) >> 8) + 128
    v = ((112 * rm - 94 * gm - 18 * bm + 128) >> 8) + 128
    if is_np:
        uv = np.stack([u, v], axis=-1).reshape(u.shape[:-1] + (w,))
        return np.concatenate([y, uv], axis=-2).clip(0, 255).astype(np.uint8)
    import torch
    uv = torch.stack([u, v], dim=-1).reshape(u.shape[:-1] + (w,))
    return torch.cat([y, uv], dim=-2).clamp(0, 255).to(torch.uint8)
