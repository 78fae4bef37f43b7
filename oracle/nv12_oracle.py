"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

NV12 -> RGB oracle for the decoder-surface source format.  The reference never sees YUV:
decord.VideoReader / cv2.VideoCapture hand it RGB bytes (nexar_videos.py:360,422; inference.ipynb:65-94), produced by
FFmpeg's swscale, a third-party dependency that is absent from /root/reference and whose fixed-point tables are not
reproducible here.  The conversion is therefore a STATED formula, not a restated reference function (parity for this
one step is "unpinned"; everything after it is the pinned RGB path): ITU-R BT.601 limited range in the classic 8-bit
integer form, nearest-neighbour chroma,

    C = Y - 16, D = U - 128, E = V - 128
    R = clip((298 C + 409 E + 128) >> 8), G = clip((298 C - 100 D - 208 E + 128) >> 8), B = clip((298 C + 516 D + 128) >> 8)

(include/nexar_clip_transform.h states the same formula for NEXAR_SRC_NV12).
"""
import numpy as np


def nv12_to_rgb(nv12: np.ndarray) -> np.ndarray:
    """uint8 [...,H*3/2,W] (Y plane, then interleaved UV) -> uint8 [...,H,W,3]."""
    hh, w = nv12.shape[-2], nv12.shape[-1]
    h = hh * 2 // 3
    y = nv12[..., :h, :].astype(np.int32)
    uv = nv12[..., h:, :].astype(np.int32).reshape(nv12.shape[:-2] + (h // 2, w // 2, 2))
    u = np.repeat(np.repeat(uv[..., 0], 2, axis=-2), 2, axis=-1)
    v = np.repeat(np.repeat(uv[..., 1], 2, axis=-2), 2, axis=-1)
    c, d, e = y - 16, u - 128, v - 128
    r = (298 * c + 409 * e + 128) >> 8
    g = (298 * c - 100 * d - 208 * e + 128) >> 8
    b = (298 * c + 516 * d + 128) >> 8
    return np.stack([r, g, b], axis=-1).clip(0, 255).astype(np.uint8)
