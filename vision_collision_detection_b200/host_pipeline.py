"""Host-buffer entry point: pinned uint8 clips in host memory -> H2D -> fused transform -> D2H.

This is the path a DataLoader-side caller exercises (decoded frames live in host memory,
nexar_videos.py:422-438) and what bench.py reports as ``e2e``.  The batch is cut into chunks
that round-robin over a few CUDA streams so that the H2D copy of one chunk overlaps the
kernels and the D2H copy of its neighbours (PCIe is full duplex).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch

from .engine import ClipTransformEngine
from .video_aug import GpuVideoTransform


class HostClipPipeline:
    def __init__(self, transform: GpuVideoTransform, n_clips: int, frames: int, height: int, width: int,
                 device=None, clips_per_chunk: int = 4, n_streams: int = 3, out_dtype: Optional[torch.dtype] = None):
        self.tf = transform
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.shape = (n_clips, frames, height, width, 3)
        self.chunk = max(1, min(clips_per_chunk, n_clips))
        self.out_dtype = out_dtype or transform.out_dtype
        cs = transform.crop_size
        self.streams = [torch.cuda.Stream(self.device) for _ in range(n_streams)]
        self.engines = [ClipTransformEngine(self.device) for _ in range(n_streams)]
        self.dev_in = [torch.empty((self.chunk, frames, height, width, 3), dtype=torch.uint8, device=self.device)
                       for _ in range(n_streams)]
        self.dev_out = [torch.empty((self.chunk, 3, frames, cs, cs), dtype=self.out_dtype, device=self.device)
                        for _ in range(n_streams)]
        self.host_out = torch.empty((n_clips, 3, frames, cs, cs), dtype=self.out_dtype).pin_memory()

    def pinned_input(self) -> torch.Tensor:
        return torch.empty(self.shape, dtype=torch.uint8).pin_memory()

    @torch.no_grad()
    def run(self, host_clips: torch.Tensor, params: Optional[List[Dict[str, Any]]] = None) -> torch.Tensor:
        """host_clips: pinned uint8 [B,T,H,W,3].  Returns the pinned [B,3,T,cs,cs] result (valid on return)."""
        if tuple(host_clips.shape) != self.shape or host_clips.dtype != torch.uint8:
            raise ValueError(f"expected uint8 {self.shape}")
        n = self.shape[0]
        if params is None:
            params = self.tf.sample_params(n, self.shape[2], self.shape[3])
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        for k, lo in enumerate(range(0, n, self.chunk)):
            hi = min(n, lo + self.chunk)
            i = k % len(self.streams)
            with torch.cuda.stream(self.streams[i]):
                din = self.dev_in[i][: hi - lo]
                din.copy_(host_clips[lo:hi], non_blocking=True)
                dout = self.dev_out[i][: hi - lo]
                self.tf.forward_batch(din, params=params[lo:hi], out=dout, engine=self.engines[i])
                self.host_out[lo:hi].copy_(dout, non_blocking=True)
        for s in self.streams:
            cur.wait_stream(s)
        cur.synchronize()
        return self.host_out
