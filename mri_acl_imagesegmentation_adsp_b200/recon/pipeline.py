"""Host-buffer front end of the fused stage: pinned host k-space in, host images out.

This is the end-to-end path a data loader would drive (``src/main.py:204-206`` hands one volume of
host arrays at a time to ``preprocess_records``): the batch is cut into sub-batches that alternate
between two CUDA streams, so the host->device copy of sub-batch i+1 overlaps the kernels and the
device->host copy of sub-batch i.  All arithmetic is still the one C-ABI call per sub-batch.
"""
from __future__ import annotations

from typing import Any, Optional, Tuple

import numpy as np
import torch

from .. import _device as D
from .cartesian import zero_filled_rss


class HostPipeline:
    def __init__(self, slice_shape: Tuple[int, int, int], crop: Tuple[int, int] = (320, 320),
                 normalize: Optional[str] = "instance", eps: float = 0.0, sub_batch: int = 8, n_streams: int = 2):
        dev = D.require_cuda()
        self.slice_shape, self.crop, self.normalize, self.eps = tuple(slice_shape), tuple(crop), normalize, eps
        self.sub = int(sub_batch)
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        self.stage = [torch.empty((self.sub,) + self.slice_shape, dtype=torch.complex64, device=dev)
                      for _ in range(n_streams)]
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def __call__(self, kspace_host: torch.Tensor, mask: Any, out_host: torch.Tensor,
                 mean_std_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """kspace_host: CPU complex64 ``(S,C,H,W)`` (pinned for asynchronous copies);
        out_host: CPU float32 ``(S,oh,ow)`` (pinned).  Returns ``out_host`` after all streams have drained."""
        if kspace_host.device.type != "cpu" or out_host.device.type != "cpu":
            raise ValueError("HostPipeline takes host tensors; use zero_filled_rss for device-resident k-space")
        S = kspace_host.shape[0]
        if tuple(kspace_host.shape[1:]) != self.slice_shape:
            raise ValueError(f"slice shape {tuple(kspace_host.shape[1:])} != {self.slice_shape}")
        cur = torch.cuda.current_stream()
        self.h2d_bytes = self.d2h_bytes = 0
        for st in self.streams:
            st.wait_stream(cur)
        for i, s0 in enumerate(range(0, S, self.sub)):
            n = min(self.sub, S - s0)
            st, buf = self.streams[i % len(self.streams)], self.stage[i % len(self.streams)]
            with torch.cuda.stream(st):
                buf[:n].copy_(kspace_host[s0:s0 + n], non_blocking=True)
                img, mean, std = zero_filled_rss(buf[:n], mask, self.crop, self.normalize, self.eps, chunk_slices=n)
                out_host[s0:s0 + n].copy_(img, non_blocking=True)
                self.h2d_bytes += n * int(np.prod(self.slice_shape)) * 8
                self.d2h_bytes += img.numel() * 4
                if mean_std_host is not None:
                    mean_std_host[s0:s0 + n, 0].copy_(mean, non_blocking=True)
                    mean_std_host[s0:s0 + n, 1].copy_(std, non_blocking=True)
                    self.d2h_bytes += 2 * n * 4
                img.record_stream(st)
        for st in self.streams:
            cur.wait_stream(st)
        return out_host


def zero_filled_rss_host(kspace_host: Any, mask: Any = None, crop: Tuple[int, int] = (320, 320),
                         normalize: Optional[str] = "instance", eps: float = 0.0, sub_batch: int = 8):
    """Convenience wrapper: numpy / CPU-torch ``(S,C,H,W)`` in, numpy ``(S,oh,ow)`` + ``(S,2)`` mean/std out."""
    t = torch.from_numpy(np.ascontiguousarray(kspace_host, dtype=np.complex64)) if isinstance(kspace_host, np.ndarray) \
        else kspace_host.contiguous()
    if not t.is_pinned():
        t = t.pin_memory()
    out = torch.empty((t.shape[0],) + tuple(crop), dtype=torch.float32).pin_memory()
    ms = torch.empty((t.shape[0], 2), dtype=torch.float32).pin_memory()
    pipe = HostPipeline(tuple(t.shape[1:]), crop, normalize, eps, sub_batch)
    pipe(t, mask, out, ms)
    torch.cuda.current_stream().synchronize()
    return out.numpy(), ms.numpy()
