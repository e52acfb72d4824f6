"""Host-buffer front end of the fused stage: host k-space in, host images out.

This is the end-to-end path a data loader would drive (``src/main.py:204-206`` hands one volume of
host arrays at a time to ``preprocess_records``): the batch is cut into sub-batches that alternate
between CUDA streams, so the host->device copy of sub-batch i+1 overlaps the kernels and the
device->host copy of sub-batch i.  All arithmetic is still the one C-ABI call per sub-batch.

Two ways of getting a sub-batch onto the device (``pack``):

* ``False``  one ``cudaMemcpyAsync`` of the whole sub-batch (needs pinned host k-space to be asynchronous).  Measured
  at 96 % of this box's plain pinned-copy ceiling (53.4 of 55.6 GB/s): the end-to-end rate IS the PCIe rate.
* ``True``   the undersampling mask multiplies the unsampled columns by exactly zero, so they never have to cross
  PCIe: host threads gather the sampled columns (114 of 368 at 4x) into a pinned staging buffer
  (``mriacl_pack_columns_host``), ONE ``cudaMemcpyAsync`` ships the 31 % that remain, and the column pass reads the
  packed layout (``MRIACL_PACKED_COLUMNS``).  The arithmetic sees the same values: images are bit-identical.  The
  host gather is bound by host memory bandwidth, so which mode wins depends on the box and on how many ranks share
  its memory system;
* mixed       packing is bound by host memory reads and leaves PCIe half idle, the direct copy is bound by PCIe and leaves
  the host idle: with ``direct_every = k`` every k-th sub-batch (the first of each group) goes across full width while the
  host threads gather the columns of the following ones, so both resources work at the same time;
* ``"auto"`` (default) times direct, packed and the mixed patterns on the first call and keeps the fastest.
"""
from __future__ import annotations

import time
from typing import Any, Optional, Tuple

import numpy as np
import torch

from .. import _device as D
from .cartesian import zero_filled_rss


class HostPipeline:
    def __init__(self, slice_shape: Tuple[int, int, int], crop: Tuple[int, int] = (320, 320),
                 normalize: Optional[str] = "instance", eps: float = 0.0, sub_batch: int = 8, n_streams: int = 2,
                 pack: Any = "auto", pack_threads: int = 0, collective_calibration: bool = False, direct_every: int = 0):
        dev = D.require_cuda()
        if pack not in (True, False, "auto"):
            raise ValueError("pack must be True, False or 'auto'")
        self.dev = dev
        self.slice_shape, self.crop, self.normalize, self.eps = tuple(slice_shape), tuple(crop), normalize, eps
        self.sub = int(sub_batch)
        self.pack, self.pack_threads = pack, int(pack_threads)
        # ranks of one box share the host memory system: with collective_calibration every rank of the default process
        # group times both modes at the same moment and all keep the mode with the smaller summed time (the first call
        # is then a collective: every rank must make it)
        self.collective = bool(collective_calibration)
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]
        self.stage = None            # device staging, full-width sub-batches  (allocated on first use)
        self.pstage = None           # (mask key, pinned packed buffers, device packed buffers, copy-done events)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.pack_s = self.wait_s = 0.0   # host seconds of the last call spent gathering columns / waiting for a staging buffer
        self.direct_every = int(direct_every) if pack is True else 0   # packed mode: every direct_every-th sub-batch is copied full width instead (0 = none)
        self.calibration = None      # {"direct_s": ..., "packed_s": ..., "mixed_s": {k: ...}} once "auto" has decided

    # ---- buffers -------------------------------------------------------------------------------------------------
    def _full_stage(self):
        if self.stage is None:
            self.stage = [torch.empty((self.sub,) + self.slice_shape, dtype=torch.complex64, device=self.dev)
                          for _ in self.streams]
        return self.stage

    def _packed_stage(self, m: np.ndarray):
        key = m.tobytes()
        if self.pstage is None or self.pstage[0] != key:
            c, h, _ = self.slice_shape
            n_act = int(np.count_nonzero(m))
            shape = (self.sub, c, h, max(1, n_act))
            pinned = [torch.empty(shape, dtype=torch.complex64).pin_memory() for _ in self.streams]
            device = [torch.empty(shape, dtype=torch.complex64, device=self.dev) for _ in self.streams]
            events = [None for _ in self.streams]          # recorded after the copy that reads pinned[b]; None = never used
            self.pstage = (key, pinned, device, events, n_act)
        return self.pstage

    # ---- one pass over the batch ---------------------------------------------------------------------------------
    def _run(self, kspace_host, m, out_host, mean_std_host, packed: bool, direct_every: int = 0):
        S = kspace_host.shape[0]
        c, h, w = self.slice_shape
        cur = torch.cuda.current_stream()
        self.h2d_bytes = self.d2h_bytes = 0
        for st in self.streams:
            st.wait_stream(cur)
        lib = D.lib()
        if packed:
            _, pinned, device, events, n_act = self._packed_stage(m)
            self.pack_s = self.wait_s = 0.0
        if not packed or direct_every:
            stage = self._full_stage()
        all_packed = packed
        for i, s0 in enumerate(range(0, S, self.sub)):
            n = min(self.sub, S - s0)
            b = i % len(self.streams)
            st = self.streams[b]
            packed = all_packed and not (direct_every and i % direct_every == 0)
            if packed:
                t0 = time.perf_counter()
                if events[b] is not None:
                    events[b].synchronize()          # the copy that last read this pinned buffer (this call or an earlier one) is done
                t1 = time.perf_counter()
                src = kspace_host[s0:s0 + n]
                lib.pack_columns_host(src.data_ptr(), pinned[b].data_ptr(), n * c * h, w, m, self.pack_threads)
                self.wait_s += t1 - t0
                self.pack_s += time.perf_counter() - t1
            with torch.cuda.stream(st):
                if packed:
                    device[b][:n].copy_(pinned[b][:n], non_blocking=True)
                    if events[b] is None:
                        events[b] = torch.cuda.Event()
                    events[b].record(st)
                    img, stats = zero_filled_rss(device[b][:n], m, self.crop, self.normalize, self.eps,
                                                 chunk_slices=n, packed=True, stats_2col=True)
                    self.h2d_bytes += n * c * h * n_act * 8
                else:
                    stage[b][:n].copy_(kspace_host[s0:s0 + n], non_blocking=True)
                    img, stats = zero_filled_rss(stage[b][:n], m, self.crop, self.normalize, self.eps, chunk_slices=n,
                                                 stats_2col=True)
                    self.h2d_bytes += n * c * h * w * 8
                out_host[s0:s0 + n].copy_(img, non_blocking=True)
                self.d2h_bytes += img.numel() * 4
                if mean_std_host is not None:
                    # ONE contiguous (n, 2) copy: a strided device -> host copy goes through a pageable temporary and
                    # blocks the host until the stream has drained, which serialises the sub-batches
                    mean_std_host[s0:s0 + n].copy_(stats, non_blocking=True)
                    self.d2h_bytes += 2 * n * 4
                    stats.record_stream(st)
                img.record_stream(st)
        for st in self.streams:
            cur.wait_stream(st)
        return out_host

    def _can_pack(self, m) -> bool:
        if m is None or self.slice_shape[1] != 640:       # packed k-space feeds the 640-row column pass only
            return False
        n_act = int(np.count_nonzero(m))
        return 0 < n_act < m.shape[0]

    def _calibrate(self, kspace_host, m, out_host, mean_std_host):
        """Time one pass of each mode on (at most) the first four sub-batches; keep the faster."""
        n_sub = (kspace_host.shape[0] + self.sub - 1) // self.sub
        n = min(kspace_host.shape[0], (8 if n_sub >= 8 else 4) * self.sub)
        n_sub = (n + self.sub - 1) // self.sub
        # candidates: (packed, direct_every); a mixed pattern needs at least one packed sub-batch per direct one
        cands = [(False, 0), (True, 0)] + [(True, k) for k in (8, 6, 4, 3, 2) if 2 <= k <= n_sub]
        times = []
        ms = None if mean_std_host is None else mean_std_host[:n]
        for packed, k in cands:
            self._run(kspace_host[:n], m, out_host[:n], ms, packed, k)   # warm
            torch.cuda.current_stream().synchronize()
            t0 = time.perf_counter()
            self._run(kspace_host[:n], m, out_host[:n], ms, packed, k)
            torch.cuda.current_stream().synchronize()
            times.append(time.perf_counter() - t0)
        self.calibration = {"direct_s": times[0], "packed_s": times[1], "mixed_s": {str(k): t for (_, k), t in zip(cands[2:], times[2:])},
                            "slices": n}
        if self.collective:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                t = torch.tensor(times, dtype=torch.float64, device=self.dev)
                dist.all_reduce(t)
                times = [float(x) for x in t]
                self.calibration.update(direct_s_all_ranks=times[0], packed_s_all_ranks=times[1],
                                        mixed_s_all_ranks={str(k): tt for (_, k), tt in zip(cands[2:], times[2:])})
        best = min(range(len(cands)), key=lambda i: times[i])
        self.pack, self.direct_every = cands[best]
        self.calibration["chosen"] = "direct" if not self.pack else ("packed" if not self.direct_every else f"mixed, every {self.direct_every}th sub-batch direct")
        if not self.pack:
            self.pstage = None                            # release the pinned staging buffers
        elif not self.direct_every:
            self.stage = None

    def __call__(self, kspace_host: torch.Tensor, mask: Any, out_host: torch.Tensor,
                 mean_std_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """kspace_host: CPU complex64 ``(S,C,H,W)`` (pinned for asynchronous copies when not packing);
        out_host: CPU float32 ``(S,oh,ow)`` (pinned).  Returns ``out_host`` after all streams have drained."""
        if kspace_host.device.type != "cpu" or out_host.device.type != "cpu":
            raise ValueError("HostPipeline takes host tensors; use zero_filled_rss for device-resident k-space")
        if tuple(kspace_host.shape[1:]) != self.slice_shape:
            raise ValueError(f"slice shape {tuple(kspace_host.shape[1:])} != {self.slice_shape}")
        if kspace_host.dtype != torch.complex64 or not kspace_host.is_contiguous():
            raise ValueError("kspace_host must be a contiguous complex64 tensor")
        m = D.host_mask(mask, self.slice_shape[2])
        packed = self.pack
        if packed in (True, "auto") and not self._can_pack(m):
            if packed is True:
                raise ValueError("pack=True needs a sampling mask with unsampled columns and 640-row k-space")
            packed = False
        if packed == "auto":
            self._calibrate(kspace_host, m, out_host, mean_std_host)
            packed = self.pack
        return self._run(kspace_host, m, out_host, mean_std_host, bool(packed), self.direct_every if packed else 0)


def zero_filled_rss_host(kspace_host: Any, mask: Any = None, crop: Tuple[int, int] = (320, 320),
                         normalize: Optional[str] = "instance", eps: float = 0.0, sub_batch: int = 8, pack: Any = "auto"):
    """Convenience wrapper: numpy / CPU-torch ``(S,C,H,W)`` in, numpy ``(S,oh,ow)`` + ``(S,2)`` mean/std out."""
    t = torch.from_numpy(np.ascontiguousarray(kspace_host, dtype=np.complex64)) if isinstance(kspace_host, np.ndarray) \
        else kspace_host.contiguous()
    if not t.is_pinned():
        t = t.pin_memory()
    out = torch.empty((t.shape[0],) + tuple(crop), dtype=torch.float32).pin_memory()
    ms = torch.empty((t.shape[0], 2), dtype=torch.float32).pin_memory()
    pipe = HostPipeline(tuple(t.shape[1:]), crop, normalize, eps, sub_batch, pack=pack)
    pipe(t, mask, out, ms)
    torch.cuda.current_stream().synchronize()
    return out.numpy(), ms.numpy()
