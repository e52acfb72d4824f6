"""Slice / volume sharding of the input stage across the GPUs of one box.

The path has no exchange step: slices are independent, so every rank runs the
fused stage on its own shard and nothing crosses NVLink on the hot path.  The only collective is
the OPTIONAL gather of the finished ``(n, oh, ow)`` images (1.4 % of the bytes the kernels read),
done with ``torch.distributed`` (NCCL on GPUs; the same code runs under gloo in the CPU tests).

Two partition rules:
* ``slice_shard``  -- contiguous blocks of ``ceil(N / world)`` slices (configs[4] sweep);
* ``volume_shard`` -- whole volumes to one rank, volume ``i`` of the sorted name list to rank
  ``i % world``, the rule of the vendored ``VolumeSampler``
  (``ZIP!/DL_reconstruction/data/volume_sampler.py:63-90``), for consumers that need volume locality
  (prostate averages of one slice must stay together: the mean is taken after the RSS).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world_info() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def slice_shard(n_slices: int, world: Optional[int] = None, rank: Optional[int] = None) -> Tuple[int, int]:
    """[start, stop) of this rank's contiguous block; blocks are ``ceil(n/world)`` long, the tail
    ranks may be short or empty.  Every slice belongs to exactly one rank."""
    r, w = world_info()
    world = w if world is None else world
    rank = r if rank is None else rank
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} / world {world}")
    per = -(-n_slices // world) if n_slices > 0 else 0
    start = min(n_slices, rank * per)
    return start, min(n_slices, start + per)


def volume_shard(volume_names: Sequence[str], world: Optional[int] = None, rank: Optional[int] = None) -> List[str]:
    """Volumes of this rank: sorted unique names, every ``world``-th starting at ``rank``."""
    r, w = world_info()
    world = w if world is None else world
    rank = r if rank is None else rank
    names = sorted(set(str(v) for v in volume_names))
    return [names[i] for i in range(rank, len(names), world)]


def volume_shard_indices(example_volumes: Sequence[str], world: Optional[int] = None, rank: Optional[int] = None
                         ) -> List[int]:
    """Indices of the examples (one entry per slice, naming its volume) that land on this rank."""
    mine = set(volume_shard(example_volumes, world, rank))
    return [i for i, v in enumerate(example_volumes) if str(v) in mine]


def gather_slices(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather the per-rank result blocks of ``slice_shard`` back into ``(n_total, ...)`` on every
    rank.  Blocks are padded to the common ``ceil(n/world)`` length for ``all_gather_into_tensor`` and
    trimmed afterwards.  No-op without an initialised process group."""
    rank, world = world_info()
    if world == 1:
        return local
    per = -(-n_total // world)
    tail = local.shape[1:]
    padded = local
    if local.shape[0] < per:
        padded = torch.zeros((per,) + tuple(tail), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(tail), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous())
    return out[:n_total]


def recon_sharded(kspace_of_shard, n_total: int, mask=None, gather: bool = False, **kw):
    """Run the fused stage on this rank's shard; optionally gather the images on every rank.

    ``kspace_of_shard(start, stop)`` returns this rank's ``(n, C, H, W)`` k-space (device or host);
    keyword arguments go to ``zero_filled_rss``.  Returns ``(images, (start, stop))``."""
    from .recon.cartesian import zero_filled_rss
    start, stop = slice_shard(n_total)
    img, mean, std = zero_filled_rss(kspace_of_shard(start, stop), mask, **kw)
    if gather:
        img = gather_slices(img if isinstance(img, torch.Tensor) else torch.from_numpy(img), n_total)
    return img, (start, stop)
