"""Host-side logic of the N>1 path on CPU: shard rules and the optional result gather under a
world_size-2 gloo process group (the GPU run uses the same code over NCCL)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mri_acl_imagesegmentation_adsp_b200 import sharding


def test_slice_shard_partitions_exactly():
    for n in [0, 1, 7, 64, 10000, 10001]:
        for world in [1, 2, 3, 4, 8]:
            covered = []
            for r in range(world):
                a, b = sharding.slice_shard(n, world, r)
                assert 0 <= a <= b <= n
                covered += list(range(a, b))
            assert covered == list(range(n))
    assert sharding.slice_shard(10000, 8, 0) == (0, 1250) and sharding.slice_shard(10000, 8, 7) == (8750, 10000)
    with pytest.raises(ValueError):
        sharding.slice_shard(10, 2, 2)


def test_volume_shard_rule():
    examples = [f"vol{v:02d}" for v in [3, 3, 3, 1, 1, 2, 0, 0, 0, 0, 4]]
    w = 2
    r0 = sharding.volume_shard(examples, w, 0)
    r1 = sharding.volume_shard(examples, w, 1)
    assert r0 == ["vol00", "vol02", "vol04"] and r1 == ["vol01", "vol03"]
    i0 = sharding.volume_shard_indices(examples, w, 0)
    i1 = sharding.volume_shard_indices(examples, w, 1)
    assert sorted(i0 + i1) == list(range(len(examples)))
    # all slices of a volume stay on one rank
    assert {examples[i] for i in i0}.isdisjoint({examples[i] for i in i1})


def _worker(rank: int, world: int, port: int, n_total: int, out_dir: str):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = sharding.slice_shard(n_total)
        # stand-in for the per-rank stage output: slice s is filled with the value s
        local = torch.arange(a, b, dtype=torch.float32).reshape(-1, 1, 1).expand(-1, 4, 5).contiguous()
        full = sharding.gather_slices(local, n_total)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 8])
def test_gather_world2_gloo(tmp_path, n_total):
    world, port = 2, 29641 + n_total
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    want = np.broadcast_to(np.arange(n_total, dtype=np.float32).reshape(-1, 1, 1), (n_total, 4, 5))
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"rank{r}.npy"), want)
