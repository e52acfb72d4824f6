#!/usr/bin/env python
"""Host memory bandwidth of this box as seen by 16 torch threads: read-only reduction and copy."""
import time, torch
torch.set_num_threads(len(__import__("os").sched_getaffinity(0)))
a = torch.empty(1 << 28, dtype=torch.float32).normal_()      # 1 GiB
b = torch.empty_like(a)
for name, fn, byts in (("read (sum)", lambda: a.sum(), a.numel() * 4), ("copy", lambda: b.copy_(a), 2 * a.numel() * 4)):
    fn()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {byts / dt / 1e9:.1f} GB/s ({torch.get_num_threads()} threads)")
