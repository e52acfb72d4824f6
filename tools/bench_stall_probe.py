#!/usr/bin/env python
"""Replicates bench.py's start-up sequence and prints per-step GPU times to locate one-off stalls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import zero_filled_rss

use_sampler = "--sampler" in sys.argv
torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(1234)
k = torch.view_as_complex(torch.randn((64, 15, 640, 368, 2), device="cuda", generator=g))
m = synth.knee_mask()
if use_sampler:
    s = bench.ClockSampler(0); s.start()
for _ in range(3):
    zero_filled_rss(k, m, (320, 320), "instance", chunk_slices=64)
torch.cuda.synchronize()
if use_sampler:
    s.active = True
K = 20
ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
t = [time.perf_counter()]
ev[0].record()
for i in range(K):
    out = zero_filled_rss(k, m, (320, 320), "instance", chunk_slices=64)
    ev[i + 1].record()
    t.append(time.perf_counter())
torch.cuda.synchronize()
print("sampler", use_sampler)
print("gpu ms per step:", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(K)])
print("cpu ms per step:", [round((t[i + 1] - t[i]) * 1e3, 3) for i in range(K)])
