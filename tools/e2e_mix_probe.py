#!/usr/bin/env python
"""End-to-end step time (pinned host k-space -> host images, 64 slices of configs[1]) per transfer pattern of
recon.pipeline.HostPipeline: direct, packed, mixed (every k-th sub-batch direct), for several sub-batch sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.recon.pipeline import HostPipeline
B, C, H, W = 64, 15, 640, 368
g = torch.Generator().manual_seed(0)
k = torch.view_as_complex(torch.randn((B, C, H, W, 2), generator=g)).pin_memory()
out = torch.empty((B, 320, 320), dtype=torch.float32).pin_memory()
ms = torch.empty((B, 2), dtype=torch.float32).pin_memory()
m = synth.knee_mask()
ref = None
for sub in (8, 16, 32):
    for pack, k_dir in [(False, 0), (True, 0), (True, 8), (True, 6), (True, 4), (True, 3), (True, 2)]:
        if k_dir > B // sub: continue
        pipe = HostPipeline((C, H, W), (320, 320), "instance", 0.0, sub_batch=sub, n_streams=2, pack=pack, direct_every=k_dir)
        for _ in range(2): pipe(k, m, out, ms)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 6
        for _ in range(n): pipe(k, m, out, ms)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        if ref is None: ref = out.clone()
        same = bool(torch.equal(ref, out))
        print(f"sub {sub:2d} pack {int(pack)} direct_every {k_dir}: {dt*1e3:7.2f} ms/step  {B/dt:7.0f} slices/s  h2d {pipe.h2d_bytes/1e6:7.0f} MB  bit-equal {same}", flush=True)
pipe = HostPipeline((C, H, W), (320, 320), "instance", 0.0, sub_batch=8, n_streams=2, pack="auto")
pipe(k, m, out, ms); torch.cuda.synchronize()
print("auto (sub 8):", pipe.calibration)
