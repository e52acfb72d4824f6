#!/usr/bin/env python
"""Quick probe of the overlapped schedule at full batch: correctness vs the sequential schedule and timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import zero_filled_rss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator(device="cuda").manual_seed(0)
k = torch.view_as_complex(torch.randn((B, 15, 640, 368, 2), device="cuda", generator=g))
m = synth.knee_mask()
ref, _, _ = zero_filled_rss(k, m, (320, 320), None, sequential=True)
torch.cuda.synchronize()
for name, kw in (("sequential", dict(sequential=True)), ("overlapped", dict())):
    for _ in range(3):
        out, _, _ = zero_filled_rss(k, m, (320, 320), None, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        out, _, _ = zero_filled_rss(k, m, (320, 320), None, **kw)
    b.record()
    torch.cuda.synchronize()
    print(name, "ms/step", a.elapsed_time(b) / 10, "equal", bool(torch.equal(out, ref)), flush=True)
