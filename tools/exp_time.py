#!/usr/bin/env python
"""Step time of configs[1] under whatever MRIACL_* environment is set (experimental library A/B; no parity check)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import zero_filled_rss
g = torch.Generator(device="cuda").manual_seed(0)
k = torch.view_as_complex(torch.randn((64, 15, 640, 368, 2), device="cuda", generator=g))
m = synth.knee_mask()
f = lambda: zero_filled_rss(k, m, (320, 320), "instance")
for _ in range(5): f()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(30): f()
b.record(); torch.cuda.synchronize()
print(" ".join(f"{k_}={v}" for k_, v in sorted(os.environ.items()) if k_.startswith("MRIACL_") and k_ != "MRIACL_RECON_LIBRARY"), "->", round(a.elapsed_time(b) / 30, 4), "ms")
