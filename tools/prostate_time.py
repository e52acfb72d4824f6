#!/usr/bin/env python
"""configs[2] volume time (CUDA events) + per-kernel split through the ONLY_* flags are not needed here: one number."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.prostate.t2 import t2_average_combine
g = torch.Generator(device="cuda").manual_seed(0)
k = torch.view_as_complex(torch.randn((3, 30, 16, 640, 451, 2), device="cuda", generator=g))
m = synth.prostate_mask()
f = lambda: t2_average_combine(k, (94, 95), (320, 320), m)
for _ in range(5): f()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): f()
b.record(); torch.cuda.synchronize()
print("configs[2] volume:", round(a.elapsed_time(b) / 20, 4), "ms")
