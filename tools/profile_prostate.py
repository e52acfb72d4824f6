#!/usr/bin/env python
"""Minimal driver for ncu: two passes of the configs[2] prostate-shape volume through the fused 640-wide plan."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.prostate.t2 import t2_average_combine
g = torch.Generator(device="cuda").manual_seed(0)
k = torch.view_as_complex(torch.randn((3, 30, 16, 640, 451, 2), device="cuda", generator=g))
m = synth.prostate_mask()
for _ in range(2):
    img = t2_average_combine(k, (94, 95), (320, 320), m)
torch.cuda.synchronize()
print("ok", float(img.abs().mean()))
