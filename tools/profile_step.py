#!/usr/bin/env python
"""Minimal driver for ncu: a few steps of the bench workload (device-resident synthetic k-space),
nothing else.  usage: python tools/profile_step.py [--batch 16] [--steps 3] [--chunk 16]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import zero_filled_rss

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--chunk", type=int, default=16)
args = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(0)
k = torch.view_as_complex(torch.randn((args.batch, 15, 640, 368, 2), device="cuda", generator=g))
m = synth.knee_mask()
for _ in range(args.steps):
    img, mean, std = zero_filled_rss(k, m, (320, 320), "instance", chunk_slices=args.chunk)
torch.cuda.synchronize()
print("ok", float(img.abs().mean()))
