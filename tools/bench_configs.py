#!/usr/bin/env python
"""Side measurements for BASELINE.json configs[2] (prostate-shape T2 volume) and configs[3] (fused input stage feeding
the U-Net consumer).  Not the contract bench (bench.py measures configs[1]); prints one JSON line per config.
usage: python tools/bench_configs.py [--steps 10]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.models.unet_factory import build_unet
from mri_acl_imagesegmentation_adsp_b200.infer.segment import segment_kspace
from mri_acl_imagesegmentation_adsp_b200.prostate.t2 import t2_average_combine
from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import recon_to_unet_input

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=10)
args = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


# ---- configs[2]: (3 averages, 30 slices, 16 coils, 640 x 451), 8x equispaced + 18 ACS, pad (94, 95), flipud, mean, crop
A, S, C, RO, PE = 3, 30, 16, 640, 451
k = torch.view_as_complex(torch.randn((A, S, C, RO, PE, 2), device=dev, generator=g))
m = synth.equispaced_mask(PE, 8, 0.04) if hasattr(synth, "equispaced_mask") else None
ms = timed(lambda: t2_average_combine(k, (94, 95), (320, 320), m), max(2, args.steps // 2))
byts = A * S * C * RO * PE * 8 + S * 320 * 320 * 4
print(json.dumps({"config": "configs[2] prostate-shape T2 volume (3,30,16,640,451) c64, 8x mask, pad (94,95), flipud, mean over averages, crop 320",
                  "path": "fused 640-wide plan (colpass640 + rowpass640)", "ms_per_volume": ms, "output_slices_per_s": S / (ms * 1e-3),
                  "algorithmic_GBps": byts / (ms * 1e-3) / 1e9, "hbm_frac_of_6544": byts / (ms * 1e-3) / 1e9 / 6544.3}))
del k
torch.cuda.empty_cache()

# ---- configs[3]: batch-64 knee slices -> fused stage -> ResNet34 U-Net (random seeded weights, fp16 autocast)
B = 64
k = torch.view_as_complex(torch.randn((B, 15, 640, 368, 2), device=dev, generator=g))
mk = synth.knee_mask()
torch.manual_seed(0)
net = build_unet().to(dev).eval()
t_recon = timed(lambda: recon_to_unet_input(k, mk), args.steps)
t_all = timed(lambda: segment_kspace(net, k, mk, amp=True), args.steps)
print(json.dumps({"config": "configs[3] batch-64 15-coil knee slices -> fused input stage -> ResNet34 U-Net (seeded random weights, fp16 autocast) -> mask",
                  "recon_only_ms": t_recon, "recon_only_slices_per_s": B / (t_recon * 1e-3),
                  "recon_plus_unet_ms": t_all, "recon_plus_unet_slices_per_s": B / (t_all * 1e-3),
                  "input_stage_share": t_recon / t_all}))
