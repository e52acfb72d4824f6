#!/usr/bin/env python
"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck, one tool per gpurun call): one tiny call of
every kernel family of the library."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import zero_filled_rss
from mri_acl_imagesegmentation_adsp_b200.utils import kspace as K

m = synth.knee_mask()
k = torch.from_numpy(synth.gaussian_kspace((2, 3, 640, 368), 1)).cuda()
ref = None
for sched in ("sequential", "pair", "coresident", "overlapped", "fused", "pipelined"):
    out, mean, std = zero_filled_rss(k, m, (320, 320), "instance", schedule=sched, chunk_slices=2)
    torch.cuda.synchronize()
    ref = out if ref is None else ref
    print(sched, float((out - ref).abs().max()))
os.environ["MRIACL_KC_RING"] = "1"
out, _, _ = zero_filled_rss(k, m, (320, 320), "instance", schedule="coresident")
print("coresident ring", float((out - ref).abs().max()))
os.environ.pop("MRIACL_KC_RING")
k5 = torch.from_numpy(synth.gaussian_kspace((2, 1, 2, 640, 368), 2)).cuda()           # (A, S, C, H, W): cooperative co-resident team
out, _, _ = zero_filled_rss(k5, m, (77, 200), None, average_axis=0, flip_rows=True, schedule="coresident")
kp = torch.from_numpy(synth.gaussian_kspace((2, 1, 2, 640, 451), 3)).cuda()
out, _, _ = zero_filled_rss(kp, synth.prostate_mask(), (320, 320), "instance", average_axis=0, flip_rows=True, pad=(94, 95))
out, _, _ = zero_filled_rss(kp, None, (75, 640), None, average_axis=0, pad=(94, 95))
k372 = torch.from_numpy(synth.gaussian_kspace((1, 2, 640, 372), 4)).cuda()
out, _, _ = zero_filled_rss(k372, synth.equispaced_mask(372, 4, 0.08), (320, 320), "instance")
x = synth.gaussian_kspace((2, 30, 23), 5)
K.ifft2c(x)
torch.cuda.synchronize()
print("done")
