import os, sys
sys.path.insert(0, "/root/repo")
import torch, numpy as np
from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import zero_filled_rss
g = torch.Generator(device="cuda").manual_seed(0)
for W in (400, 372, 368):
    k = torch.view_as_complex(torch.randn((16, 15, 640, W, 2), device="cuda", generator=g))
    m = synth.equispaced_mask(W, 4, 0.08)
    for _ in range(2): zero_filled_rss(k, m, (320, 320), "instance")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): zero_filled_rss(k, m, (320, 320), "instance")
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(W, "ms per 16 slices", ms, "slices/s", 16 / ms * 1e3)
