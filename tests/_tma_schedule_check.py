"""Parity of the TMA-fed experimental schedules against the oracle; run as a subprocess by test_gpu_parity.py with
MRIACL_RECON_LIBRARY pointing at the experimental build and the MRIACL_* environment that selects the schedule
(the schedules read their knobs once per process).  Prints "OK <max rel-L2>" on success."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import zero_filled_rss
from oracle import recon_oracle as O

worst = 0.0
rng = np.random.default_rng(5)
masks = [synth.knee_mask(), synth.equispaced_mask(368, 8, 0.04, offset=3)]
w = np.zeros(368, np.float32); w[rng.choice(368, 70, replace=False)] = rng.uniform(0.5, 1.5, 70).astype(np.float32)
masks.append(w)                                            # ragged, weighted columns: bands with 0..8 sampled columns
for mi, m in enumerate(masks):
    k = synth.gaussian_kspace((5,) + synth.KNEE_SHAPE, 20 + mi)
    kd = torch.from_numpy(k).cuda()
    for chunk in (5, 2):
        img, mean, std = zero_filled_rss(kd, m, synth.CROP, "instance", chunk_slices=chunk)
        ref, rmean, rstd = O.knee_chain_numpy(k, m, synth.CROP, "instance")
        for s in range(5):
            worst = max(worst, O.rel_l2(img[s].cpu().numpy(), ref[s]))
        np.testing.assert_allclose(mean.cpu().numpy(), rmean, rtol=1e-5)
    raw, _, _ = zero_filled_rss(kd[:, 2:9], m, (640, 368), None)     # a coil subset: slice stride != C * H * W in the tensor map
    rref, _, _ = O.knee_chain_numpy(k[:, 2:9], m, (640, 368), None)
    worst = max(worst, O.rel_l2(raw.cpu().numpy(), rref))
# a 64-slice batch exercises the ring of T slots and the per-slice counters in both directions
g = torch.Generator(device="cuda").manual_seed(1)
kb = torch.view_as_complex(torch.randn((64, 15, 640, 368, 2), device="cuda", generator=g))
img, _, _ = zero_filled_rss(kb, masks[0], synth.CROP, "instance")
for _ in range(3):
    img2, _, _ = zero_filled_rss(kb, masks[0], synth.CROP, "instance")
assert torch.equal(img, img2)
for s in (0, 31, 63):
    ref, _, _ = O.knee_chain_numpy(kb[s:s + 1].cpu().numpy(), masks[0], synth.CROP, "instance")
    worst = max(worst, O.rel_l2(img[s].cpu().numpy(), ref[0]))
assert worst <= 1e-5, worst
print("OK", worst)
