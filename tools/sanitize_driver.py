#!/usr/bin/env python
"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck, one tool per gpurun call): one tiny call of
every kernel family of the PRODUCT library (libmriacl_recon.so).  usage: compute-sanitizer --tool memcheck python tools/sanitize_driver.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.fastmri.sense import sens_combine
from mri_acl_imagesegmentation_adsp_b200.preprocess.mri_preprocess import MRIKneePreprocessor
from mri_acl_imagesegmentation_adsp_b200.prostate.grappa import Grappa
from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import stack_2p5d, zero_filled_rss
from mri_acl_imagesegmentation_adsp_b200.recon.pipeline import HostPipeline
from mri_acl_imagesegmentation_adsp_b200.utils import kspace as K

m = synth.knee_mask()
k_np = synth.gaussian_kspace((3, 3, 640, 368), 1)
k = torch.from_numpy(k_np).cuda()
out, mean, std = zero_filled_rss(k, m, (320, 320), "instance", chunk_slices=2)                       # 4-column column pass, ragged last group
idx = np.nonzero(m)[0]
pk, _, _ = zero_filled_rss(torch.from_numpy(np.ascontiguousarray(k_np[..., idx])).cuda(), m, (320, 320), "instance", packed=True)
print("packed == full:", bool(torch.equal(pk, out)))
oh = torch.empty((3, 320, 320)).pin_memory(); ms = torch.empty((3, 2)).pin_memory()
HostPipeline((3, 640, 368), (320, 320), "instance", 0.0, sub_batch=2, pack=True)(torch.from_numpy(k_np), m, oh, ms)
torch.cuda.synchronize()
k5 = torch.from_numpy(synth.gaussian_kspace((2, 1, 2, 640, 368), 2)).cuda()                         # (A, S, C, H, W)
zero_filled_rss(k5, m, (77, 200), None, average_axis=0, flip_rows=True)
kp = torch.from_numpy(synth.gaussian_kspace((2, 1, 2, 640, 451), 3)).cuda()
zero_filled_rss(kp, synth.prostate_mask(), (320, 320), "instance", average_axis=0, flip_rows=True, pad=(94, 95))
zero_filled_rss(kp, None, (75, 640), None, average_axis=0, pad=(94, 95))
for w in (372, 400, 320):
    kw = torch.from_numpy(synth.gaussian_kspace((1, 2, 640, w), 4)).cuda()
    zero_filled_rss(kw, synth.equispaced_mask(w, 4, 0.08), (320, 320), "instance")
K.ifft2c(synth.gaussian_kspace((2, 30, 23), 5))
pre = MRIKneePreprocessor(out_size=(40, 24))
img = torch.from_numpy(np.stack([synth.magnitude_image((37, 53), 7), synth.magnitude_image((37, 53), 8)])).cuda()
pre.clip_resize_zscore(img, (img > 0.3).to(torch.uint8))
pre.clip_resize_zscore(torch.from_numpy(synth.magnitude_image((640, 368), 9)[None]).cuda(), None)
stack_2p5d(np.random.default_rng(0).standard_normal((5, 1, 8, 12)).astype(np.float32), 3, imagenet_norm=True)
kg, calib = synth.grappa_case_inputs("small_r2")
g = Grappa(kg.copy(), (5, 5), 1)
g.apply_weights(kg.copy(), g.compute_weights(calib.copy()))
sens_combine(synth.gaussian_kspace((2, 3, 8, 6), 1), synth.gaussian_kspace((2, 3, 8, 6), 2), magnitude=True)
torch.cuda.synchronize()
print("done")
