#!/usr/bin/env python
"""Timing of the post-reconstruction epilogue (clip -> resize -> z-score) on 64 images, per kernel via CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mri_acl_imagesegmentation_adsp_b200.adapters import recon_cabi
from mri_acl_imagesegmentation_adsp_b200.preprocess.mri_preprocess import MRIKneePreprocessor
lib = recon_cabi.library()
B, H, W = 64, 320, 320
x = torch.rand((B, H, W), device="cuda") ** 3 * 5
m = (x > 0.5).to(torch.uint8)
pre = MRIKneePreprocessor(out_size=(320, 320))
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
out = torch.empty_like(x); lh = torch.empty((B, 2), device="cuda"); st = torch.empty((B, 6), device="cuda")
print("epilogue (one call)  %.3f ms" % timed(lambda: pre.clip_resize_zscore(x, m)))
print("percentile only      %.3f ms" % timed(lambda: lib.percentile_clip(x.data_ptr(), 0, lh.data_ptr(), B, H * W, 1.0, 99.5, 0)))
print("resize only          %.3f ms" % timed(lambda: lib.resize_bilinear(x.data_ptr(), out.data_ptr(), B, H, W, 320, 320, 0)))
print("zscore+preview only  %.3f ms" % timed(lambda: lib.zscore_preview(x.data_ptr(), m.data_ptr(), out.data_ptr(), 0, st.data_ptr(), B, H * W, 0)))
