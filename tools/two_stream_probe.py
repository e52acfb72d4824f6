#!/usr/bin/env python
"""Throughput of configs[1] when independent batches alternate between N CUDA streams (the tail of one batch's row pass and
normalise launch overlaps the next batch's column pass) against the single-stream step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import zero_filled_rss
g = torch.Generator(device="cuda").manual_seed(0)
ks = [torch.view_as_complex(torch.randn((64, 15, 640, 368, 2), device="cuda", generator=g)) for _ in range(2)]
m = synth.knee_mask()
for n_streams in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    def run(steps):
        cur = torch.cuda.current_stream()
        for st in streams: st.wait_stream(cur)
        for i in range(steps):
            with torch.cuda.stream(streams[i % n_streams]):
                zero_filled_rss(ks[i % 2], m, (320, 320), "instance")
        for st in streams: cur.wait_stream(st)
    run(6); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(60); b.record(); torch.cuda.synchronize()
    print(f"{n_streams} stream(s): {a.elapsed_time(b) / 60:.4f} ms per 64-slice batch")
