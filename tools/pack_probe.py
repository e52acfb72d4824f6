#!/usr/bin/env python
"""Host gather throughput of mriacl_pack_columns_host on this box: threads x store mode (MRIACL_PACK_STREAM is read once
per process, so each store mode runs in its own process).  usage: python tools/pack_probe.py"""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from mri_acl_imagesegmentation_adsp_b200 import synth
    from mri_acl_imagesegmentation_adsp_b200.adapters import recon_cabi
    lib = recon_cabi.ReconLibrary(recon_cabi.DEFAULT_LIBRARY)
    S = 32
    k = torch.empty((S, 15, 640, 368), dtype=torch.complex64)
    k.view(torch.float32).normal_()
    m = synth.knee_mask()
    dst = torch.empty((S, 15, 640, 114), dtype=torch.complex64).pin_memory() if torch.cuda.is_available() else torch.empty((S, 15, 640, 114), dtype=torch.complex64)
    for nt in (1, 2, 4, 8, 16, 32, 0):
        lib.pack_columns_host(k.data_ptr(), dst.data_ptr(), S * 15 * 640, 368, m, nt)
        t0 = time.perf_counter()
        for _ in range(3):
            lib.pack_columns_host(k.data_ptr(), dst.data_ptr(), S * 15 * 640, 368, m, nt)
        dt = (time.perf_counter() - t0) / 3
        print(f"stream={os.environ.get('MRIACL_PACK_STREAM', '0')} threads={nt:2d}: {S / dt:7.0f} slices/s  {k.numel() * 8 / dt / 1e9:6.1f} GB/s of source", flush=True)
else:
    print("cores:", len(os.sched_getaffinity(0)))
    for mode in ("0", "1"):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, MRIACL_PACK_STREAM=mode))
