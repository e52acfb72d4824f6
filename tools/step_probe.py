#!/usr/bin/env python
"""Where does a bench step spend its time: CPU enqueue vs GPU execution, with / without NVML polling."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import zero_filled_rss

g = torch.Generator(device="cuda").manual_seed(0)
k = torch.view_as_complex(torch.randn((64, 15, 640, 368, 2), device="cuda", generator=g))
m = synth.knee_mask()

def measure(tag, steps=50):
    for _ in range(5):
        zero_filled_rss(k, m, (320, 320), "instance")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(steps):
        zero_filled_rss(k, m, (320, 320), "instance")
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{tag:28s} gpu {a.elapsed_time(b)/steps:.4f} ms/step  cpu-enqueue {(t1-t0)/steps*1e3:.4f} ms/step  wall {(t2-t0)/steps*1e3:.4f}", flush=True)

measure("plain")
measure("plain again")
stop = threading.Event()
def poll(period):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    while not stop.is_set():
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        time.sleep(period)
for period in (0.1, 0.02):
    stop.clear()
    th = threading.Thread(target=poll, args=(period,), daemon=True); th.start()
    time.sleep(0.3)
    measure(f"nvml poll every {period}s")
    measure(f"nvml poll every {period}s (200 steps)", 200)
    stop.set(); th.join()
measure("plain after")
