#!/usr/bin/env python
"""bench.py -- k-space -> image slices/s of the fused input stage (BASELINE.json metric).

Workload (configs[1]): batches of 64 fifteen-coil 640x368 complex64 slices per GPU, 4x equispaced
mask + 29 ACS columns (114 of 368 kept), zero-filled centred iFFT per coil, RSS, crop 320x320,
instance normalise.  One "step" = one pass of the stage over one batch.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU)
  python bench.py --impl reference [--gpus N] ...                reference CPU path (oracle port)

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` goes through the
public host-buffer API (pinned host k-space in, host images out, copies inside the timed
region), `roofline` relates the step to the measured HBM bandwidth, `cpu_baseline` times the
oracle's numpy chain on this box's cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "kspace_to_image_slices_per_sec"
UNIT = "slices/s"
C, H, W = 15, 640, 368
CROP = (320, 320)
BYTES_PER_SLICE = C * H * W * 8 + W * 4 + CROP[0] * CROP[1] * 4     # 28 673 472
FALLBACK_HBM_GBS = 6650.0


def workload_config(batch: int) -> dict:
    return {"workload": "configs[1]: batched 15-coil 640x368 knee slices, 4x equispaced mask + 29 ACS (114/368 "
                        "columns), zero-filled ifft2c per coil + RSS + crop 320x320 + instance normalise",
            "batch_per_gpu": batch, "coils": C, "H": H, "W": W, "crop": list(CROP), "mask_columns": 114,
            "normalize": "instance", "algorithmic_bytes_per_slice": BYTES_PER_SLICE,
            "l2": f"inputs are {batch * C * H * W * 8 / 1e6:.0f} MB per step, larger than the 126 MB L2; no flush needed"}


def hbm_peak() -> tuple:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# CPU legs: the oracle's numpy chain (reference functions restated; kind "port")
# ---------------------------------------------------------------------------------------------
_W_SLICES = None
_W_MASK = None


def _cpu_worker_init(seed_base: int, per_worker: int):
    global _W_SLICES, _W_MASK
    import multiprocessing as mp
    from mri_acl_imagesegmentation_adsp_b200 import synth
    ident = mp.current_process()._identity
    wid = ident[0] if ident else 0
    _W_SLICES = [synth.gaussian_kspace((C, H, W), seed_base + 1000 * wid + i) for i in range(per_worker)]
    _W_MASK = synth.knee_mask()


def _cpu_worker_run(n: int) -> int:
    from oracle import recon_oracle as O
    for i in range(n):
        O.knee_chain_numpy(_W_SLICES[i % len(_W_SLICES)], _W_MASK, CROP, "instance")
    return n


class CpuPool:
    """All host cores, one process each, every worker holding a few pre-generated slices."""

    def __init__(self, cores: int):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("fork").Pool(cores, initializer=_cpu_worker_init, initargs=(12345, 2))
        self.pool.map(_cpu_worker_run, [1] * cores)      # warm every worker (imports, pocketfft plans)

    def run(self, slices_per_worker: int) -> tuple:
        t0 = time.perf_counter()
        done = sum(self.pool.map(_cpu_worker_run, [slices_per_worker] * self.cores))
        return done, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_baseline_leg() -> dict:
    cores = host_cores()
    pool = CpuPool(cores)
    # ~0.12 s per slice per core: 16 slices per core ~ 2 s wall, ~2 s x cores of CPU work (bounded to 10-30 s)
    per = max(2, min(16, int(round(20.0 / (0.125 * cores)))))
    done, dt = pool.run(per)
    pool.close()
    return {"value": done / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done} slices of the same workload ({per} per core, {cores} processes), oracle numpy chain "
                      f"(ifft2c -> complex_abs -> RSS -> center_crop_or_pad -> normalize_instance), {dt:.2f} s wall"}


def reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    pool = CpuPool(cores)
    per = 2
    for _ in range(args.warmup):
        pool.run(per)
    total, t = 0, 0.0
    for _ in range(args.steps):
        done, dt = pool.run(per)
        total += done
        t += dt
    pool.close()
    val = total / t
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.batch),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"each step = {per * cores} slices ({per} per core, {cores} processes) of the "
                                       "workload through the oracle's numpy restatement of the reference chain"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML.  It polls from before the warm-up
    (the first NVML queries take milliseconds and contend with kernel launches for the driver) and keeps
    only the samples that fall inside the timed regions (`active`)."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reason_bits, self.active = [], 0, False
        self.stop_flag = threading.Event()
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self._query()
            self.ok = True
        except Exception:
            self.ok = False

    def _query(self):
        mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
        try:
            bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            bits = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        return mhz, int(bits)

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            try:
                mhz, bits = self._query()
                if self.active:
                    self.samples.append(mhz)
                    self.reason_bits |= bits
            except Exception:
                pass
            time.sleep(self.period)

    def summary(self) -> dict:
        reasons = [n for b, n in self.REASONS.items() if self.reason_bits & b and n != "gpu_idle"]
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def ours(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_base = cpu_baseline_leg()          # before CUDA is initialised (fork-safe), rank 0 at N=1 only

    import torch
    import torch.distributed as dist
    from mri_acl_imagesegmentation_adsp_b200 import _device as D
    from mri_acl_imagesegmentation_adsp_b200 import synth
    from mri_acl_imagesegmentation_adsp_b200.adapters import recon_cabi as cabi
    from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import zero_filled_rss
    from mri_acl_imagesegmentation_adsp_b200.recon.pipeline import HostPipeline

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("MRIACL_NCCL_DEBUG", "WARN")   # no version banner on stdout: ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = cabi.library()
    B = args.batch
    mask = synth.knee_mask()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # synthetic k-space, resident in HBM before the timed region (seeded per rank)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    k = torch.view_as_complex(torch.randn((B, C, H, W, 2), device=dev, generator=g))
    chunk = args.chunk or D.DEFAULT_CHUNK_SLICES

    def step():
        return zero_filled_rss(k, mask, CROP, "instance", chunk_slices=chunk)

    sampler = ClockSampler(local)
    sampler.start()

    out = None
    for _ in range(max(args.warmup, 3)):
        out = step()       # same allocation pattern as the timed loop (the caching allocator settles here)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.launch_count()
    sampler.active = True
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    barrier()
    sampler.active = False
    launches = lib.launch_count() - n0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- per-kernel live timing (profiling flags run single phases of the same plan) ----
    peak, peak_src = hbm_peak()
    kern = {}
    if rank == 0:
        S, Wp = B, W
        ows = D.workspace(lib.recon_rss_workspace_bytes(min(chunk, S), 1, C, H, W, 0, Wp, CROP[0], CROP[1], mask, 0))
        o = torch.empty((S,) + CROP, dtype=torch.float32, device=dev)
        msd = torch.empty((S, 2), dtype=torch.float32, device=dev)

        def phase(flag):
            lib.recon_rss(k.data_ptr(), C * H * W, 0, mask, o.data_ptr(), msd.data_ptr(), S, 1, C, H, W, 0, Wp,
                          CROP[0], CROP[1], cabi.NORM_INSTANCE | flag, 0.0, ows.data_ptr(), ows.numel(), D.stream_ptr())

        reps = max(5, min(args.steps, 20))
        sampler.active = True
        names = (("colpass640", cabi.ONLY_COLPASS), ("rowpass_23x16", cabi.ONLY_ROWPASS),
                 ("normalize_instance", cabi.ONLY_NORM))
        # (a) in pipeline order: the three phases of the same plan back to back, events in between
        for _ in range(3):
            for _, flag in names:
                phase(flag)
        torch.cuda.synchronize()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(reps)]
        for r in range(reps):
            for i, (_, flag) in enumerate(names):
                evs[r][i].record()
                phase(flag)
            evs[r][3].record()
        torch.cuda.synchronize()
        for i, (name, _) in enumerate(names):
            kern[name] = float(np.mean([evs[r][i].elapsed_time(evs[r][i + 1]) for r in range(reps)]))
        kern["sum_in_pipeline"] = float(np.mean([evs[r][0].elapsed_time(evs[r][3]) for r in range(reps)]))
        # (b) each phase alone, repeated (warm caches for its own working set)
        for name, flag in names:
            for _ in range(3):
                phase(flag)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                phase(flag)
            b.record()
            torch.cuda.synchronize()
            kern[name + "_alone"] = a.elapsed_time(b) / reps
        sampler.active = False
    if world > 1:
        dist.barrier()

    # ---- end to end through the host-buffer API (pinned host in, host out) ----
    k_host = torch.empty((B, C, H, W), dtype=torch.complex64).pin_memory()
    k_host.copy_(k)
    out_host = torch.empty((B,) + CROP, dtype=torch.float32).pin_memory()
    ms_host = torch.empty((B, 2), dtype=torch.float32).pin_memory()
    pipe = HostPipeline((C, H, W), CROP, "instance", 0.0, sub_batch=args.sub_batch)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(2):
        pipe(k_host, mask, out_host, ms_host)
    barrier()
    sampler.active = True
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        pipe(k_host, mask, out_host, ms_host)
    e1.record()
    barrier()
    sampler.active = False
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
    e2e_check = float(out_host[0].abs().mean())
    sampler.stop_flag.set()

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    step_bytes = BYTES_PER_SLICE * B
    achieved = step_bytes / (ms_step * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic.get("step_dram_bytes") if traffic else None,
            "peak_source": peak_src,
            "what": "whole fused stage per step (colpass640 + rowpass + normalise launches): algorithmic "
                    f"{BYTES_PER_SLICE} B/slice x {B} slices / step time",
            "kernels_ms": kern,
            "dominant_kernel": {"name": "colpass640_ws_kernel",
                                "achieved": step_bytes / (kern["colpass640"] * 1e-3) / 1e9 if kern else None,
                                "frac": step_bytes / (kern["colpass640"] * 1e-3) / 1e9 / peak if kern else None,
                                "note": "reads all of k-space once; timed alone with CUDA events (ONLY_COLPASS)"}}
    if traffic:
        roof["traffic_detail"] = traffic
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(workload_config(B), chunk_slices=chunk, parallelism=f"slice-sharded x{world}, no collective"),
            "roofline": roof, "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "sub_batch": args.sub_batch,
                    "api": "recon.pipeline.HostPipeline (pinned host k-space -> host images)", "check_mean_abs": e2e_check},
            "clocks": sampler.summary(), "gpu_launches": int(launches)}
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=0, help="slices in flight per launch group (0 = library default)")
    ap.add_argument("--sub-batch", type=int, default=32, help="slices per host->device copy in the e2e pipeline")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it (one process per GPU)
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533")] + sys.argv
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
