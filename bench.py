#!/usr/bin/env python
"""bench.py -- k-space -> image slices/s of the fused input stage (BASELINE.json metric).

Workload (configs[1]): batches of 64 fifteen-coil 640x368 complex64 slices per GPU, 4x equispaced
mask + 29 ACS columns (114 of 368 kept), zero-filled centred iFFT per coil, RSS, crop 320x320,
instance normalise.  One "step" = one pass of the stage over one batch.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU)
  python bench.py --impl reference [--gpus N] ...                reference CPU path (oracle port)

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` goes through the
public host-buffer API (pinned host k-space in, host images out, copies inside the timed
region), `roofline` relates the step to the measured HBM bandwidth (burst and a >= 2 s sustained loop),
`cpu_baseline` times the oracle's numpy chain on this box's cores, `extra_configs` carries the other
BASELINE.json configs (prostate volume, U-Net consumer, 10k-slice sharded sweep).
"""
from __future__ import annotations

import argparse
import glob
import hashlib
import json
import math
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
KSPACE_BYTES_PER_SLICE = C * H * W * 8 + W * 4                       # what the column pass alone must read
FALLBACK_HBM_GBS = 6650.0
SWEEP_SLICES = 10000                                                 # configs[4]


def workload_config(batch: int, chunk: int, world: int) -> dict:
    """The config dict both arms print (identical keys and values for the same command line)."""
    return {"workload": "configs[1]: batched 15-coil 640x368 knee slices, 4x equispaced mask + 29 ACS (114/368 "
                        "columns), zero-filled ifft2c per coil + RSS + crop 320x320 + instance normalise",
            "batch_per_gpu": batch, "coils": C, "H": H, "W": W, "crop": list(CROP), "mask_columns": 114,
            "normalize": "instance", "algorithmic_bytes_per_slice": BYTES_PER_SLICE,
            "l2": f"inputs are {batch * C * H * W * 8 / 1e6:.0f} MB per step, larger than the 126 MB L2; no flush needed",
            "chunk_slices": chunk, "parallelism": f"slice-sharded x{world}, no collective"}


_JSON_OUT = None


def emit(line: dict) -> None:
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def hbm_peak() -> tuple:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def csrc_sha() -> str:
    """Content hash of the kernel sources + C header: stamps profiles/traffic.json so that a capture taken from
    other kernels is never printed beside this build's timings (.git does not travel to the GPU box)."""
    h = hashlib.sha1()
    files = sorted(glob.glob(os.path.join(ROOT, "mri_acl_imagesegmentation_adsp_b200", "csrc", "*.cu")) +
                   glob.glob(os.path.join(ROOT, "mri_acl_imagesegmentation_adsp_b200", "csrc", "*.cuh")) +
                   glob.glob(os.path.join(ROOT, "mri_acl_imagesegmentation_adsp_b200", "csrc", "*.h")) +
                   glob.glob(os.path.join(ROOT, "include", "*.h")))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


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

    def run_total(self, n_slices: int) -> tuple:
        """Exactly n_slices one-slice tasks, taken by whichever worker is free."""
        t0 = time.perf_counter()
        done = sum(self.pool.imap_unordered(_cpu_worker_run, [1] * n_slices, chunksize=1))
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
    """The reference's CPU path (oracle numpy port: the reference is Python and /root/reference does not travel),
    all host cores, one step = one batch of `--batch` slices of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    pool = CpuPool(cores)
    B = args.batch
    for _ in range(max(1, min(args.warmup, 2))):
        pool.run_total(B)
    # bounded: a 64-slice step is ~0.4-0.6 s on 16-32 cores; cap the run at ~90 s of wall clock
    t_probe = pool.run_total(B)[1]
    steps_run = max(1, min(args.steps, int(90.0 / max(t_probe, 1e-3))))
    total, t = 0, 0.0
    for _ in range(steps_run):
        done, dt = pool.run_total(B)
        total += done
        t += dt
    pool.close()
    val = total / t
    chunk = args.chunk or int(os.environ.get("MRIACL_CHUNK_SLICES", "64"))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / steps_run, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(B, chunk, args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"each step = one batch of {B} slices of the workload spread over {cores} processes, "
                                       f"oracle's numpy restatement of the reference chain; {steps_run} of the {args.steps} "
                                       "requested steps were run (bounded to ~90 s)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "steps_run": steps_run, "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons of one GPU through NVML.  It polls from before the warm-up
    (the first NVML queries take milliseconds and contend with kernel launches for the driver) and keeps
    only the samples that fall inside the timed regions (`active`); `window()` opens a separately reported region."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reason_bits, self.active = [], 0, False
        self.win = None                  # {"mhz": [], "w": [], "bits": 0} while a window is open
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
        try:
            watts = self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        except Exception:
            watts = None
        return mhz, int(bits), watts

    def run(self):
        if not self.ok:
            return
        while not self.stop_flag.is_set():
            try:
                mhz, bits, watts = self._query()
                if self.active:
                    self.samples.append(mhz)
                    self.reason_bits |= bits
                w = self.win
                if w is not None:
                    w["mhz"].append(mhz)
                    w["bits"] |= bits
                    if watts is not None:
                        w["w"].append(watts)
            except Exception:
                pass
            time.sleep(self.period)

    def _names(self, bits: int) -> list:
        return [n for b, n in self.REASONS.items() if bits & b and n != "gpu_idle"]

    def window_open(self):
        self.win = {"mhz": [], "w": [], "bits": 0}

    def window_close(self) -> dict:
        w, self.win = self.win, None
        if not w or not w["mhz"]:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(w["mhz"])), "sm_mhz_min": float(min(w["mhz"])), "sm_max_mhz": self.max_mhz,
                "power_w_median": float(np.median(w["w"])) if w["w"] else None,
                "power_w_max": float(max(w["w"])) if w["w"] else None,
                "reasons": self._names(w["bits"]), "samples": len(w["mhz"])}

    def summary(self) -> dict:
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": self._names(self.reason_bits),
                "samples": len(self.samples)}


def bind_to_gpu_numa(local: int) -> dict:
    """Best effort: run this rank (and first-touch its pinned buffers) on the CPUs of its GPU's NUMA node when the
    box exposes more than one node.  Returns what was found for the JSON line."""
    info = {"nodes": None, "gpu_node": None, "bound": False}
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["nodes"] = len(nodes)
        import pynvml
        pynvml.nvmlInit()
        pci = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        pci = (pci.decode() if isinstance(pci, bytes) else pci).lower()
        dom, rest = pci.split(":", 1)
        pci = dom[-4:] + ":" + rest                     # NVML prints an 8-digit domain, sysfs a 4-digit one
        with open(f"/sys/bus/pci/devices/{pci}/numa_node") as f:
            node = int(f.read().strip())
        info["gpu_node"] = node
        if node >= 0 and len(nodes) > 1:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus = set()
                for part in f.read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            allowed = cpus & set(os.sched_getaffinity(0))
            if allowed:
                os.sched_setaffinity(0, allowed)
                info["bound"] = True
    except Exception as e:          # containers often hide sysfs: report and carry on
        info["error"] = type(e).__name__
    return info


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
    from mri_acl_imagesegmentation_adsp_b200 import sharding, synth
    from mri_acl_imagesegmentation_adsp_b200.adapters import recon_cabi as cabi
    from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import zero_filled_rss
    from mri_acl_imagesegmentation_adsp_b200.recon.pipeline import HostPipeline

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        # NCCL's own log (communicator size, transport, NVLS) goes to stderr so that stdout stays ONE JSON line
        os.environ.setdefault("NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
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

    def sum_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
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
    out_first = out[0].clone()

    # ---- sustained: the same step back to back for >= 2 s (clocks and power sampled over that window only) ----
    sustained = None
    if args.sustained_s > 0:
        n_sus = max(args.steps, int(math.ceil(args.sustained_s * 1e3 / ms_step)))
        barrier()
        sampler.window_open()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(n_sus):
            step()
        s1.record()
        barrier()
        clk = sampler.window_close()
        sus_ms = max_over_ranks(s0.elapsed_time(s1))
        sustained = {"seconds": sus_ms * 1e-3, "steps": n_sus, "ms_per_step": sus_ms / n_sus,
                     "value": world * B * n_sus / (sus_ms * 1e-3), "unit": UNIT, "clocks": clk}

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

    # ---- optional result gather over NCCL (never on the hot path): every rank checks the gathered block ----
    gather = None
    if world > 1:
        imgs = step()[0]
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        allimg = sharding.gather_slices(imgs, world * B)
        g1.record()
        torch.cuda.synchronize()
        own_ok = bool(torch.equal(allimg[rank * B:(rank + 1) * B], imgs))
        sums = torch.stack([allimg[r * B:(r + 1) * B].double().sum() for r in range(world)])
        mine = torch.zeros(world, dtype=torch.float64, device=dev)
        mine[rank] = imgs.double().sum()
        dist.all_reduce(mine, op=dist.ReduceOp.SUM)
        others_ok = bool(torch.equal(sums, mine))
        flag = torch.tensor([1.0 if (own_ok and others_ok) else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather = {"api": "sharding.gather_slices (all_gather_into_tensor, NCCL)", "bytes_per_rank": int(imgs.numel() * 4),
                  "ms": max_over_ranks(g0.elapsed_time(g1)), "own_block_bit_equal": own_ok, "all_blocks_checksum_equal": others_ok,
                  "ok_on_every_rank": bool(flag.item() == 1.0)}
        if not gather["ok_on_every_rank"]:
            raise SystemExit("gather_slices over NCCL returned a block that differs from the rank's own images")
        del allimg

    # ---- configs[4]: slice-sharded sweep of 10k slices (strong scaling: total work fixed, every rank cycles its pool) ----
    sweep = None
    if not args.no_extra:
        start, stop = sharding.slice_shard(SWEEP_SLICES, world, rank)
        n_mine = stop - start
        for _ in range(3):
            step()
        barrier()
        sampler.window_open()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        calls = 0
        for off in range(0, n_mine, B):
            m = min(B, n_mine - off)
            zero_filled_rss(k[:m], mask, CROP, "instance", chunk_slices=chunk)
            calls += 1
        w1.record()
        barrier()
        sweep_clk = sampler.window_close()
        mine_ms = w0.elapsed_time(w1)
        sweep_ms = max_over_ranks(mine_ms)
        sweep_min = -max_over_ranks(-mine_ms)
        sweep = {"workload": f"configs[4]: {SWEEP_SLICES} slices of the configs[1] shape, sharding.slice_shard -> contiguous "
                             f"blocks of ceil(n/world), every rank cycles its resident {B}-slice pool",
                 "scaling": "strong", "slices": SWEEP_SLICES, "slices_rank0": n_mine, "calls_rank0": calls,
                 "last_call_slices_rank0": (n_mine - 1) % B + 1 if n_mine else 0,
                 "ms_to_last_rank": sweep_ms, "ms_fastest_rank": sweep_min,
                 "value": SWEEP_SLICES / (sweep_ms * 1e-3), "unit": UNIT, "clocks": sweep_clk,
                 "hbm_frac_per_gpu": SWEEP_SLICES * BYTES_PER_SLICE / (sweep_ms * 1e-3) / 1e9 / peak / world}

    # ---- configs[2] and configs[3] (rank 0 at N=1) ----
    extra = {}
    if world == 1 and not args.no_extra:
        extra = extra_configs_leg(torch, dev, k, mask, peak)
    if sweep is not None:
        extra["configs[4]"] = sweep

    # ---- end to end through the host-buffer API (pinned host in, host out) ----
    k_host = torch.empty((B, C, H, W), dtype=torch.complex64).pin_memory()
    k_host.copy_(k)
    out_host = torch.empty((B,) + CROP, dtype=torch.float32).pin_memory()
    ms_host = torch.empty((B, 2), dtype=torch.float32).pin_memory()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    pack_arg = {"auto": "auto", "on": True, "off": False}[args.e2e_pack]

    def e2e_run(pipe, steps):
        for _ in range(2):
            pipe(k_host, mask, out_host, ms_host)      # warm-up (pack="auto" decides here, outside the timed region)
        barrier()
        sampler.active = True
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            pipe(k_host, mask, out_host, ms_host)
        e1.record()
        barrier()
        sampler.active = False
        ms = max_over_ranks(e0.elapsed_time(e1))
        return {"value": world * B * steps / (ms * 1e-3), "ms_per_step": ms / steps, "h2d_bytes_per_step": pipe.h2d_bytes,
                "d2h_bytes_per_step": pipe.d2h_bytes, "h2d_gbs": world * pipe.h2d_bytes * steps / (ms * 1e-3) / 1e9,
                "host_pack_ms_last_step": pipe.pack_s * 1e3, "host_wait_ms_last_step": pipe.wait_s * 1e3}

    pipe = HostPipeline((C, H, W), CROP, "instance", 0.0, sub_batch=args.sub_batch, n_streams=args.e2e_streams, pack=pack_arg,
                        pack_threads=args.pack_threads, collective_calibration=world > 1)
    head = e2e_run(pipe, e2e_steps)
    if world > 1 and pack_arg == "auto":
        # every rank decided on its own; report whether they agree (they share the host memory system)
        packed_ranks = int(sum_over_ranks(1.0 if pipe.pack else 0.0))
    else:
        packed_ranks = world if pipe.pack else 0
    head_mode = "direct" if not pipe.pack else ("packed" if not pipe.direct_every else "mixed")
    e2e_modes = {head_mode: head}
    if pack_arg == "auto":
        for name, pk in (("direct", False), ("packed", True)):      # the pure modes beside the one "auto" chose
            if name in e2e_modes:
                continue
            other = HostPipeline((C, H, W), CROP, "instance", 0.0, sub_batch=args.sub_batch, n_streams=args.e2e_streams,
                                 pack=pk, pack_threads=args.pack_threads)
            e2e_modes[name] = e2e_run(other, max(2, e2e_steps // 2))
            del other
        pipe(k_host, mask, out_host, ms_host)          # the parity check below reads the headline mode's output
        torch.cuda.synchronize()
    e2e_ms, e2e_value = head["ms_per_step"] * e2e_steps, head["value"]
    h2d_gbs = e2e_modes["direct"]["h2d_gbs"] if "direct" in e2e_modes else None
    # plain-copy ceiling: the same sub-batches, pinned host -> device, nothing else, all ranks at once
    sub = min(args.sub_batch, B)
    stage = torch.empty((sub, C, H, W), dtype=torch.complex64, device=dev)
    for _ in range(2):
        stage.copy_(k_host[:sub], non_blocking=True)
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_copies = 4 * ((B + sub - 1) // sub)
    c0.record()
    for i in range(n_copies):
        s0 = (i * sub) % max(1, B - sub + 1)
        stage.copy_(k_host[s0:s0 + sub], non_blocking=True)
    c1.record()
    barrier()
    copy_ms = max_over_ranks(c0.elapsed_time(c1))
    h2d_ceiling = world * n_copies * sub * C * H * W * 8 / (copy_ms * 1e-3) / 1e9
    del stage
    sampler.stop_flag.set()

    # e2e parity: first / middle / last slice of the host output against the oracle's numpy chain (rank 0)
    e2e_parity = None
    if rank == 0:
        from oracle import recon_oracle as O          # checker only: not on the measured path
        errs = {}
        for i in sorted({0, B // 2 - 1 if B > 1 else 0, B - 1}):
            ref, rmean, rstd = O.knee_chain_numpy(k_host[i].numpy()[None], mask, CROP, "instance")
            errs[str(i)] = float(O.rel_l2(out_host[i].numpy(), ref[0]))
        e2e_parity = {"rel_l2_vs_oracle": errs, "tolerance": 1e-5, "ok": all(v <= 1e-5 for v in errs.values()),
                      "device_resident_first_slice_bit_equal": bool(torch.equal(out_first.cpu(), out_host[0]))}
        if not e2e_parity["ok"]:
            raise SystemExit(f"e2e parity failed: {errs}")

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    step_bytes = BYTES_PER_SLICE * B
    achieved = step_bytes / (ms_step * 1e-3) / 1e9
    traffic, traffic_note = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
        if traffic.get("csrc_sha") != csrc_sha():
            traffic_note = (f"profiles/traffic.json was captured from kernel sources {traffic.get('csrc_sha')}, this build is "
                            f"{csrc_sha()}: not reported")
            traffic = None
    except Exception:
        pass
    own = KSPACE_BYTES_PER_SLICE * B
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic.get("step_dram_bytes") if traffic else None,
            "peak_source": peak_src,
            "what": "whole fused stage per step (colpass640 + rowpass + normalise launches): algorithmic "
                    f"{BYTES_PER_SLICE} B/slice x {B} slices / step time",
            "kernels_ms": kern,
            "dominant_kernel": {"name": "colpass640_ws_kernel", "algorithmic_bytes": own,
                                "achieved": own / (kern["colpass640"] * 1e-3) / 1e9 if kern else None,
                                "frac": own / (kern["colpass640"] * 1e-3) / 1e9 / peak if kern else None,
                                "note": "the kernel's own algorithmic bytes (k-space + mask, read once) / its own time "
                                        "(CUDA events, ONLY_COLPASS); the intermediate it also writes is not counted"}}
    if sustained is not None:
        sustained["achieved"] = BYTES_PER_SLICE * sustained["value"] / world / 1e9
        sustained["frac"] = sustained["achieved"] / peak
        roof["sustained"] = sustained
    if traffic:
        roof["traffic_detail"] = traffic
    if traffic_note:
        roof["traffic_note"] = traffic_note
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(B, chunk, world),
            "roofline": roof, "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "sub_batch": args.sub_batch, "streams": args.e2e_streams,
                    "api": "recon.pipeline.HostPipeline (pinned host k-space -> host images), one cudaMemcpyAsync per sub-batch; "
                           "packed = host threads gather the sampled columns first, so only they cross PCIe (bit-identical images); "
                           "mixed = every direct_every-th sub-batch goes across full width while the host gathers the others",
                    "mode": head_mode, "direct_every": pipe.direct_every, "pack_arg": args.e2e_pack, "ranks_packed": packed_ranks,
                    "pack_calibration": pipe.calibration, "modes": e2e_modes,
                    "host_kspace_gbs": e2e_value * C * H * W * 8 / 1e9,
                    "h2d_gbs": h2d_gbs, "h2d_ceiling_gbs": h2d_ceiling,
                    "h2d_frac_of_ceiling": h2d_gbs / h2d_ceiling if (h2d_ceiling and h2d_gbs) else None,
                    "h2d_ceiling_what": f"{n_copies} plain pinned->device copies of {sub}-slice sub-batches per rank, all "
                                        f"{world} ranks at once (aggregate GB/s)",
                    "numa": numa, "parity": e2e_parity},
            "clocks": sampler.summary(), "gpu_launches": int(launches), "csrc_sha": csrc_sha()}
    if gather is not None:
        line["gather"] = gather
    if extra:
        line["extra_configs"] = extra
    emit(line)


def extra_configs_leg(torch, dev, k, mask, peak) -> dict:
    """configs[2] (prostate-shape T2 volume through the 640-wide plan) and configs[3] (fused stage feeding the U-Net
    consumer), measured in the same run as the headline so that the driver's record carries them."""
    from mri_acl_imagesegmentation_adsp_b200 import synth
    from mri_acl_imagesegmentation_adsp_b200.infer.segment import segment_kspace
    from mri_acl_imagesegmentation_adsp_b200.models.unet_factory import build_unet
    from mri_acl_imagesegmentation_adsp_b200.prostate.t2 import t2_average_combine
    from mri_acl_imagesegmentation_adsp_b200.recon.cartesian import recon_to_unet_input, zero_filled_rss
    from oracle import recon_oracle as O              # checker only

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

    out = {}
    # configs[2]: (3 averages, 30 slices, 16 coils, 640 x 451), 8x equispaced + 18 ACS, pad (94, 95), flipud, mean, crop
    A, S, Cc, RO, PE = synth.PROSTATE_SHAPE
    g = torch.Generator(device=dev).manual_seed(77)
    kv = torch.view_as_complex(torch.randn((A, S, Cc, RO, PE, 2), device=dev, generator=g))
    pm = synth.prostate_mask()
    ms = timed(lambda: t2_average_combine(kv, synth.PROSTATE_PAD, CROP, pm), 10)
    img = t2_average_combine(kv, synth.PROSTATE_PAD, CROP, pm)
    sl = S // 2
    ref = O.prostate_chain(kv[:, sl:sl + 1].cpu().numpy(), pm, synth.PROSTATE_PAD, CROP)
    err = float(O.rel_l2(img[sl].cpu().numpy(), ref[0]))
    byts = A * S * Cc * RO * PE * 8 + S * CROP[0] * CROP[1] * 4
    n_act = int((pm != 0).sum())
    # at a 64-byte sampling stride every 32-byte sector holding a sampled column is fetched: sector-touch bytes
    sect = set()
    for w in np.nonzero(pm)[0]:
        sect.add((int(w) * 8) // 32)
    touch = A * S * Cc * RO * len(sect) * 32 + S * CROP[0] * CROP[1] * 4
    out["configs[2]"] = {"workload": "prostate-shape T2 volume (3,30,16,640,451) c64, 8x mask + 18 ACS, pad (94,95), flipud, "
                                     "mean over averages after RSS, crop 320x320; one t2_average_combine call",
                         "path": "colpass640_ws_kernel + rowpass640_kernel", "ms_per_volume": ms,
                         "output_slices_per_s": S / (ms * 1e-3), "sampled_columns": n_act,
                         "algorithmic_GBps": byts / (ms * 1e-3) / 1e9, "hbm_frac": byts / (ms * 1e-3) / 1e9 / peak,
                         "sector_touch_GBps": touch / (ms * 1e-3) / 1e9, "hbm_frac_sector_touch": touch / (ms * 1e-3) / 1e9 / peak,
                         "parity_rel_l2_slice": {"slice": sl, "value": err, "tolerance": 1e-5, "ok": err <= 1e-5}}
    del kv, img
    torch.cuda.empty_cache()
    # configs[3]: batch-64 knee slices -> fused stage -> ResNet34 U-Net (seeded random weights, fp16 autocast)
    torch.manual_seed(0)
    net = build_unet().to(dev).eval()
    B = k.shape[0]
    t_recon = timed(lambda: recon_to_unet_input(k, mask), 20)
    t_all = timed(lambda: segment_kspace(net, k, mask, amp=True), 5)
    x = recon_to_unet_input(k[:2], mask)
    ref, _, _ = O.knee_chain_numpy(k[:2].cpu().numpy(), mask, CROP, "instance")
    err3 = float(O.rel_l2(x[:, 0].cpu().numpy(), ref))
    out["configs[3]"] = {"workload": f"batch-{B} 15-coil knee slices -> fused input stage -> ResNet34 U-Net (seeded random "
                                     "weights, fp16 autocast, library convolutions) -> sigmoid > 0.5",
                         "recon_only_ms": t_recon, "recon_only_slices_per_s": B / (t_recon * 1e-3),
                         "recon_plus_unet_ms": t_all, "recon_plus_unet_slices_per_s": B / (t_all * 1e-3),
                         "input_stage_share": t_recon / t_all,
                         "segmentation_input_parity_rel_l2": {"value": err3, "tolerance": 1e-5, "ok": err3 <= 1e-5}}
    del net
    torch.cuda.empty_cache()
    # the per-slice steps after the reconstruction (percentile clip -> resize -> in-mask z-score + preview)
    from mri_acl_imagesegmentation_adsp_b200.preprocess.mri_preprocess import MRIKneePreprocessor
    pre = MRIKneePreprocessor(out_size=CROP)
    raw, _, _ = zero_filled_rss(k, mask, CROP, None)
    bm = (raw > 0.3 * raw.amax(dim=(1, 2), keepdim=True)).to(torch.uint8)
    t_post = timed(lambda: pre.clip_resize_zscore(raw, bm), 20)
    r = pre.clip_resize_zscore(raw[:1], bm[:1])
    wz, wp, wm, (wlo, whi) = O.post_chain(raw[0].cpu().numpy(), bm[0].cpu().numpy(), CROP, pre.clip_percentiles)
    out["post_steps"] = {"workload": f"{B} RSS images 320x320 (output of the fused stage) -> percentile clip (1, 99.5) -> bilinear "
                                     "resize 320x320 -> in-mask z-score + [0,1] preview, one C-ABI call",
                         "ms": t_post, "slices_per_s": B / (t_post * 1e-3),
                         "percentiles_bit_exact": bool(r["clip"][0, 0].item() == float(wlo) and r["clip"][0, 1].item() == float(whi)),
                         "mask_bit_exact": bool(np.array_equal(r["mask"][0].cpu().numpy(), wm)),
                         "img_z_max_abs_err": float(np.abs(r["img_z"][0].cpu().numpy() - wz).max())}
    # GRAPPA weight application at the prostate file shape, one average (30 slices) per launch
    from mri_acl_imagesegmentation_adsp_b200.prostate.grappa import Grappa
    S, Cc, RO, PE = 30, 16, 640, 451
    keep = np.zeros(PE, dtype=bool)
    keep[::2] = True
    keep[(PE - 24) // 2:(PE - 24) // 2 + 24] = True
    g = torch.Generator(device=dev).manual_seed(78)
    kg = torch.view_as_complex(torch.randn((S, Cc, RO, PE, 2), device=dev, generator=g))
    kg[..., torch.from_numpy(~keep).to(dev)] = 0
    gr = Grappa(np.transpose(kg[0].cpu().numpy(), (2, 0, 1)), kernel_size=(5, 5), coil_axis=1)
    rngw = np.random.default_rng(5)
    kv = gr.kernel_var_dict
    wd = {int(i): (0.05 * (rngw.standard_normal((Cc, int(kv["patches"][i].sum()))) +
                           1j * rngw.standard_normal((Cc, int(kv["patches"][i].sum()))))).astype(np.complex64) for i in kv["patch_indices"]}
    plan = gr._device_plan(dev, lanes_along_x=True)      # PE (x) is the contiguous axis of the file layout
    Wd = gr._pack_weights([wd] * S, plan)
    lib = cabi_library()

    def run_grappa():
        lib.grappa_apply(kg.data_ptr(), Cc * RO * PE, 1, PE, RO * PE, S, PE, RO, Cc, 5, 5, plan["hole_xy"].data_ptr(), plan["n_items"],
                         plan["item_geom"].data_ptr(), plan["item_first"].data_ptr(), plan["item_count"].data_ptr(),
                         plan["src_start"].data_ptr(), plan["src_off"].data_ptr(), plan["max_src"], plan["w_start"].data_ptr(),
                         Wd.data_ptr(), Wd.shape[1], 0 if torch.cuda.current_stream().cuda_stream == 0 else torch.cuda.current_stream().cuda_stream)

    t_gr = timed(run_grappa, 5)
    holes = int(plan["hole_xy"].numel())
    flops = 0.0
    for gi, gidx in enumerate(plan["geoms"]):
        n_s = int(kv["patches"][gidx][..., 0].sum())
        flops += 8.0 * Cc * n_s * Cc * len(kv["holes_x"][gidx])
    flops *= S
    sm_clock = 1.965e9
    fp32_peak = 148 * 128 * 2 * sm_clock
    out["grappa_apply"] = {"workload": f"GRAPPA 5x5 weight application, {S} slices x {Cc} coils x {RO} x {PE} (R=2 + 24 ACS lines), one launch, "
                                       "random weights (timing only; parity: tests/test_grappa_sense.py)",
                           "ms": t_gr, "holes_per_slice": holes, "geometries": len(plan["geoms"]), "gflop": flops / 1e9,
                           "tflops": flops / (t_gr * 1e-3) / 1e12, "fp32_peak_tflops": fp32_peak / 1e12,
                           "frac_of_fp32_peak": flops / (t_gr * 1e-3) / fp32_peak, "slices_per_s": S / (t_gr * 1e-3)}
    return out


def cabi_library():
    from mri_acl_imagesegmentation_adsp_b200.adapters import recon_cabi
    return recon_cabi.library()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=0, help="slices in flight per launch group (0 = library default)")
    ap.add_argument("--sub-batch", type=int, default=16, help="slices per host->device copy in the e2e pipeline")
    ap.add_argument("--e2e-streams", type=int, default=2)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--pack-threads", type=int, default=0, help="host threads of the column gather (0 = one per core)")
    ap.add_argument("--e2e-pack", default="auto", choices=["auto", "on", "off"],
                    help="host-side column packing before the host->device copy (HostPipeline pack=)")
    ap.add_argument("--sustained-s", type=float, default=2.5, help="length of the sustained device-resident loop (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_configs block (configs[2], [3], [4])")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it (one process per GPU)
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533")] + sys.argv
        raise SystemExit(subprocess.call(cmd))
    # stdout carries ONE JSON line: everything else any library prints there (NCCL's version banner, torchrun notes)
    # is sent to stderr by pointing fd 1 at fd 2 for the run; the line itself goes to the saved descriptor
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    global _JSON_OUT
    _JSON_OUT = os.fdopen(saved, "w")
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
