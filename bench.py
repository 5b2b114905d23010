#!/usr/bin/env python
"""bench.py -- benchmark of the B200-native Savitzky-Golay engine.

Metric (BASELINE.json): Gsamples/s and fraction of the HBM roofline of the 1D batch filter.
Headline workload at every N (BASELINE.json configs[1], "c2"): 65,536 signals x 4,096 samples fp32,
half_window 16, poly_order 3, derivative 1, REFLECT boundary -- PER GPU (weak scaling: the batch of
independent signals is sharded with no collective, each rank filters its own 65,536 signals).

One step = one pass of the hot path over the whole per-GPU batch = one kernel launch through the C ABI
(savgol_apply_batch) on device-resident buffers (1 GiB in + 1 GiB out, far larger than the 126 MB L2, so no
explicit L2 flush is needed between steps).

The default invocation also runs the other BASELINE configs after the headline and attaches them to the
same JSON line under "configs": c1 (one 1M-sample signal), c3 (one 2^29-sample slice per GPU of a single
periodic signal; at N > 1 the n-sample halos are read from the ring neighbours' HBM over NVLink inside the
kernel, parity checked on EVERY rank at both seams), c5 (multichannel chunked stream), c4 (savgol2d, 256
images of 4096^2).  Each record carries value, ms_per_step, both rooflines (HBM and fp32 pipe), parity,
a `sustained` sub-record (seconds of back-to-back launches with clocks / power), `e2e` (the C-ABI call on
pinned HOST buffers, copies inside the timed region) and, at N = 1, `cpu_baseline` (the unmodified reference
on the box's host cores).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload all|c1|c2|c3|c4|c5]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gsamples/s & % of HBM roofline, 1D batch SG filter at 1/2/4/8 B200"
UNIT = "Gsamples/s"

WORKLOADS = {
    # name: dict(kind, shape..., filter params)
    "c2": dict(kind="batch", rows=65536, length=4096, n=16, m=3, d=1, dt=1.0, boundary="reflect",
               desc="1D batch: 65,536 signals x 4,096 samples per GPU, half_window=16 poly_order=3 derivative=1 reflect"),
    "c1": dict(kind="batch", rows=1, length=1_000_000, n=12, m=4, d=0, dt=1.0, boundary="polynomial",
               desc="one 1,000,000-sample signal, half_window=12 poly_order=4 derivative=0 polynomial"),
    "c3": dict(kind="long", length=1 << 29, n=32, m=4, d=2, dt=1.0, boundary="periodic",
               desc="2^29-sample slice per GPU of one periodic signal, half_window=32 poly_order=4 derivative=2, "
                    "n-sample halo exchange between ring neighbours"),
    "c5": dict(kind="stream", rows=1 << 20, length=1024, n=10, m=2, d=1, dt=1.0, boundary="polynomial",
               desc="multichannel stream: 1,048,576 channels x 1,024-sample chunks per GPU, half_window=10 poly_order=2 derivative=1"),
    "c4": dict(kind="2d", images=256, rows=4096, cols=4096, nx=7, ny=7, order=3, boundary="constant",
               desc="savgol2d: 256 images of 4096x4096 per GPU, 15x15 window, order 3, constant boundary"),
}
ORDER = ["c2", "c1", "c3", "c5", "c4"]


def config_of(name):
    """Static description of a workload -- identical in both arms (the driver compares them)."""
    wl = WORKLOADS[name]
    big = name != "c1"
    return {"workload": wl["desc"],
            "l2": "inputs larger than L2 (no flush needed)" if big else "L2 flushed between steps by a 256 MiB fill",
            "sharding": "contiguous slices of one signal per rank, n-sample halos from the ring neighbours" if wl["kind"] == "long"
            else "independent units per rank, no collective"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU with NVML while a region runs."""

    def __init__(self, index: int):
        self.samples = []  # (t, sm_mhz, reasons_bitmask, power_w)
        self.stop_flag = False
        self.thread = None
        self.nv = None
        self.h = None
        self.sm_max = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.perf_counter(), sm, rs, pw))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self.samples = []
        self.stop_flag = False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread:
            self.stop_flag = True
            self.thread.join()
            self.thread = None

    def summary(self, t0, t1):
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0}
        mask = 0
        for s in inside:
            mask |= s[2]
        return {"sm_mhz": statistics.median(s[1] for s in inside), "sm_max_mhz": self.sm_max,
                "reasons": sorted(v for k, v in names.items() if mask & k), "samples": len(inside),
                "power_w_max": round(max(s[3] for s in inside), 1)}


# ----------------------------------------------------------------------------------------------
def run_reference_arm(args, rank, world):
    """The reference's own CPU implementation on the box's host cores (oracle/_ref = the unmodified reference
    compiled from its sources; the oracle port only where that is absent).  Rank 0 alone works."""
    if rank != 0:
        return
    from bench_workloads import cpu_rate
    names = ORDER if args.workload == "all" else [args.workload]
    nthreads = os.cpu_count() or 1
    recs = {}
    for name in names:
        wl = WORKLOADS[name]
        # each step is one pass over a bounded sample so that steps + warmup stay within minutes
        budget = 1.0 if name == names[0] else 0.5
        rate, kind, sample, units = cpu_rate(wl, nthreads, reps=max(1, min(args.steps, 3)), budget=budget)
        recs[name] = {"value": round(rate, 4), "unit": UNIT, "cores": nthreads, "kind": kind, "sample": sample,
                      "ms_per_pass": round(units / rate / 1e6, 3)}
    head = names[0]
    r = recs[head]
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_pass"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(head),
        "cpu_baseline": r,
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if len(names) > 1:
        line["configs"] = {n: {"value": recs[n]["value"], "unit": UNIT, "config": config_of(n), "cpu_baseline": recs[n],
                               "e2e": {"value": recs[n]["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
                           for n in names[1:]}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
def record_of(name, res, world, peaks, peak_kind, traffic, steps, warmup):
    from bench_workloads import FP32_PEAK_TFMA
    units_per_step = res["units_per_step_per_rank"] * world
    ms = res["ms_per_step"]
    value = units_per_step / (ms * 1e-3) / 1e9
    kern_ms = res["kernel_ms"]
    alg_bytes = res["alg_bytes_per_launch"]
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    tfma = res["fp32_ops_per_unit"] * res["units_per_step_per_rank"] / (kern_ms * 1e-3) / 1e12
    rec = {
        "value": round(value, 3), "unit": UNIT, "steps": res.get("steps", steps), "warmup": warmup, "ms_per_step": round(ms, 5),
        "config": dict(config_of(name), **res.get("config", {})),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": traffic, "peak_kind": peak_kind,
                     "kernel": res["kernel"], "kernel_ms": round(kern_ms, 5), "alg_bytes_per_launch": alg_bytes,
                     "best_step_ms": round(res["best_step_ms"], 5) if res.get("best_step_ms") else None,
                     "frac_best_step": round(alg_bytes / (res["best_step_ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"], 4) if res.get("best_step_ms") else None,
                     "fp32": {"tfma_s": round(tfma, 2), "peak": FP32_PEAK_TFMA, "frac": round(tfma / FP32_PEAK_TFMA, 4),
                              "ops_per_unit": res["fp32_ops_per_unit"],
                              "peak_kind": "FFMA lane-operations/s, tools/microbench.cu on B200 at 1965 MHz"}},
        "clocks": res["clocks"],
        "gpu_launches": res["gpu_launches"], "tma_launches": res.get("tma_launches"),
        "parity": res.get("parity"),
    }
    for k in ("sustained", "strong_scaling", "e2e", "cpu_baseline"):
        if res.get(k):
            rec[k] = res[k]
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all"] + sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import savgol_b200 as sg
    from bench_workloads import Ctx, bind_near_gpu, run_1d_family, run_2d

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    numa = bind_near_gpu(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lib = sg.lib()
    try:  # DRAM bytes per launch of the dominant kernels, from the committed ncu --set full captures
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        traffic_tab = {}
    peaks, peak_kind = measured_peaks()
    ctx = Ctx(torch=torch, np=np, lib=lib, sg=sg, dev=dev, rank=rank, world=world, local=local, dist=dist, barrier=barrier,
              max_over_ranks=max_over_ranks, sampler=ClockSampler(local), sampler_cls=ClockSampler, peaks=peaks)

    names = ORDER if args.workload == "all" else [args.workload]
    recs = {}
    for i, name in enumerate(names):
        wl = WORKLOADS[name]
        head = i == 0
        steps = args.steps if head else max(3, min(args.steps, 10))
        sustain = 0.0 if args.no_sustained else (2.0 if name in ("c2", "c4") else 1.0)
        want_cpu = world == 1 and rank == 0 and not args.no_cpu
        fn = run_2d if wl["kind"] == "2d" else run_1d_family
        res = fn(wl, ctx, steps, args.warmup, not args.no_e2e, want_cpu, sustain)
        if rank == 0:
            recs[name] = record_of(name, res, world, peaks, peak_kind, traffic_tab.get(name, {}).get("bytes"), steps, args.warmup)
        del res
        torch.cuda.synchronize()
        torch.cuda.empty_cache()

    if rank == 0:
        head = recs[names[0]]
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic"}
        for k, v in head.items():
            if k not in ("value", "unit", "steps", "warmup", "ms_per_step"):
                line[k] = v
        if numa is not None:
            line["numa_node"] = numa
        if len(names) > 1:
            line["configs"] = {n: recs[n] for n in names[1:]}
            line["gpu_launches_all_configs"] = sum(r["gpu_launches"] for r in recs.values())
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
