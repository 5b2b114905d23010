#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Savitzky-Golay engine.

Metric (BASELINE.json): Gsamples/s and fraction of the HBM roofline of the 1D batch filter.
Workload at every N (BASELINE.json configs[1], "c2"): 65,536 signals x 4,096 samples fp32,
half_window 16, poly_order 3, derivative 1, REFLECT boundary -- PER GPU (weak scaling: the batch
of independent signals is sharded with no collective, each rank filters its own 65,536 signals).

One step = one pass of the hot path over the whole per-GPU batch = one launch of sg1d_kernel
through the C ABI (savgol_apply_batch) on device-resident buffers (1 GiB in + 1 GiB out, far
larger than the 126 MB L2, so no explicit L2 flush is needed between steps).

JSON line keys: see the task contract.  `value` = device-resident throughput (CUDA events, max
over ranks); `e2e` = the same call with pinned HOST buffers, H2D + D2H inside the timed region;
`roofline` = algorithmic bytes (8 B/sample) / kernel time vs the measured HBM peak;
`cpu_baseline` = the unmodified reference C code (oracle/_ref) on all host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c1|c3|c5|c4]
"""
from __future__ import annotations

import argparse
import ctypes as C
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
def make_batch_numpy(rows, length, seed, np):
    """Seeded synthetic batch: N(0,1) noise + a per-signal sinusoid."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((rows, length), dtype=np.float32)
    t = np.arange(length, dtype=np.float32)
    amp = rng.uniform(0.5, 2.0, (rows, 1)).astype(np.float32)
    frq = rng.uniform(0.002, 0.05, (rows, 1)).astype(np.float32)
    x += amp * np.sin(frq * t[None, :])
    return x


def cpu_reference_rate(wl, nthreads, reps=2, rows_cap=None):
    """Times the unmodified reference (oracle/_ref) -- or the oracle port when it is absent -- over
    independent signals on `nthreads` host threads.  Returns (Gsamples/s, kind, sample description)."""
    import numpy as np
    from oracle import oracle as O
    rows = wl["rows"] if rows_cap is None else min(wl["rows"], rows_cap)
    L = wl["length"]
    x = make_batch_numpy(rows, L, 1, np)
    y = np.empty_like(x)
    lib = O.lib()
    if O.have_ref():
        R = O.ref()
        cfg = O.make_config(wl["n"], wl["m"], wl["d"], wl["dt"], wl["boundary"])
        f = R.savgol_create(C.byref(cfg))
        run = lambda: lib.sgh_apply_rows(O.fnptr(R, "savgol_apply"), C.cast(f, C.c_void_p), O._fp(x), O._fp(y),
                                         rows, L, L, L, nthreads)
        kind = "reference"
    else:
        of = O.Filter1D(wl["n"], wl["m"], wl["d"], wl["dt"], wl["boundary"])
        import concurrent.futures as cf
        pool = cf.ThreadPoolExecutor(nthreads)
        step = (rows + nthreads - 1) // nthreads

        def part(i):
            a, b = i * step, min(rows, (i + 1) * step)
            if a < b:
                lib.sgo_apply_batch(of.n, O._fp(of.center), O._fp(of.edge), of.dt_inv, of.mode,
                                    x[a:b].ctypes.data_as(O.f32p), y[a:b].ctypes.data_as(O.f32p), b - a, L, L, L)
        run = lambda: list(pool.map(part, range(nthreads)))
        kind = "port"
    run()  # warm (page faults, thread start)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        run()
        best = min(best, time.perf_counter() - t0)
    rate = rows * L / best / 1e9
    sample = f"{rows}x{L} signals ({'full per-GPU workload' if rows == wl['rows'] else 'subset'}), best of {reps}, " \
             f"gcc -O2 -ffp-contract=off, {nthreads} threads over independent signals"
    return rate, kind, sample, (x, y)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    if wl["kind"] != "batch":
        print(json.dumps({"impl": "reference", "unavailable": "reference arm implemented for the 1D batch workloads"}))
        return
    nthreads = os.cpu_count() or 1
    # each step is one pass over a bounded sample so that steps+warmup stay within minutes
    rows_cap = max(256, min(wl["rows"], int(4e8 // wl["length"] // max(1, args.steps + args.warmup) * 4)))
    import numpy as np  # noqa: F401
    rate, kind, sample, _ = cpu_reference_rate(wl, nthreads, reps=max(1, args.steps), rows_cap=rows_cap)
    rows = min(wl["rows"], rows_cap)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(rate, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(rows * wl["length"] / rate / 1e6, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "step": sample},
        "cpu_baseline": {"value": round(rate, 4), "unit": UNIT, "cores": nthreads, "kind": kind, "sample": sample},
        "e2e": {"value": round(rate, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
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

    wl = WORKLOADS[args.workload]
    lib = sg.lib()
    traffic = None
    try:  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {}).get("bytes")
    except Exception:
        pass
    peaks, peak_kind = measured_peaks()
    sampler = ClockSampler(local)
    extra = {}

    if wl["kind"] in ("batch", "long", "stream"):
        from bench_workloads import run_1d_family
        res = run_1d_family(wl, args, sg, lib, torch, np, dev, rank, world, barrier, max_over_ranks, sampler, dist)
    else:
        from bench_workloads import run_2d
        res = run_2d(wl, args, sg, lib, torch, np, dev, rank, world, barrier, max_over_ranks, sampler, dist)

    if rank == 0:
        units_per_step = res["units_per_step_per_rank"] * world
        ms = res["ms_per_step"]
        value = units_per_step / (ms * 1e-3) / 1e9
        kern_ms = res["kernel_ms"]
        alg_bytes = res["alg_bytes_per_launch"]
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms, 5), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict({"workload": wl["desc"], "l2": "inputs larger than L2 (no flush needed)" if res["bytes_in"] > 2e8
                            else "L2 flushed between steps by a 256 MiB memset", "sharding": "independent units per rank, no collective"
                            if wl["kind"] != "long" else "contiguous slices of one signal per rank"}, **res.get("config", {})),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": round(achieved / peaks["hbm_gbs"], 4), "traffic": res.get("traffic", traffic), "peak_kind": peak_kind,
                         "kernel": res["kernel"], "kernel_ms": round(kern_ms, 5), "alg_bytes_per_launch": alg_bytes},
            "clocks": res["clocks"],
            "gpu_launches": res["gpu_launches"],
            "parity": res.get("parity"),
        }
        if res.get("e2e"):
            line["e2e"] = res["e2e"]
        if res.get("cpu_baseline"):
            line["cpu_baseline"] = res["cpu_baseline"]
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
