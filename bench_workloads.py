"""Workload drivers for bench.py (kept separate so bench.py stays readable).

One driver per BASELINE config family.  Every driver returns a dict with
  units_per_step_per_rank, ms_per_step (max over ranks), kernel / kernel_ms / alg_bytes_per_launch /
  fp32_ops_per_unit (dominant kernel, for the two rooflines), bytes_in, gpu_launches, clocks, parity,
  sustained (>= 1-2 s of back-to-back launches), e2e (the C-ABI call on pinned HOST buffers, copies inside
  the timed region) and, at N = 1 on rank 0, cpu_baseline (the unmodified reference on the host cores).
"""
from __future__ import annotations

import ctypes as C
import os
import time

FP32_PEAK_TFMA = 35.5   # fp32 FMA lane-operations per second the chip sustains at 1965 MHz: tools/microbench.cu on B200


# ----------------------------------------------------------------------------------------------
# timing helpers
def time_region(torch, steps, call, barrier, sampler, flush=None):
    """Times `steps` calls with CUDA events on the current stream.  Without `flush` one event pair
    brackets the whole region; with it (small, L2-resident workloads) every step gets its own pair
    and the flush between steps is excluded."""
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    t0 = time.perf_counter()
    best = None
    if flush is None:
        # one event pair brackets the K steps (that is the reported time); the events between steps only tell how
        # the steps differ (the first launches of a burst run at full clock, later ones may meet the power cap)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            call()
            ev[i + 1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[steps]) / steps
        best = min(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
    else:
        pairs = []
        for _ in range(steps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            call()
            b.record()
            pairs.append((a, b))
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in pairs) / steps
        best = min(a.elapsed_time(b) for a, b in pairs)
    t1 = time.perf_counter()
    sampler.stop()
    barrier()
    time_region.best_ms = best
    return ms, (t0, t1)


def sustained_region(torch, call, est_ms, seconds, units, alg_bytes, fp32_ops, peaks, sampler_cls, local):
    """>= `seconds` of back-to-back launches of the same call (no flush: only meaningful for inputs larger than
    L2).  Clocks, power and throttle reasons are sampled over the whole region."""
    n = max(10, int(seconds * 1e3 / max(est_ms, 1e-3)) + 1)
    sampler = sampler_cls(local)
    torch.cuda.synchronize()
    sampler.start()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        call()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    sampler.stop()
    ms = e0.elapsed_time(e1) / n
    clk = sampler.summary(t0, t1)
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    tf = fp32_ops * units / (ms * 1e-3) / 1e12
    return {"seconds": round(t1 - t0, 3), "launches": n, "ms_per_step": round(ms, 5), "value": round(units / (ms * 1e-3) / 1e9, 3),
            "frac_hbm": round(gbs / peaks["hbm_gbs"], 4), "frac_fp32": round(tf / FP32_PEAK_TFMA, 4),
            "sm_mhz_median": clk.get("sm_mhz"), "power_w_max": clk.get("power_w_max"), "reasons": clk.get("reasons")}


def bind_near_gpu(local):
    """Best effort: pin this process (and so the first-touch placement of its pinned buffers) to the NUMA node of
    its GPU.  Returns the node id or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:      # nvml prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & ids
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def pcie_probe(torch, xh, yh, xd, yd):
    """Plain pinned-memory copies of the e2e buffers, both directions at once: the host-side ceiling."""
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        with torch.cuda.stream(s1):
            xd.copy_(xh, non_blocking=True)
        with torch.cuda.stream(s2):
            yh.copy_(yd, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 2
    return {"h2d_gbs": round(xh.numel() * 4 / dt / 1e9, 2), "d2h_gbs": round(yh.numel() * 4 / dt / 1e9, 2),
            "note": "cudaMemcpyAsync of the same pinned buffers, both directions concurrently, all ranks at once (rank 0's figures)"}


def e2e_region(torch, hcall, steps, barrier, max_over_ranks, world, units, h2d, d2h, api, extra=None):
    assert hcall() is not False   # warm-up (allocates the staging ring)
    torch.cuda.synchronize()
    barrier()
    ta = time.perf_counter()
    for _ in range(steps):
        hcall()
    torch.cuda.synchronize()
    tb = time.perf_counter()
    barrier()
    own = (tb - ta) / steps
    sec = max_over_ranks(own)
    rec = {"value": round(world * units / sec / 1e9, 3), "unit": "Gsamples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "steps": steps, "ms_per_step": round(sec * 1e3, 3), "api": api,
           "h2d_gbs_per_rank": round(h2d / own / 1e9, 2), "d2h_gbs_per_rank": round(d2h / own / 1e9, 2)}
    if extra:
        rec.update(extra)
    return rec


def synthetic_batch(torch, rows, length, dev, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    x = torch.randn(rows, length, device=dev, generator=g, dtype=torch.float32)
    t = torch.arange(length, device=dev, dtype=torch.float32)
    amp = torch.empty(rows, 1, device=dev).uniform_(0.5, 2.0, generator=g)
    frq = torch.empty(rows, 1, device=dev).uniform_(0.002, 0.05, generator=g)
    step = max(1, (1 << 26) // length)
    for r0 in range(0, rows, step):  # chunked so the temporary stays small
        x[r0:r0 + step] += amp[r0:r0 + step] * torch.sin(frq[r0:r0 + step] * t[None, :])
    return x


def make_batch_numpy(rows, length, seed, np):
    """Seeded synthetic batch for the CPU arm: N(0,1) noise + a per-signal sinusoid."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((rows, length), dtype=np.float32)
    t = np.arange(length, dtype=np.float32)
    amp = rng.uniform(0.5, 2.0, (rows, 1)).astype(np.float32)
    frq = rng.uniform(0.002, 0.05, (rows, 1)).astype(np.float32)
    x += amp * np.sin(frq * t[None, :])
    return x


# ----------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref) -- or the oracle port when it is absent -- on host threads
def _best(run, reps):
    run()  # warm (page faults, thread start)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        run()
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_rate(wl, nthreads, reps=2, budget=1.0):
    """Times the reference's CPU path of workload `wl` on `nthreads` host threads over a bounded sample.
    Returns (Gunits/s, kind, sample description, units per pass).  `budget` scales the sample (1.0 = default size)."""
    import numpy as np
    from oracle import oracle as O
    lib = O.lib()
    have = O.have_ref()
    R = O.ref() if have else None
    kind = wl["kind"]
    flags = "gcc -O2 -ffp-contract=off"
    if kind == "batch":
        rows = max(1, min(wl["rows"], int(budget * 2.7e8 // wl["length"])))
        L = wl["length"]
        x = make_batch_numpy(rows, L, 1, np)
        y = np.empty_like(x)
        if have:
            cfg = O.make_config(wl["n"], wl["m"], wl["d"], wl["dt"], wl["boundary"])
            f = R.savgol_create(C.byref(cfg))
            run = lambda: lib.sgh_apply_rows(O.fnptr(R, "savgol_apply"), C.cast(f, C.c_void_p), O._fp(x), O._fp(y), rows, L, L, L, nthreads)
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
        units = rows * L
        what = f"{rows}x{L} signals ({'full per-GPU workload' if rows == wl['rows'] else 'subset'}), savgol_apply per signal"
    elif kind == "long":
        n = wl["n"]
        total = int(min(wl["length"], budget * (1 << 27)))
        rng = np.random.default_rng(2)
        core = rng.standard_normal(total, dtype=np.float32)
        xp = np.concatenate([core[-n:], core, core[:n]])          # wrap-padded: periodic == VALID over this
        y = np.empty(total, np.float32)
        chunk = 1 << 20
        if have:
            cfg = O.make_config(wl["n"], wl["m"], wl["d"], wl["dt"], wl["boundary"])
            f = R.savgol_create(C.byref(cfg))
            run = lambda: lib.sgh_valid_chunks(O.fnptr(R, "savgol_apply_valid"), C.cast(f, C.c_void_p), O._fp(xp), O._fp(y), total, n, chunk, nthreads)
        else:
            of = O.Filter1D(wl["n"], wl["m"], wl["d"], wl["dt"], wl["boundary"])
            run = lambda: y.__setitem__(slice(None), of.apply_valid(xp))
        units = total
        what = (f"{total}-sample periodic signal (subset of the 2^29 slice) as savgol_apply_valid over the wrap-padded signal in 2^20-sample "
                f"chunks (savgol_apply itself overflows int beyond 2^31 samples, SURVEY Q5)")
    elif kind == "2d":
        rows = cols = wl["rows"]
        images = max(1, min(wl["images"], int(round(budget * nthreads))))
        rng = np.random.default_rng(3)
        x = rng.random((images, rows, cols), dtype=np.float32)
        y = np.zeros_like(x)
        b = {"valid": 0, "constant": 1, "reflect": 2}[wl["boundary"]]
        if have:
            cfg = O.Savgol2DConfig(wl["nx"], wl["ny"], wl["order"], 0, 0, 1.0, 1.0)
            f = R.savgol2d_create(C.byref(cfg))
            run = lambda: lib.sgh_apply2d_images(O.fnptr(R, "savgol2d_apply"), C.cast(f, C.c_void_p), O._fp(x), O._fp(y), images, rows, cols, b, nthreads)
        else:
            of = O.Filter2D(wl["nx"], wl["ny"], wl["order"])
            run = lambda: [of.apply(x[i], wl["boundary"], out=y[i]) for i in range(images)]
        units = images * rows * cols
        what = f"{images} images of {rows}x{cols} (subset of {wl['images']}), savgol2d_apply per image"
        reps = 1
        best = float("inf")
        t0 = time.perf_counter()
        run()
        best = time.perf_counter() - t0      # one pass only: ~3 Mpixel/s per core (y was touched by np.zeros_like)
        return units / best / 1e9, ("reference" if have else "port"), f"{what}, one pass, {flags}, {nthreads} threads over independent images", units
    else:  # stream
        K = wl["length"]
        channels = max(1, min(wl["rows"], int(budget * (1 << 18))))
        x = make_batch_numpy(channels, K, 4, np)
        y = np.empty((channels, K + 64), np.float32)
        if have:
            cfg = O.make_config(wl["n"], wl["m"], wl["d"], wl["dt"], 0)
            f = R.savgol_create(C.byref(cfg))
            run = lambda: lib.sgh_stream_channels(O.fnptr(R, "savgol_stream_init"), O.fnptr(R, "savgol_stream_push_full"), O.fnptr(R, "savgol_stream_flush"),
                                                  C.cast(f, C.c_void_p), O._fp(x), O._fp(y), channels, K, K, K + 64, 0, nthreads)
        else:
            of = O.Filter1D(wl["n"], wl["m"], wl["d"], wl["dt"])
            run = lambda: [of.stream_run(r) for r in x]
        units = channels * K
        what = f"{channels} channels x {K}-sample chunk (subset of {wl['rows']}), savgol_stream_push_full per sample"
    best = _best(run, reps)
    return units / best / 1e9, ("reference" if have else "port"), f"{what}, best of {reps}, {flags}, {nthreads} threads over independent units", units


def cpu_baseline_record(wl, single_thread=True):
    nthreads = os.cpu_count() or 1
    rate, kind, sample, _ = cpu_rate(wl, nthreads)
    rec = {"value": round(rate, 4), "unit": "Gsamples/s", "cores": nthreads, "kind": kind, "sample": sample}
    if single_thread and wl["kind"] != "2d":
        r1, _, _, _ = cpu_rate(wl, 1, reps=1, budget=0.1)
        rec["single_thread_value"] = round(r1, 4)
    return rec


# ----------------------------------------------------------------------------------------------
class Ctx:
    """Everything a driver needs from bench.py."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def _finish(ctx, res, call, units, steps, warmup, flush, parity_fn, sustain_s):
    torch, lib = ctx.torch, ctx.lib
    for _ in range(warmup):
        call()
    torch.cuda.synchronize()
    c0 = lib.savgol_b200_launch_count()
    t0c = lib.savgol_b200_tma_launch_count()
    ms, (t0, t1) = time_region(torch, steps, call, ctx.barrier, ctx.sampler, flush)
    res["gpu_launches"] = int(lib.savgol_b200_launch_count() - c0)
    res["tma_launches"] = int(lib.savgol_b200_tma_launch_count() - t0c)
    res["steps"] = steps
    res["kernel_ms"] = ms
    res["best_step_ms"] = time_region.best_ms
    res["ms_per_step"] = ctx.max_over_ranks(ms)
    res["units_per_step_per_rank"] = units
    res["clocks"] = ctx.sampler.summary(t0, t1)
    if parity_fn is not None:
        res["parity"] = parity_fn()
    if sustain_s > 0 and flush is None:
        res["sustained"] = sustained_region(torch, call, ms, sustain_s, units, res["alg_bytes_per_launch"], res["fp32_ops_per_unit"],
                                            ctx.peaks, ctx.sampler_cls, ctx.local)
    return res


def _gather_parity(ctx, own):
    """Every rank checks its own slice; rank 0 reports all of them."""
    if ctx.world == 1:
        own["ranks_ok"] = [own["ok"]]
        return own
    allp = [None] * ctx.world
    ctx.dist.all_gather_object(allp, own)
    out = dict(allp[0])
    out["max_abs_err"] = max(p["max_abs_err"] for p in allp)
    out["ranks_ok"] = [bool(p["ok"]) for p in allp]
    out["ok"] = all(out["ranks_ok"])
    return out


def run_1d_family(wl, ctx, steps, warmup, want_e2e, want_cpu, sustain_s):
    torch, np, lib, sg, dev, rank, world = ctx.torch, ctx.np, ctx.lib, ctx.sg, ctx.dev, ctx.rank, ctx.world
    n, m, d, dt = wl["n"], wl["m"], wl["d"], wl["dt"]
    kind = wl["kind"]
    lib.savgol_b200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
    res = {"config": {}, "fp32_ops_per_unit": 2 * n + 1}
    tol_scale = 1e-6 / (dt ** d)
    ring = None

    if kind == "batch":
        rows, L = wl["rows"], wl["length"]
        f = sg.SavgolFilter(n, m, d, dt, wl["boundary"])
        x = synthetic_batch(torch, rows, L, dev, 1 + rank)
        y = torch.empty_like(x)
        xp, yp = x.data_ptr(), y.data_ptr()

        def call():
            rc = lib.savgol_apply_batch(f.handle, xp, yp, rows, L, L, L)
            assert rc == 0
        units = rows * L
        res["alg_bytes_per_launch"] = 8 * units
        res["bytes_in"] = 4 * units

        def parity():
            from oracle import oracle as O
            o = O.Filter1D(n, m, d, dt, wl["boundary"])
            pick = list(range(min(rows, 32))) + list(range(max(0, rows - 32), rows))
            xs = x[pick].cpu().numpy()
            ys = y[pick].cpu().numpy()
            ref = o.apply(xs) if xs.shape[0] > 1 else o.apply(xs[0])[None, :]
            err = float(np.max(np.abs(ys - ref)))
            t = tol_scale * float(np.max(np.abs(xs)))
            return _gather_parity(ctx, {"max_abs_err": err, "tol": t, "signals_checked": len(pick), "ok": bool(err <= t)})
    elif kind == "long":
        L = wl["length"]
        f = sg.SavgolFilter(n, m, d, dt, wl["boundary"])
        g = torch.Generator(device=dev)
        g.manual_seed(2 + rank)
        x = torch.randn(L, device=dev, generator=g, dtype=torch.float32)
        y = torch.empty_like(x)
        strip = torch.empty(2 * n, device=dev)
        allstrips = torch.empty(world * 2 * n, device=dev)
        xp, yp = x.data_ptr(), y.data_ptr()
        halo_mode = os.environ.get("SG_C3_HALO", "p2p")   # p2p: halos read from the neighbours' HBM inside the kernel
        if world > 1:
            strip[:n] = x[:n]
            strip[n:] = x[L - n:]
            ctx.dist.all_gather_into_tensor(allstrips, strip)   # the parity checker's copy of the neighbours' edges
            if halo_mode == "p2p":
                from savgol_b200 import dist as sgdist
                ring = sgdist.PeerRing(x, n, True)
                left_p, right_p = C.c_void_p(ring.left_ptr), C.c_void_p(ring.right_ptr)
            torch.cuda.synchronize()
            ctx.dist.barrier()
        res["config"]["signal"] = f"{world} x 2^29 = {world * L} samples, one periodic signal, contiguous slice per rank"
        res["config"]["halo"] = ("none (one slice, periodic wrap inside the kernel)" if world == 1 else
                                 "CUDA IPC: ring neighbours' slices mapped once, 2n samples read over NVLink inside the kernel, no collective per step"
                                 if ring else "NCCL all_gather of 2n floats per rank per step")

        def call():
            if world == 1:
                rc = lib.savgol_apply(f.handle, xp, yp, L)  # periodic wrap inside the kernel
            elif ring is not None:
                rc = lib.savgol_apply_halo(f.handle, xp, yp, L, left_p, right_p)
            else:
                strip[:n] = x[:n]
                strip[n:] = x[L - n:]
                ctx.dist.all_gather_into_tensor(allstrips, strip)
                prev, nxt = (rank - 1) % world, (rank + 1) % world
                left = allstrips[prev * 2 * n + n: prev * 2 * n + 2 * n]
                right = allstrips[nxt * 2 * n: nxt * 2 * n + n]
                rc = lib.savgol_apply_halo(f.handle, xp, yp, L, left.data_ptr(), right.data_ptr())
            assert rc == 0
        units = L
        res["alg_bytes_per_launch"] = 8 * units
        res["bytes_in"] = 4 * units

        def parity():
            from oracle import oracle as O
            # Q6 identity: periodic == VALID over the wrap-padded signal.  Every rank checks both ends of its slice:
            # the seams with its ring neighbours (rank 0 / N-1: the wrap seam of the whole signal).
            o = O.Filter1D(n, m, d, dt, "periodic")
            W = 4096
            if world == 1:
                head = torch.cat([x[L - n:], x[:W + n]]).cpu().numpy()
                tail = torch.cat([x[L - W - n:], x[:n]]).cpu().numpy()
            else:
                prev, nxt = (rank - 1) % world, (rank + 1) % world
                left = allstrips[prev * 2 * n + n: prev * 2 * n + 2 * n]
                right = allstrips[nxt * 2 * n: nxt * 2 * n + n]
                head = torch.cat([left, x[:W + n]]).cpu().numpy()
                tail = torch.cat([x[L - W - n:], right]).cpu().numpy()
            mid0 = L // 2 - W // 2
            mid = x[mid0 - n: mid0 + W + n].cpu().numpy()
            err = max(float(np.max(np.abs(o.apply_valid(head) - y[:W].cpu().numpy()))),
                      float(np.max(np.abs(o.apply_valid(tail) - y[L - W:].cpu().numpy()))),
                      float(np.max(np.abs(o.apply_valid(mid) - y[mid0:mid0 + W].cpu().numpy()))))
            t = tol_scale * float(np.max(np.abs(head)))
            return _gather_parity(ctx, {"max_abs_err": err, "tol": t, "samples_checked_per_rank": 3 * W,
                                        "where": "both seams of every slice (incl. the wrap seam) + mid-slice", "ok": bool(err <= t)})
    else:  # stream
        rows, K = wl["rows"], wl["length"]
        s = sg.SavgolMCStream(rows, n, m, d, dt)
        x = synthetic_batch(torch, rows, K, dev, 4 + rank)
        OP = (K + n + 3) & ~3  # output pitch: >= K + half_window, 16-byte aligned rows
        y = torch.empty(rows, OP, device=dev)
        xp, yp = x.data_ptr(), y.data_ptr()
        state = {}

        def first_fill():
            s.reset()
            k0 = lib.savgol_mcstream_push(s._h, xp, K, K, yp, OP)   # first fill (leading edge), outside the timed region
            assert k0 == K - n
            state["y0"] = y[:32, :K - n].cpu().numpy().copy()

        first_fill()

        def call():
            k = lib.savgol_mcstream_push(s._h, xp, K, K, yp, OP)
            assert k == K
        units = rows * K
        res["alg_bytes_per_launch"] = 8 * units + 2 * (2 * n + 1) * 4 * rows
        res["bytes_in"] = 4 * units
        res["config"]["latency"] = f"fixed: output index T-{n} is produced by the chunk that brings sample T"

        def parity():
            from oracle import oracle as O
            o = O.Filter1D(n, m, d, dt)
            first_fill()     # "first chunk then one more" == the stream over [chunk | chunk]
            call()
            xs = x[:32].cpu().numpy()
            ys = np.concatenate([state["y0"], y[:32, :K].cpu().numpy()], axis=1)
            ref = np.stack([o.stream_run(np.concatenate([r, r]))[: 2 * K - n] for r in xs])
            err = float(np.max(np.abs(ys - ref)))
            t = tol_scale * float(np.max(np.abs(xs)))
            return _gather_parity(ctx, {"max_abs_err": err, "tol": t, "channels_checked": 32, "ok": bool(err <= t)})

    res["kernel"] = f"sg1d_{{tma_}}kernel<N={n},{'stream' if kind == 'stream' else 'batch'},FFMA2>"
    if kind == "batch" and rows == 1:
        # a launch this small is dominated by what brackets it: the same call on a 256-sample signal, timed the same way
        # (L2 flush, one CUDA-event pair per launch), is the floor of this measurement method
        tiny_x = torch.randn(256, device=dev)
        tiny_y = torch.empty_like(tiny_x)
        scr = torch.empty(64 << 20, device=dev, dtype=torch.float32)
        tiny = lambda: lib.savgol_apply(f.handle, tiny_x.data_ptr(), tiny_y.data_ptr(), 256)
        for _ in range(3):
            tiny()
        floor_ms, _ = time_region(torch, 10, tiny, ctx.barrier, ctx.sampler, lambda: scr.fill_(1.0))
        res["config"]["launch_floor_ms"] = round(floor_ms, 5)
        res["config"]["note"] = "launch bound: 8 MB, fewer segments than resident warps; the kernel alone takes 6 us under ncu (profiles/r2_bench_launches_summary.txt)"
        del scr
    flush = None
    if res["bytes_in"] < 2e8:
        scratch = torch.empty(64 << 20, device=dev, dtype=torch.float32)  # 256 MiB > 126 MB L2
        flush = lambda: scratch.fill_(1.0)
    _finish(ctx, res, call, units, steps, warmup, flush, parity, sustain_s)
    res["kernel"] = res["kernel"].replace("{tma_}", "tma_" if res["tma_launches"] else "")

    # ---- strong scaling of the literal config ("65,536 signals sharded across N GPUs"): every rank filters 1/N of ONE
    # 65,536-signal batch.  The per-rank share shrinks towards the L2 size, so L2 is flushed before every launch.
    if kind == "batch" and world > 1 and rows >= world:
        rs = rows // world
        scratch_s = torch.empty(64 << 20, device=dev, dtype=torch.float32)

        def call_s():
            assert lib.savgol_apply_batch(f.handle, xp, yp, rs, L, L, L) == 0
        ms_s, _ = time_region(torch, max(5, min(steps, 10)), call_s, ctx.barrier, ctx.sampler, lambda: scratch_s.fill_(1.0))
        ms_s = ctx.max_over_ranks(ms_s)
        res["strong_scaling"] = {"signals_total": rs * world, "signals_per_rank": rs, "ms_per_step": round(ms_s, 5),
                                 "value": round(rs * world * L / (ms_s * 1e-3) / 1e9, 3),
                                 "frac_hbm_per_gpu": round(8 * rs * L / (ms_s * 1e-3) / 1e9 / ctx.peaks["hbm_gbs"], 4),
                                 "l2": "flushed before every launch (256 MiB fill, excluded from the timing)"}
        del scratch_s

    # ---- end to end: the same C-ABI call on pinned HOST buffers (H2D + D2H inside the timed region)
    if want_e2e:
        ke = max(1, min(steps, 5))
        if kind == "batch":
            xh = torch.empty(rows, L, dtype=torch.float32, pin_memory=True)
            yh = torch.empty(rows, L, dtype=torch.float32, pin_memory=True)
            xh.copy_(x)
            torch.cuda.synchronize()
            hcall = lambda: lib.savgol_apply_batch(f.handle, xh.data_ptr(), yh.data_ptr(), rows, L, L, L) == 0
            res["e2e"] = e2e_region(torch, hcall, ke, ctx.barrier, ctx.max_over_ranks, world, units, 4 * units, 4 * units,
                                    "savgol_apply_batch(host pinned in, host pinned out)")
            res["e2e"]["matches_device_result"] = bool(torch.equal(yh[:64], y[:64].cpu()))
            if rows * L >= (1 << 26):
                ctx.barrier()   # every rank copies at the same time: the probe sees the same host-side contention as the e2e steps
                res["e2e"]["pcie_probe"] = pcie_probe(torch, xh, yh, x, y)
                ctx.barrier()
        elif kind == "long":
            Le = min(L, 1 << 28)
            xh = torch.empty(Le, dtype=torch.float32, pin_memory=True)
            yh = torch.empty(Le, dtype=torch.float32, pin_memory=True)
            xh.copy_(x[:Le])
            torch.cuda.synchronize()
            hcall = lambda: lib.savgol_apply(f.handle, xh.data_ptr(), yh.data_ptr(), Le) == 0
            res["e2e"] = e2e_region(torch, hcall, ke, ctx.barrier, ctx.max_over_ranks, world, Le, 4 * Le, 4 * Le,
                                    "savgol_apply(host pinned in, host pinned out)",
                                    {"sample": f"2^28-sample host signal per rank (half a slice), periodic within itself; staged in {os.environ.get('SAVGOL_B200_CHUNK_MIB', '64')} MiB pieces with n-sample halos"})
            mid = Le // 2
            res["e2e"]["matches_device_result"] = bool(torch.equal(yh[mid:mid + 4096], y[mid:mid + 4096].cpu()))
        else:
            Ce = min(rows, 1 << 18)
            se = sg.SavgolMCStream(Ce, n, m, d, dt)
            xh = torch.empty(Ce, K, dtype=torch.float32, pin_memory=True)
            yh = torch.empty(Ce, OP, dtype=torch.float32, pin_memory=True)
            xh.copy_(x[:Ce])
            torch.cuda.synchronize()
            assert lib.savgol_mcstream_push(se._h, xh.data_ptr(), K, K, yh.data_ptr(), OP) == K - n   # first fill
            hcall = lambda: lib.savgol_mcstream_push(se._h, xh.data_ptr(), K, K, yh.data_ptr(), OP) == K
            res["e2e"] = e2e_region(torch, hcall, ke, ctx.barrier, ctx.max_over_ranks, world, Ce * K, 4 * Ce * K, 4 * Ce * K,
                                    "savgol_mcstream_push(host pinned chunk in, host pinned out)",
                                    {"sample": f"{Ce} channels x {K}-sample chunk per rank (subset of {rows} channels)"})
            res["e2e"]["matches_device_result"] = bool(torch.equal(yh[:64, :K], y[:64, :K].cpu()))
            del se
        del xh, yh

    if want_cpu:
        res["cpu_baseline"] = cpu_baseline_record(wl)
    if ring is not None:
        torch.cuda.synchronize()
        ctx.barrier()
        ring.close()
    return res


def run_2d(wl, ctx, steps, warmup, want_e2e, want_cpu, sustain_s):
    torch, np, lib, sg, dev, rank, world = ctx.torch, ctx.np, ctx.lib, ctx.sg, ctx.dev, ctx.rank, ctx.world
    images, rows, cols = wl["images"], wl["rows"], wl["cols"]
    images = int(os.environ.get("SG_C4_IMAGES", images))   # experiment knobs (default = the config)
    rows = cols = int(os.environ.get("SG_C4_SIZE", rows))
    lib.savgol_b200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
    f = sg.Savgol2DFilter(wl["nx"], wl["ny"], wl["order"])
    plan_kind = int(lib.savgol2d_b200_plan_kind(f.handle))
    g = torch.Generator(device=dev)
    g.manual_seed(3 + rank)
    x = torch.rand(images, rows, cols, device=dev, generator=g, dtype=torch.float32)
    y = torch.empty_like(x)
    b = sg.BOUNDARY_2D[wl["boundary"]]
    xp, yp = x.data_ptr(), y.data_ptr()

    def call():
        rc = lib.savgol2d_apply_batch(f.handle, xp, rows, cols, cols, rows * cols, yp, cols, rows * cols, images, b)
        assert rc == 0
    units = images * rows * cols
    n = wl["nx"]
    # fp32 pipe operations the kernel executes per pixel (not the 225 MACs of the literal window):
    #   additive kernel: row pass 2n+1 (u, sample-broadcast FFMA2) + box (~2n/4 + 1.5 + 1.5), column pass ~ (6n - 1) / 2
    #   rank-R kernel:   fold n + R (n + 1) + R (2n + 1)
    ops = (2 * n + 1) + (n / 2 + 3) + (3 * n - 0.5) if plan_kind == 2 else n + 2 * (n + 1) + 2 * (2 * n + 1)
    res = {"config": {"images": images}, "kernel": {2: "sg2d::sep_kernel<ADD> (additive u(x)+v(y), box-sum factors)", 1: "sg2d::sep_kernel (rank-R factors)"}.get(plan_kind, "sg2d::direct_kernel"),
           "alg_bytes_per_launch": 8 * units, "bytes_in": 4 * units, "fp32_ops_per_unit": round(ops, 2)}

    def parity():
        from oracle import oracle as O
        o = O.Filter2D(wl["nx"], wl["ny"], wl["order"])
        # borders/corners + an interior block of the first, middle and last image
        H = 96
        errs = []
        for im in sorted({0, images // 2, images - 1}):
            for (r0, c0_) in ((0, 0), (0, cols - H), (rows - H, 0), (rows - H, cols - H), (rows // 2, cols // 2)):
                ra, rb = max(0, r0 - 16), min(rows, r0 + H + 16)
                ca, cb = max(0, c0_ - 16), min(cols, c0_ + H + 16)
                crop = x[im, ra:rb, ca:cb].cpu().numpy()
                ref = o.apply(crop, wl["boundary"])
                got = y[im, r0:r0 + H, c0_:c0_ + H].cpu().numpy()
                errs.append(float(np.max(np.abs(ref[r0 - ra:r0 - ra + H, c0_ - ca:c0_ - ca + H] - got))))
        t = 1e-6 * float(x[0].max().item())
        return _gather_parity(ctx, {"max_abs_err": max(errs), "tol": t, "blocks_checked": len(errs), "ok": bool(max(errs) <= t)})

    _finish(ctx, res, call, units, steps, warmup, None, parity, sustain_s)

    if want_e2e:
        ie = min(images, 16)
        xh = torch.empty(ie, rows, cols, dtype=torch.float32, pin_memory=True)
        yh = torch.empty(ie, rows, cols, dtype=torch.float32, pin_memory=True)
        xh.copy_(x[:ie])
        torch.cuda.synchronize()
        hcall = lambda: lib.savgol2d_apply_batch(f.handle, xh.data_ptr(), rows, cols, cols, rows * cols, yh.data_ptr(), cols, rows * cols, ie, b) == 0
        ue = ie * rows * cols
        res["e2e"] = e2e_region(torch, hcall, max(1, min(steps, 3)), ctx.barrier, ctx.max_over_ranks, world, ue, 4 * ue, 4 * ue,
                                "savgol2d_apply_batch(host pinned images in, host pinned out)",
                                {"sample": f"{ie} images of {rows}x{cols} per rank (subset of {images}), one image per staging slot"})
        res["e2e"]["matches_device_result"] = bool(torch.equal(yh[0, :64], y[0, :64].cpu()))
        del xh, yh
    if want_cpu:
        res["cpu_baseline"] = cpu_baseline_record(wl, single_thread=False)
    return res
