"""Workload drivers for bench.py (kept separate so bench.py stays readable).

Every driver returns a dict with: units_per_step_per_rank, ms_per_step (max over ranks),
kernel_ms / kernel / alg_bytes_per_launch (dominant kernel, for the roofline), bytes_in,
gpu_launches, clocks, parity, and optionally e2e / cpu_baseline / config.
"""
from __future__ import annotations

import ctypes as C
import os
import time


def _time_region(torch, steps, call, barrier, sampler, flush=None):
    """Times `steps` calls with CUDA events on the current stream.  Without `flush` one event pair
    brackets the whole region; with it (small, L2-resident workloads) every step gets its own pair
    and the flush between steps is excluded."""
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    t0 = time.perf_counter()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
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
    t1 = time.perf_counter()
    sampler.stop()
    barrier()
    return ms, (t0, t1)


def _synthetic_batch(torch, rows, length, dev, seed):
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


def run_1d_family(wl, args, sg, lib, torch, np, dev, rank, world, barrier, max_over_ranks, sampler, dist):
    n, m, d, dt = wl["n"], wl["m"], wl["d"], wl["dt"]
    kind = wl["kind"]
    lib.savgol_b200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
    res = {"config": {}}
    tol_scale = 1e-6 / (dt ** d)

    if kind == "batch":
        rows, L = wl["rows"], wl["length"]
        f = sg.SavgolFilter(n, m, d, dt, wl["boundary"])
        x = _synthetic_batch(torch, rows, L, dev, 1 + rank)
        y = torch.empty_like(x)
        xp, yp = x.data_ptr(), y.data_ptr()

        def call():
            rc = lib.savgol_apply_batch(f.handle, xp, yp, rows, L, L, L)
            assert rc == 0
        units = rows * L
        res["kernel"] = f"sg1d_kernel<N={n},batch,FFMA2>"
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
            return {"max_abs_err": err, "tol": t, "signals_checked": len(pick), "ok": bool(err <= t)}
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
        ring = None
        halo_mode = os.environ.get("SG_C3_HALO", "p2p")   # p2p: halos read from the neighbours' HBM inside the kernel
        if world > 1:
            strip[:n] = x[:n]
            strip[n:] = x[L - n:]
            dist.all_gather_into_tensor(allstrips, strip)   # the parity checker's copy of the neighbours' edges
            if halo_mode == "p2p":
                from savgol_b200 import dist as sgdist
                ring = sgdist.PeerRing(x, n, True)
                left_p, right_p = C.c_void_p(ring.left_ptr), C.c_void_p(ring.right_ptr)
            torch.cuda.synchronize()
            dist.barrier()
        res.setdefault("config", {})["halo"] = ("none (one slice, periodic wrap inside the kernel)" if world == 1 else
                                                "ring neighbours' slices mapped by CUDA IPC, 2n samples read over NVLink inside the kernel, "
                                                "no collective per step" if ring else "NCCL all_gather of 2n floats per rank per step")

        def call():
            if world == 1:
                rc = lib.savgol_apply(f.handle, xp, yp, L)  # periodic wrap inside the kernel
            elif ring is not None:
                rc = lib.savgol_apply_halo(f.handle, xp, yp, L, left_p, right_p)
            else:
                # n-sample halo exchange between ring neighbours (NCCL all_gather of 2n floats per rank)
                strip[:n] = x[:n]
                strip[n:] = x[L - n:]
                dist.all_gather_into_tensor(allstrips, strip)
                prev, nxt = (rank - 1) % world, (rank + 1) % world
                left = allstrips[prev * 2 * n + n: prev * 2 * n + 2 * n]
                right = allstrips[nxt * 2 * n: nxt * 2 * n + n]
                rc = lib.savgol_apply_halo(f.handle, xp, yp, L, left.data_ptr(), right.data_ptr())
            assert rc == 0
        units = L
        res["kernel"] = f"sg1d_kernel<N={n},batch,FFMA2>"
        res["alg_bytes_per_launch"] = 8 * units
        res["bytes_in"] = 4 * units

        def parity():
            from oracle import oracle as O
            # Q6 identity: periodic == VALID over the wrap-padded signal; check both ends of the slice
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
            err = max(float(np.max(np.abs(o.apply_valid(head) - y[:W].cpu().numpy()))),
                      float(np.max(np.abs(o.apply_valid(tail) - y[L - W:].cpu().numpy()))))
            t = tol_scale * float(np.max(np.abs(head)))
            return {"max_abs_err": err, "tol": t, "samples_checked": 2 * W, "ok": bool(err <= t)}
    else:  # stream
        rows, K = wl["rows"], wl["length"]
        s = sg.SavgolMCStream(rows, n, m, d, dt)
        x = _synthetic_batch(torch, rows, K, dev, 4 + rank)
        OP = (K + n + 3) & ~3  # output pitch: >= K + half_window, 16-byte aligned rows
        y = torch.empty(rows, OP, device=dev)
        xp, yp = x.data_ptr(), y.data_ptr()
        k0 = lib.savgol_mcstream_push(s._h, xp, K, K, yp, OP)   # first fill (leading edge), outside the timed region
        assert k0 == K - n
        y0 = y[:32, :K - n].cpu().numpy().copy()

        def call():
            k = lib.savgol_mcstream_push(s._h, xp, K, K, yp, OP)
            assert k == K
        units = rows * K
        res["kernel"] = f"sg1d_kernel<N={n},stream,FFMA2>"
        res["alg_bytes_per_launch"] = 8 * units + 2 * (2 * n + 1) * 4 * rows
        res["bytes_in"] = 4 * units

        def parity():
            from oracle import oracle as O
            o = O.Filter1D(n, m, d, dt)
            xs = x[:32].cpu().numpy()
            # the same chunk pushed twice == the stream over [chunk | chunk]; compare the first 2K-n outputs
            ys = np.concatenate([y0, y[:32, :K].cpu().numpy()], axis=1)
            ref = np.stack([o.stream_run(np.concatenate([r, r]))[: 2 * K - n] for r in xs])
            err = float(np.max(np.abs(ys - ref)))
            t = tol_scale * float(np.max(np.abs(xs)))
            return {"max_abs_err": err, "tol": t, "channels_checked": 32, "ok": bool(err <= t)}

    flush = None
    if res["bytes_in"] < 2e8:
        scratch = torch.empty(64 << 20, device=dev, dtype=torch.float32)  # 256 MiB > 126 MB L2
        flush = lambda: scratch.fill_(1.0)
    for _ in range(args.warmup):
        call()
    torch.cuda.synchronize()
    if kind == "stream":
        # re-establish "first chunk then one more" so that parity() sees [chunk | chunk]
        s.reset()
        lib.savgol_mcstream_push(s._h, xp, K, K, yp, OP)
        y0 = y[:32, :K - n].cpu().numpy().copy()
        call()
        res["parity"] = parity() if rank == 0 else None
    c0 = lib.savgol_b200_launch_count()
    ms, (t0, t1) = _time_region(torch, args.steps, call, barrier, sampler, flush)
    res["gpu_launches"] = int(lib.savgol_b200_launch_count() - c0)
    res["kernel_ms"] = ms
    res["ms_per_step"] = max_over_ranks(ms)
    res["units_per_step_per_rank"] = units
    res["clocks"] = sampler.summary(t0, t1)
    if kind != "stream":
        res["parity"] = parity() if rank == 0 else None

    # ---- end to end: the same C-ABI call on pinned HOST buffers (H2D + D2H inside the timed region)
    if kind == "batch" and not args.no_e2e:
        ke = args.e2e_steps or max(1, min(args.steps, 5))
        xh = torch.empty(rows, L, dtype=torch.float32, pin_memory=True)
        yh = torch.empty(rows, L, dtype=torch.float32, pin_memory=True)
        xh.copy_(x)
        torch.cuda.synchronize()
        hcall = lambda: lib.savgol_apply_batch(f.handle, xh.data_ptr(), yh.data_ptr(), rows, L, L, L)
        assert hcall() == 0  # warm-up (allocates the staging ring)
        barrier()
        ta = time.perf_counter()
        for _ in range(ke):
            assert hcall() == 0
        torch.cuda.synchronize()
        tb = time.perf_counter()
        barrier()
        sec = max_over_ranks((tb - ta) / ke)
        same = bool(torch.equal(yh[:64], y[:64].cpu()))
        res["e2e"] = {"value": round(world * units / sec / 1e9, 3), "unit": "Gsamples/s", "h2d_bytes_per_step": 4 * units,
                      "d2h_bytes_per_step": 4 * units, "steps": ke, "ms_per_step": round(sec * 1e3, 3),
                      "api": "savgol_apply_batch(host pinned in, host pinned out)", "matches_device_result": same}
        del xh, yh

    # ---- CPU baseline: the unmodified reference on this box's host cores (rank 0, N=1 only)
    if kind == "batch" and world == 1 and rank == 0 and not args.no_cpu:
        from bench import cpu_reference_rate
        nthreads = os.cpu_count() or 1
        rate, ckind, sample, _ = cpu_reference_rate(wl, nthreads, reps=2)
        rate1, _, _, _ = cpu_reference_rate(wl, 1, reps=1, rows_cap=max(1, int(3e7 // wl["length"])))
        res["cpu_baseline"] = {"value": round(rate, 4), "unit": "Gsamples/s", "cores": nthreads, "kind": ckind,
                               "sample": sample, "single_thread_value": round(rate1, 4)}
    return res


def run_2d(wl, args, sg, lib, torch, np, dev, rank, world, barrier, max_over_ranks, sampler, dist):
    images, rows, cols = wl["images"], wl["rows"], wl["cols"]
    images = int(os.environ.get("SG_C4_IMAGES", images))   # experiment knobs (default = the config)
    rows = cols = int(os.environ.get("SG_C4_SIZE", rows))
    lib.savgol_b200_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
    f = sg.Savgol2DFilter(wl["nx"], wl["ny"], wl["order"])
    g = torch.Generator(device=dev)
    g.manual_seed(3 + rank)
    x = torch.rand(images, rows, cols, device=dev, generator=g, dtype=torch.float32)
    y = torch.empty_like(x)
    b = sg.BOUNDARY_2D[wl["boundary"]]
    xp, yp = x.data_ptr(), y.data_ptr()

    def call():
        rc = lib.savgol2d_apply_batch(f.handle, xp, rows, cols, cols, rows * cols, yp, cols, rows * cols, images, b)
        assert rc == 0
    for _ in range(args.warmup):
        call()
    c0 = lib.savgol_b200_launch_count()
    ms, (t0, t1) = _time_region(torch, args.steps, call, barrier, sampler, None)
    units = images * rows * cols
    res = {"config": {}, "kernel": "sg2d", "alg_bytes_per_launch": 8 * units, "bytes_in": 4 * units,
           "gpu_launches": int(lib.savgol_b200_launch_count() - c0), "kernel_ms": ms,
           "ms_per_step": max_over_ranks(ms), "units_per_step_per_rank": units, "clocks": sampler.summary(t0, t1)}
    if rank == 0:
        from oracle import oracle as O
        o = O.Filter2D(wl["nx"], wl["ny"], wl["order"])
        # borders/corners + an interior block of the first and last image
        H = 96
        errs = []
        for im in (0, images - 1):
            for (r0, c0_) in ((0, 0), (0, cols - H), (rows - H, 0), (rows - H, cols - H), (rows // 2, cols // 2)):
                # crop with enough context; compare the part whose window lies inside the crop or at a true border
                ra, rb = max(0, r0 - 16), min(rows, r0 + H + 16)
                ca, cb = max(0, c0_ - 16), min(cols, c0_ + H + 16)
                crop = x[im, ra:rb, ca:cb].cpu().numpy()
                ref = o.apply(crop, wl["boundary"])
                got = y[im, r0:r0 + H, c0_:c0_ + H].cpu().numpy()
                errs.append(float(np.max(np.abs(ref[r0 - ra:r0 - ra + H, c0_ - ca:c0_ - ca + H] - got))))
        t = 1e-6 * float(x[0].max().item())
        res["parity"] = {"max_abs_err": max(errs), "tol": t, "blocks_checked": len(errs), "ok": bool(max(errs) <= t)}
    else:
        res["parity"] = None
    return res
