"""End to end through the C ABI with PAGEABLE host buffers (plain malloc / numpy memory, what a drop-in caller of the
reference passes) vs pinned ones, over the host copy-pool width and bounce chunk.
usage (GPU box): python tools/r2_pageable.py"""
import os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import savgol_b200 as sg

    lib = sg.lib()
    f = sg.SavgolFilter(16, 3, 1, 1.0, "reflect")
    tag = sys.argv[2]
    for rows in (4096, 65536):
        L = 4096
        x = np.random.default_rng(0).standard_normal((rows, L), dtype=np.float32)
        y = np.empty_like(x)
        xp = torch.empty(rows, L, dtype=torch.float32, pin_memory=True); xp.copy_(torch.from_numpy(x))
        yp = torch.empty(rows, L, dtype=torch.float32, pin_memory=True)
        kinds = [("pageable", x.ctypes.data, y.ctypes.data)]
        if tag == "default":
            kinds.append(("pinned", xp.data_ptr(), yp.data_ptr()))
        for name, a, b in kinds:
            t0 = time.perf_counter()
            assert lib.savgol_apply_batch(f.handle, a, b, rows, L, L, L) == 0
            first = time.perf_counter() - t0
            t0 = time.perf_counter()
            for _ in range(3):
                assert lib.savgol_apply_batch(f.handle, a, b, rows, L, L, L) == 0
            dt = (time.perf_counter() - t0) / 3
            print(f"[{tag:22s}] {rows:6d} x {L} {name:9s}: {dt * 1e3:8.2f} ms  {rows * L / dt / 1e9:6.2f} Gsamples/s  "
                  f"({rows * L * 4 / dt / 1e9:5.1f} GB/s each way)  first call {first * 1e3:7.1f} ms", flush=True)
        if tag == "default":
            assert np.array_equal(y, yp.numpy())
    sys.exit(0)

print("host cpus:", os.cpu_count(), flush=True)
runs = [("default", {})]
runs += [(f"threads={t}", {"SAVGOL_B200_COPY_THREADS": str(t)}) for t in (1, 4, 8, 16)]
runs += [(f"bounce_mib={m}", {"SAVGOL_B200_BOUNCE_MIB": str(m)}) for m in (8, 32)]
runs += [("cached stores", {"SAVGOL_B200_COPY_NT": "0"})]
runs += [("driver staging", {"SAVGOL_B200_NO_BOUNCE": "1"})]
for tag, env in runs:
    subprocess.run([sys.executable, os.path.abspath(__file__), "child", tag], env=dict(os.environ, **env), check=False)
