"""End to end through the C ABI with PAGEABLE host buffers (plain malloc / numpy memory, what a drop-in caller of the
reference passes) vs pinned ones.  usage (GPU box): python tools/r2_pageable.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import savgol_b200 as sg

lib = sg.lib()
f = sg.SavgolFilter(16, 3, 1, 1.0, "reflect")
for rows in (4096, 65536):
    L = 4096
    x = np.random.default_rng(0).standard_normal((rows, L), dtype=np.float32)
    y = np.empty_like(x)
    xp = torch.empty(rows, L, dtype=torch.float32, pin_memory=True); xp.copy_(torch.from_numpy(x))
    yp = torch.empty(rows, L, dtype=torch.float32, pin_memory=True)
    for name, a, b in (("pageable", x.ctypes.data, y.ctypes.data), ("pinned", xp.data_ptr(), yp.data_ptr())):
        assert lib.savgol_apply_batch(f.handle, a, b, rows, L, L, L) == 0
        t0 = time.perf_counter()
        for _ in range(3):
            assert lib.savgol_apply_batch(f.handle, a, b, rows, L, L, L) == 0
        dt = (time.perf_counter() - t0) / 3
        print(f"{rows} x {L} {name:9s}: {dt * 1e3:8.2f} ms  {rows * L / dt / 1e9:6.2f} Gsamples/s  ({rows * L * 4 / dt / 1e9:5.1f} GB/s each way)")
    assert np.array_equal(y, yp.numpy())
