"""TMA vs cp.async 1D kernels over half-windows and launch sizes (device-resident, CUDA events).
usage (GPU box): python tools/r2_sweep1d.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import savgol_b200 as sg

lib = sg.lib()
flush = torch.empty(64 << 20, device="cuda")


def timeit(f, x, y, reps, do_flush):
    for _ in range(3):
        f.apply(x, out=y)
    torch.cuda.synchronize()
    tot = 0.0
    if do_flush:
        for _ in range(reps):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f.apply(x, out=y); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f.apply(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("== half-window sweep, 65536 x 4096, m3 d0 reflect (d0: no zero centre weight)")
x = torch.randn(65536, 4096, device="cuda"); y = torch.empty_like(x)
for n in range(1, 33):
    f = sg.SavgolFilter(n, min(3, 2 * n), 0, 1.0, "reflect")
    r = []
    for on in (1, 0):
        lib.savgol_b200_set_tma(2 if on else 0)
        r.append(timeit(f, x, y, 10, False))
    lib.savgol_b200_set_tma(1)
    print(f"n={n:2d} tma {r[0]:.4f} ms ({8*x.numel()/r[0]/1e6/6553.6:.3f})  cp.async {r[1]:.4f} ms ({8*x.numel()/r[1]/1e6/6553.6:.3f})  ratio {r[1]/r[0]:.3f}")
del x, y
print("== launch-size sweep, n16 m3 d1 reflect, rows x 4096 (L2 flushed between launches)")
for rows in (16, 64, 256, 1024, 2048, 4096, 8192, 16384):
    x = torch.randn(rows, 4096, device="cuda"); y = torch.empty_like(x)
    f = sg.SavgolFilter(16, 3, 1, 1.0, "reflect")
    r = []
    for on in (1, 0):
        lib.savgol_b200_set_tma(2 if on else 0)
        r.append(timeit(f, x, y, 10, True))
    lib.savgol_b200_set_tma(1)
    print(f"rows={rows:6d} segments={rows*4:6d} tma {r[0]*1e3:.1f} us  cp.async {r[1]*1e3:.1f} us")
print("== one long row")
for L, n in ((1_000_000, 12), (1 << 24, 12), (1 << 28, 32), (1 << 28, 16)):
    x = torch.randn(L, device="cuda"); y = torch.empty_like(x)
    f = sg.SavgolFilter(n, 4, 0, 1.0, "polynomial")
    r = []
    for on in (1, 0):
        lib.savgol_b200_set_tma(2 if on else 0)
        r.append(timeit(f, x, y, 10, L < (1 << 26)))
    lib.savgol_b200_set_tma(1)
    print(f"L={L} n={n} tma {r[0]*1e3:.1f} us  cp.async {r[1]*1e3:.1f} us")
