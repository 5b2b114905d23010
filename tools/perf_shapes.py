"""Device-resident throughput of the 1D batch kernel for awkward shapes (ragged, misaligned, short rows).
usage (on a GPU box): python tools/perf_shapes.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import savgol_b200 as sg

f = sg.SavgolFilter(16, 3, 1, 1.0, "reflect")
total = 1 << 28
for L, pitch in [(4096, 4096), (4097, 4097), (4096, 4100), (4095, 4095), (5000, 5000), (2060, 2060), (1040, 1040), (1024, 1024), (1000, 1000), (512, 512),
                 (360, 360), (256, 256), (250, 250), (128, 128), (100, 101), (64, 64), (48, 48), (33, 33), (65536, 65536), (1 << 20, 1 << 20)]:
    rows = max(1, total // L)
    x = torch.randn(rows, pitch, device="cuda")[:, :L]
    y = torch.empty(rows, pitch, device="cuda")[:, :L]
    for _ in range(3):
        f.apply(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f.apply(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"L={L:8d} pitch={pitch:8d} rows={rows:8d}: {ms:8.3f} ms  {rows * L / ms / 1e6:8.1f} Gsamples/s")
