"""Device-resident time of the 1D batch kernel for one half-window in a fresh process: python tools/r2_time_n.py <n> [tma 0|1|2]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import savgol_b200 as sg
n = int(sys.argv[1]); how = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sg.lib().savgol_b200_set_tma(how)
x = torch.randn(65536, 4096, device="cuda"); y = torch.empty_like(x)
f = sg.SavgolFilter(n, min(3, 2 * n), 0, 1.0, "reflect")
for _ in range(5):
    f.apply(x, out=y)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
ev[0].record()
for i in range(20):
    f.apply(x, out=y)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(20)]
ms = ev[0].elapsed_time(ev[20]) / 20
print(f"n={n:2d} tma={how} mean {ms:.4f} ms ({8 * x.numel() / ms / 1e6 / 6553.6:.3f})  min {min(ts):.4f} max {max(ts):.4f}  first5 {[round(t, 4) for t in ts[:5]]}")
