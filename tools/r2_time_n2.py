"""What makes back-to-back launches of the mid half-windows slower than a lone launch?  python tools/r2_time_n2.py <n>"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import savgol_b200 as sg
n = int(sys.argv[1])
sg.lib().savgol_b200_set_tma(2)
x = torch.randn(65536, 4096, device="cuda"); y = torch.empty_like(x)
scr = torch.empty(64 << 20, device="cuda")
f = sg.SavgolFilter(n, min(3, 2 * n), 0, 1.0, "reflect")
for _ in range(5):
    f.apply(x, out=y)
torch.cuda.synchronize()

def run(mode):
    ts = []
    for i in range(12):
        if mode == "flush":
            scr.fill_(1.0)
        if mode == "sleep":
            torch.cuda.synchronize(); time.sleep(0.02)
        if mode == "readflush":
            scr.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f.apply(x, out=y); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = ts[2:]
    print(f"n={n:2d} {mode:10s} mean {sum(ts)/len(ts):.4f} min {min(ts):.4f} max {max(ts):.4f}")
for mode in ("b2b", "sleep", "flush", "readflush", "b2b"):
    run(mode)
