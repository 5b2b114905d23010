"""One 1D batch shape, a few launches (for ncu captures of the short-row kernel).  usage: python tools/run_shape.py L [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import savgol_b200 as sg

L = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rows = (1 << 28) // L
f = sg.SavgolFilter(n, 3, 1, 1.0, "reflect")
x = torch.randn(rows, L, device="cuda"); y = torch.empty_like(x)
for _ in range(6):
    f.apply(x, out=y)
torch.cuda.synchronize()
print("rows", rows, "len", L)
