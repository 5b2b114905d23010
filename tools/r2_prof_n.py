"""One launch of the 1D batch kernel for a given half-window (for ncu): python tools/r2_prof_n.py <n>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import savgol_b200 as sg
n = int(sys.argv[1])
x = torch.randn(65536, 4096, device="cuda"); y = torch.empty_like(x)
f = sg.SavgolFilter(n, 3, 0, 1.0, "reflect")
for _ in range(4):
    f.apply(x, out=y)
torch.cuda.synchronize()
