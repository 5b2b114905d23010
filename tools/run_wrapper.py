"""A few gradient / Hessian calls on one device image (for ncu captures of the multi-output 2D kernel).
usage: python tools/run_wrapper.py [half_window] [order] [size]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import savgol_b200 as sg

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 2
order = int(sys.argv[2]) if len(sys.argv) > 2 else 2
size = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
img = torch.rand(size, size, device="cuda")
for _ in range(4):
    sg.gradient(img, hw, hw, order, 1.0, 1.0, "constant")
    sg.hessian(img, hw, hw, order, 1.0, 1.0, "constant")
torch.cuda.synchronize()
print("launches", sg.launch_count())
