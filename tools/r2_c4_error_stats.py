"""Error statistics of config 4 at its literal size (256 x 4096 x 4096): fast (additive kernel) vs exact (reference
summation order) over every pixel, and both against a float64 evaluation of the same fp32 weight table.
usage (GPU box): python tools/r2_c4_error_stats.py [images]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import savgol_b200 as sg

images = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rows = cols = 4096
g = torch.Generator(device="cuda"); g.manual_seed(3)
x = torch.rand(images, rows, cols, device="cuda", generator=g)
f = sg.Savgol2DFilter(7, 7, 3)
y = f.apply(x, "constant")
sg.set_exact(True)
ye = f.apply(x, "constant")
sg.set_exact(False)
worst, over, over9, tot = 0.0, 0, 0, 0
for i in range(0, images, 8):
    d = (y[i:i + 8] - ye[i:i + 8]).abs()
    worst = max(worst, float(d.max()))
    over += int((d > 1e-6).sum()); over9 += int((d > 9e-7).sum()); tot += d.numel()
print(f"fast vs exact over {tot} pixels: max {worst:.4e}, > 1e-6: {over} ({over / tot:.2e}), > 9e-7: {over9}")
W = torch.from_numpy(np.asarray(f.weights, np.float32).reshape(15, 15)).cuda().double()
scale = float(np.float32(f.scale))
ef = ee = 0.0
for i in range(0, min(images, 16)):
    xi = F.pad(x[i].double()[None, None], (7, 7, 7, 7), mode="replicate")
    truth = F.conv2d(xi, W[None, None])[0, 0] * scale
    ef = max(ef, float((y[i].double() - truth).abs().max()))
    ee = max(ee, float((ye[i].double() - truth).abs().max()))
print(f"against float64 truth over {min(images, 16)} images: fast max {ef:.4e}, exact (= the reference's order) max {ee:.4e}")
