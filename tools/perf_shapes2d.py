"""Device-resident throughput of the 2D filter for a few shapes (aligned / odd widths, window sizes).
usage (on a GPU box): python tools/perf_shapes2d.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import savgol_b200 as sg

for (nx, ny, order, images, rows, cols) in [(7, 7, 3, 16, 4096, 4096), (7, 7, 3, 16, 4096, 4095), (7, 7, 3, 1, 4096, 4096), (7, 7, 3, 64, 1024, 1024),
                                             (2, 2, 2, 16, 4096, 4096), (4, 4, 4, 16, 4096, 4096), (7, 3, 3, 16, 4096, 4096), (9, 9, 3, 16, 4096, 4096), (10, 10, 5, 16, 4096, 4096), (12, 12, 3, 16, 4096, 4096), (16, 16, 3, 16, 4096, 4096), (8, 8, 3, 16, 4096, 4096),
                                             (16, 16, 6, 16, 4096, 4096), (8, 8, 6, 16, 4096, 4096), (8, 8, 4, 16, 4096, 4096), (7, 7, 5, 16, 4096, 4096),
                                             (6, 6, 6, 16, 4096, 4096), (4, 4, 6, 16, 4096, 4096)]:
    f = sg.Savgol2DFilter(nx, ny, order)
    x = torch.rand(images, rows, cols, device="cuda")
    y = torch.empty_like(x)
    for _ in range(2):
        f.apply(x, "constant", out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f.apply(x, "constant", out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    px = images * rows * cols
    print(f"window {2*nx+1}x{2*ny+1} order {order} images {images} x {rows}x{cols}: {ms:8.3f} ms {px / ms / 1e6:8.1f} Gpixel/s  {8 * px / ms / 1e6 / 6553.6 * 100:5.1f} % of HBM roofline")
