"""savgol2d_gradient / _hessian on device images: ONE multi-output launch (default) vs concurrent per-component launches
(SAVGOL_B200_WRAP_FUSED=0) vs the sequential composition (SAVGOL_B200_WRAP_SEQ=1), all against the single-filter time.
usage (GPU box): python tools/r2_wrappers.py"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import savgol_b200 as sg
    flush = torch.empty(64 << 20, device="cuda")

    def t(fn, reps=20):
        for _ in range(3):
            fn()
        tot = 0.0
        for _ in range(reps):
            flush.fill_(1.0)          # the image must come from HBM, not from the previous repetition's L2 lines
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps * 1e3
    for hw, order in ((2, 2), (2, 3), (3, 3), (4, 4), (7, 3)):
      for size in (1024, 4096, 8192):
        img = torch.rand(size, size, device="cuda")
        f = sg.Savgol2DFilter(hw, hw, order, 1, 0)
        o1 = torch.empty_like(img)
        single = t(lambda: f.apply(img, "constant", out=o1))
        c0 = sg.launch_count()
        sg.gradient(img, hw, hw, order, 1.0, 1.0, "constant")
        nl = sg.launch_count() - c0
        g = t(lambda: sg.gradient(img, hw, hw, order, 1.0, 1.0, "constant"))
        h = t(lambda: sg.hessian(img, hw, hw, order, 1.0, 1.0, "constant"))
        lap = t(lambda: sg.laplacian(img, hw, hw, order, 1.0, 1.0, "constant"))
        print(f"{2 * hw + 1}x{2 * hw + 1} order {order} {size}x{size}: single filter {single:7.1f} us | gradient {g:7.1f} us ({g / single:.2f}x, {nl} launch) | hessian {h:7.1f} us ({h / single:.2f}x) | laplacian (one table) {lap:7.1f} us")
else:
    for name, extra in (("default: ONE multi-output launch (half-windows <= 8)", {}),
                        ("SAVGOL_B200_WRAP_FUSED=0: concurrent per-component launches", {"SAVGOL_B200_WRAP_FUSED": "0"}),
                        ("SAVGOL_B200_WRAP_SEQ=1: sequential composition", {"SAVGOL_B200_WRAP_SEQ": "1"})):
        print("== " + name, flush=True)
        r = subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True, env=dict(os.environ, **extra))
        print(r.stdout + r.stderr[-2000:], flush=True)
