"""savgol2d_gradient / _hessian on device images: concurrent component launches (default) vs the sequential composition
(SAVGOL_B200_WRAP_SEQ=1), and both against N x the single-filter time.  usage (GPU box): python tools/r2_wrappers.py"""
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
    for size in (1024, 4096, 8192):
        img = torch.rand(size, size, device="cuda")
        f = sg.Savgol2DFilter(7, 7, 3, 1, 0)
        o1 = torch.empty_like(img)
        single = t(lambda: f.apply(img, "constant", out=o1))
        g = t(lambda: sg.gradient(img, 7, 7, 3, 1.0, 1.0, "constant"))
        h = t(lambda: sg.hessian(img, 7, 7, 3, 1.0, 1.0, "constant"))
        lap = t(lambda: sg.laplacian(img, 7, 7, 3, 1.0, 1.0, "constant"))
        print(f"{size}x{size}: single filter {single:7.1f} us | gradient {g:7.1f} us ({g / single:.2f}x) | hessian {h:7.1f} us ({h / single:.2f}x) | laplacian (one fused table) {lap:7.1f} us")
else:
    for seq in ("0", "1"):
        print("== SAVGOL_B200_WRAP_SEQ=" + seq + (" (sequential composition)" if seq == "1" else " (concurrent component launches)"))
        env = dict(os.environ, SAVGOL_B200_WRAP_SEQ=seq)
        print(subprocess.run([sys.executable, __file__, "child"], capture_output=True, text=True, env=env).stdout)
