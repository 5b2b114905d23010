"""End to end through the C ABI for calls that fit ONE staging chunk (the reference's typical calls: one image, one
signal, a modest batch): adaptive chunks / row bands vs one chunk per call.  usage (GPU box): python tools/r2_host_small.py"""
import os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import savgol_b200 as sg

    tag = sys.argv[2]

    def pin(a):
        t = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
        t.copy_(torch.from_numpy(a))
        return t

    def timeit(fn, reps=10):
        fn(); fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3

    rng = np.random.default_rng(0)
    f1 = sg.SavgolFilter(16, 3, 1, 1.0, "reflect")
    fc1 = sg.SavgolFilter(12, 4, 0, 1.0, "polynomial")
    for name, f, shape in (("1D 4096 x 4096", f1, (4096, 4096)), ("1D 1024 x 4096", f1, (1024, 4096)), ("1D one 1,000,000-sample signal (C1)", fc1, (1000000,)),
                           ("1D one 16,777,216-sample signal", fc1, (1 << 24,))):
        x = rng.standard_normal(shape, dtype=np.float32)
        xp, yp = pin(x), pin(x)
        y = np.empty_like(x)
        tp = timeit(lambda: f.apply(xp, out=yp))
        tg = timeit(lambda: f.apply(x, out=y))
        assert np.array_equal(y, yp.numpy())
        print(f"[{tag:12s}] {name:38s}: pinned {tp:8.3f} ms ({x.size / tp / 1e6:6.2f} Gsamples/s) | pageable {tg:8.3f} ms ({x.size / tg / 1e6:6.2f})", flush=True)
    for hw, order in ((7, 3), (2, 2)):
        f2 = sg.Savgol2DFilter(hw, hw, order)
        for shape in ((4096, 4096), (2048, 2048), (3, 2048, 2048)):
            img = rng.random(shape, dtype=np.float32)
            ip, op = pin(img), pin(img)
            o = np.empty_like(img)
            tp = timeit(lambda: f2.apply(ip, "reflect", out=op))
            tg = timeit(lambda: f2.apply(img, "reflect", out=o))
            dev = f2.apply(torch.from_numpy(img).cuda(), "reflect").cpu().numpy()
            assert np.array_equal(o, op.numpy()) and np.array_equal(o, dev)
            print(f"[{tag:12s}] 2D {2 * hw + 1}x{2 * hw + 1} image(s) {str(shape):18s}: pinned {tp:8.3f} ms ({img.size / tp / 1e6:6.2f} Gpixel/s) | pageable {tg:8.3f} ms ({img.size / tg / 1e6:6.2f})", flush=True)
    # gradient / Hessian of ONE host image: one upload + one multi-output launch + a download per component
    # (SAVGOL_B200_WRAP_SEQ=1: the reference's composition, every component uploads the image again)
    img = rng.random((4096, 4096), dtype=np.float32)
    ip = pin(img)
    lib = sg.lib()
    outs_p = [pin(img) for _ in range(3)]
    outs_g = [np.zeros_like(img) for _ in range(3)]          # preallocated and touched: no first-touch faults in the timing
    def grad(src, o):
        assert lib.savgol2d_gradient(2, 2, 2, src, 4096, 4096, 4096, o[0], o[1], 1.0, 1.0, 1) == 0
    def hess(src, o):
        assert lib.savgol2d_hessian(2, 2, 2, src, 4096, 4096, 4096, o[0], o[1], o[2], 1.0, 1.0, 1) == 0
    pp = [t.data_ptr() for t in outs_p]
    gp = [a_.ctypes.data for a_ in outs_g]
    for name, fn in (("gradient 5x5", grad), ("hessian 5x5", hess)):
        tp = timeit(lambda: fn(ip.data_ptr(), pp), reps=5)
        tg = timeit(lambda: fn(img.ctypes.data, gp), reps=5)
        assert np.array_equal(outs_g[0], outs_p[0].numpy()) and np.array_equal(outs_g[1], outs_p[1].numpy())
        print(f"[{tag:12s}] 2D {name} of one host 4096x4096 image: pinned {tp:8.3f} ms | pageable {tg:8.3f} ms", flush=True)
    sys.exit(0)

modes = [("adaptive", {}), ("one chunk", {"SAVGOL_B200_NO_HOST_BANDS": "1", "SAVGOL_B200_FIXED_CHUNK": "1", "SAVGOL_B200_WRAP_SEQ": "1"})]
if len(sys.argv) > 1 and sys.argv[1] == "zerocopy":
    modes = [("adaptive", {}), ("zero copy", {"SAVGOL_B200_ZEROCOPY": "1"})]
for tag, env in modes:
    subprocess.run([sys.executable, os.path.abspath(__file__), "child", tag], env=dict(os.environ, **env), check=False)
