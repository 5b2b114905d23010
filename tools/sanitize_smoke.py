"""Small run of every kernel path, meant to be executed under compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import savgol_b200 as sg

rng = np.random.default_rng(0)
for n, m, d in [(1, 1, 0), (7, 3, 1), (16, 3, 1), (21, 4, 2), (32, 4, 2)]:
    for mode in ("polynomial", "reflect", "periodic", "constant"):
        f = sg.SavgolFilter(n, m, d, 1.0, mode)
        for shape in [(2 * n + 1,), (3000,), (5, 4096), (3, 1027)]:
            x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).cuda()
            f.apply(x)
        x = torch.from_numpy(rng.standard_normal(5000).astype(np.float32)).cuda()
        f.apply_valid(x)
        f.apply(x, out=x)            # in-place (scratch path)
        f.close()
# short rows (packed kernel: 2 / 4 / 8 rows per warp), aligned and odd pitches, offset views
for n, m, d in [(2, 2, 0), (12, 4, 0), (25, 4, 2), (32, 5, 0)]:
    for mode in ("polynomial", "reflect", "periodic", "constant"):
        f = sg.SavgolFilter(n, m, d, 1.0, mode)
        for rows, L, pitch in [(7, 2 * n + 1, 2 * n + 1), (40, 20, 20), (70, 50, 51), (33, 64, 64), (33, 128, 128), (9, 250, 253), (5, 360, 360), (17, 512, 516), (3, 100, 101)]:
            if L < 2 * n + 1:
                continue
            big = torch.from_numpy(rng.standard_normal((rows, pitch + 1)).astype(np.float32)).cuda()
            f.apply(big[:, 1:1 + L])
            f.apply(big[:, :L])
        # unaligned long rows (out-of-line staging / store paths)
        big = torch.from_numpy(rng.standard_normal((3, 4099)).astype(np.float32)).cuda()
        f.apply(big[:, 1:4098])
        f.close()
f = sg.SavgolFilter(5, 2, 1, 0.5)
rec = torch.zeros(3 * 700, device="cuda")
assert f.apply_strided(rec.data_ptr(), 12, 4, rec.data_ptr(), 12, 8, 700) == 0
s = sg.SavgolMCStream(33, 10, 2, 1, 1.0)
for K in (5, 30, 1024, 77):
    s.push(torch.from_numpy(rng.standard_normal((33, K)).astype(np.float32)).cuda())
s.flush(torch.empty(1, device="cuda"))
blob = s.save()
s2 = sg.SavgolMCStream(33, 10, 2, 1, 1.0)
s2.restore(blob)
s2.push(torch.from_numpy(rng.standard_normal((33, 300)).astype(np.float32)).cuda())   # short chunk: packed stream kernel
for nx, ny, o, dx, dy in [(2, 2, 2, 0, 0), (7, 7, 3, 0, 0), (7, 7, 3, 1, 0), (12, 12, 5, 0, 0), (16, 16, 6, 0, 0), (7, 3, 3, 0, 0), (2, 9, 4, 0, 1)]:
    f2 = sg.Savgol2DFilter(nx, ny, o, dx, dy)
    for b in ("valid", "constant", "reflect"):
        for shape in [(2 * ny + 3, 2 * nx + 9), (150, 300), (2, 70, 260), (333, 131), (1, 600, 64)]:
            img = torch.from_numpy(rng.random(shape).astype(np.float32)).cuda()
            f2.apply(img, b)
fb = sg.Savgol2DFilter(7, 7, 3)
imgb = torch.from_numpy(rng.random((300, 260)).astype(np.float32)).cuda()
fb.apply_band(imgb[0:107], 0, 7, "reflect")          # top band (image border above), neighbour rows below
fb.apply_band(imgb[93:207], 7, 7, "constant")        # interior band
fb.apply_band(imgb[193:300], 7, 0, "constant")       # bottom band
# bulk-tensor (TMA) paths: 1D kernels forced on for every eligible layout (full / ragged segments, halos, stream
# chunks), 2D interior work items of aligned images (additive and rank-R kernels, half-windows 5..8)
sg.lib().savgol_b200_set_tma(2)
for n, m, d in [(1, 1, 0), (10, 2, 1), (16, 3, 1), (17, 4, 2), (32, 4, 2)]:
    for mode in ("polynomial", "reflect", "periodic", "constant"):
        f = sg.SavgolFilter(n, m, d, 1.0, mode)
        for rows, L, pitch in [(1, 1024, 1024), (3, 2048, 2052), (2, 5000, 5000), (1, 1024 + 33, 1060), (2, 4096, 4096)]:
            big = torch.from_numpy(rng.standard_normal((rows, pitch)).astype(np.float32)).cuda()
            f.apply(big[:, :L])
        x = torch.from_numpy(rng.standard_normal(3 * 4096).astype(np.float32)).cuda()
        if n % 4 == 0:
            f.apply_valid(x)
        f.apply_halo(x[4096:8192], x[4096 - n:4096].clone(), x[8192:8192 + n].clone())
        f.close()
st = sg.SavgolMCStream(40, 10, 2, 1, 1.0)
for K in (1024, 1024, 2048):
    st.push(torch.from_numpy(rng.standard_normal((40, K)).astype(np.float32)).cuda(), out=torch.empty(40, K + 12, device="cuda"))
sg.lib().savgol_b200_set_tma(1)
for nx, o in [(5, 3), (7, 3), (8, 2), (7, 4), (6, 6), (9, 3), (16, 2)]:
    f2 = sg.Savgol2DFilter(nx, nx, o)
    for b in ("valid", "constant", "reflect"):
        for shape in [(3, 700, 520), (1, 1100, 640), (2, 64, 1024), (1, 200, 517)]:   # 517: rows not 16-byte aligned, interior strips
            f2.apply(torch.from_numpy(rng.random(shape).astype(np.float32)).cuda(), b)
# host staging: pipelines, in place, VALID, stream chunks, single-process multi-GPU entry points
hx = rng.standard_normal((37, 3000)).astype(np.float32)
fh = sg.SavgolFilter(12, 4, 0, 1.0, "reflect")
fh.apply(hx)
fh.apply(hx, out=hx)
fh.apply_valid(hx[0])
sh = sg.SavgolMCStream(9, 6, 2, 0, 1.0)
for K in (4, 40, 1500):
    sh.push(rng.standard_normal((9, K)).astype(np.float32))
import ctypes as C
devs = (C.c_int * 3)(0, 0, 0)
hy = np.empty_like(hx)
assert sg.lib().savgol_apply_batch_multi(fh.handle, hx.ctypes.data, hy.ctypes.data, 37, 3000, 3000, 3000, devs, 3) == 0
imgs = rng.random((5, 120, 260)).astype(np.float32)
imgo = np.empty_like(imgs)
f7 = sg.Savgol2DFilter(7, 7, 3)
assert sg.lib().savgol2d_apply_batch_multi(f7.handle, imgs.ctypes.data, 120, 260, 260, 120 * 260, imgo.ctypes.data, 260, 120 * 260, 5, 1, devs, 3) == 0
sl = [torch.from_numpy(rng.standard_normal(5000).astype(np.float32)).cuda() for _ in range(3)]
so = [torch.empty(5000, device="cuda") for _ in range(3)]
torch.cuda.synchronize()
assert sg.lib().savgol_apply_slices(fh.handle, (C.c_void_p * 3)(*[t.data_ptr() for t in sl]), (C.c_void_p * 3)(*[t.data_ptr() for t in so]),
                                    (C.c_size_t * 3)(5000, 5000, 5000), devs, 3) == 0
# misaligned rows on the per-row phase (every phase of input and output), short tails behind the last full segment
for n, m, d, mode in [(3, 2, 0, "polynomial"), (16, 3, 1, "reflect"), (32, 4, 2, "periodic"), (10, 2, 1, "constant")]:
    f = sg.SavgolFilter(n, m, d, 1.0, mode)
    for L in (1025, 1056, 1057, 2049, 4095, 4097, 4130):
        for off_in, off_out in ((0, 0), (1, 1), (2, 3), (3, 0), (5, 5), (31, 31)):
            big = torch.from_numpy(rng.standard_normal((3, L + 40)).astype(np.float32)).cuda()
            out = torch.empty(3, L + 40, device="cuda")
            f.apply(big[:, off_in:off_in + L], out=out[:, off_out:off_out + L])
        f.apply_valid(torch.from_numpy(rng.standard_normal(L + 1).astype(np.float32)).cuda()[1:])
    f.close()
sm = sg.SavgolMCStream(7, 10, 2, 1, 1.0)
for K in (1030, 1056, 2049):
    sm.push(torch.from_numpy(rng.standard_normal((7, K + 1)).astype(np.float32)).cuda()[:, 1:])
# gradient / Hessian: one multi-output launch (every instantiated rank combination), host images uploaded once
for hw, order in [(1, 2), (2, 2), (2, 3), (3, 5), (4, 4), (5, 3), (7, 3), (8, 5)]:
    for shape in [(70, 260), (333, 131), (40, 1030)]:
        img = torch.from_numpy(rng.random(shape).astype(np.float32)).cuda()
        for b in ("constant", "reflect"):
            sg.gradient(img, hw, hw, order, 1.0, 0.5, b)
            sg.hessian(img, hw, hw, order, 1.0, 0.5, b)
sg.gradient(rng.random((90, 200)).astype(np.float32), 3, 2, 3, 1.0, 1.0, "reflect")
sg.hessian(rng.random((90, 200)).astype(np.float32), 2, 2, 2, 1.0, 1.0, "valid")
# pageable host buffers larger than a bounce chunk (host copy pool + pinned bounce buffers)
hb = rng.standard_normal((5000, 1100)).astype(np.float32)
fh.apply(hb)
fh.apply(hb, out=hb)
sg.set_exact(True)
f2 = sg.Savgol2DFilter(3, 2, 3)
f2.apply(torch.from_numpy(rng.random((40, 50)).astype(np.float32)).cuda(), "reflect")
sg.SavgolFilter(9, 3, 0).apply(torch.from_numpy(rng.standard_normal(2500).astype(np.float32)).cuda())
sg.set_exact(False)
torch.cuda.synchronize()
print("sanitize smoke done, launches:", sg.launch_count())
