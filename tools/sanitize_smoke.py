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
sg.set_exact(True)
f2 = sg.Savgol2DFilter(3, 2, 3)
f2.apply(torch.from_numpy(rng.random((40, 50)).astype(np.float32)).cuda(), "reflect")
sg.SavgolFilter(9, 3, 0).apply(torch.from_numpy(rng.standard_normal(2500).astype(np.float32)).cuda())
sg.set_exact(False)
torch.cuda.synchronize()
print("sanitize smoke done, launches:", sg.launch_count())
