#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ (run in the authoring
container, where /root/reference exists).

1. matlab_n6_m3.npz   -- the reference repo's only known-answer vector: `rawData` and
   `yourSavgolData` embedded in "tool for matlab comparisons/savgolComparison.m" lines 2 and 5
   (301 samples, window 13 = half_window 6, order 3, smoothing, polynomial edges).  The
   text has 6 decimals, so the comparison tolerance is 1e-5.
2. demo360.npy        -- the 360-sample dataset of test/iterative/test_savgol_main.c:55-92
   (input only; used as a realistic signal).
3. ref_outputs.npz    -- outputs of the UNMODIFIED reference (oracle/_ref/libsavgol_ref.so,
   gcc -O2 -ffp-contract=off) on seeded inputs for every API on the hot path: apply in
   the four boundary modes, apply_valid, apply_strided, the stream push_full+flush
   sequence, savgol2d apply (valid/constant/reflect) and a table of weight vectors.  These
   let the CPU test-suite pin oracle/savgol_oracle.c even where /root/reference and
   oracle/_ref are absent.
"""
import ctypes as C
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"


def matlab_vectors():
    txt = open(os.path.join(REF, "tool for matlab comparisons", "savgolComparison.m")).read()
    raw = re.search(r"rawData\s*=\s*\[(.*?)\]", txt, re.S).group(1)
    exp = re.search(r"yourSavgolData\s*=\s*\[(.*?)\]", txt, re.S).group(1)
    raw = np.array([float(v) for v in raw.replace(";", ",").split(",") if v.strip()], np.float32)
    exp = np.array([float(v) for v in exp.replace(";", ",").split(",") if v.strip()], np.float32)
    return raw, exp


def demo_dataset():
    txt = open(os.path.join(REF, "test", "iterative", "test_savgol_main.c")).read()
    body = re.search(r"dataset\[\]\s*=\s*\{(.*?)\};", txt, re.S).group(1)
    vals = re.findall(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?f?", body)
    return np.array([float(v.rstrip("f")) for v in vals], np.float32)


def ref_filter(R, n, m, d, dt, b):
    cfg = O.make_config(n, m, d, dt, b)
    f = R.savgol_create(C.byref(cfg))
    assert f
    return f


def main():
    R = O.ref()
    raw, exp = matlab_vectors()
    assert raw.size == exp.size == 301, (raw.size, exp.size)
    np.savez(os.path.join(HERE, "matlab_n6_m3.npz"), raw=raw, expected=exp)
    demo = demo_dataset()
    assert demo.size == 360, demo.size
    np.save(os.path.join(HERE, "demo360.npy"), demo)

    out = {}
    rng = np.random.default_rng(20261017)
    cases = [(1, 1, 0, 1.0), (2, 2, 1, 0.5), (5, 3, 0, 1.0), (6, 3, 0, 1.0), (12, 4, 0, 1.0),
             (16, 3, 1, 0.01), (32, 4, 2, 0.25), (10, 2, 1, 0.1), (32, 10, 4, 2.0), (7, 5, 3, 1.0)]
    out["cases"] = np.array(cases, np.float64)
    for ci, (n, m, d, dt) in enumerate(cases):
        L = 2 * n + 1 + int(rng.integers(0, 200))
        x = rng.standard_normal(L).astype(np.float32)
        out[f"c{ci}_x"] = x
        for b in range(4):
            f = ref_filter(R, n, m, d, dt, b)
            y = np.zeros(L, np.float32)
            assert R.savgol_apply(f, x.ctypes.data, y.ctypes.data, L) == 0
            out[f"c{ci}_apply_b{b}"] = y
            if b == 0:
                out[f"c{ci}_center"] = np.ctypeslib.as_array(f.contents.center_weights).copy()
                out[f"c{ci}_edge"] = np.ctypeslib.as_array(f.contents.edge_weights).reshape(32, 65).copy()
                yv = np.zeros(L - 2 * n, np.float32)
                k = R.savgol_apply_valid(f, x.ctypes.data, L, yv.ctypes.data)
                assert k == L - 2 * n
                out[f"c{ci}_valid"] = yv
                # strided: 12-byte records, field at offset 4 (as test_savgol.c:245-249)
                rec_in = np.zeros(L * 3, np.float32); rec_in[1::3] = x
                rec_out = np.full(L * 3, -7.0, np.float32)
                assert R.savgol_apply_strided(f, rec_in.ctypes.data, 12, 4, rec_out.ctypes.data, 12, 4, L) == 0
                out[f"c{ci}_strided"] = rec_out
                # stream: push_full for every sample, then flush
                st = O.SavgolStream()
                assert R.savgol_stream_init(C.byref(st), f) == 0
                ys, buf = [], (C.c_float * 40)()
                for v in x:
                    k = R.savgol_stream_push_full(C.byref(st), float(v), buf, 40)
                    ys.extend(buf[:k])
                k = R.savgol_stream_flush(C.byref(st), buf, 40)
                ys.extend(buf[:k])
                out[f"c{ci}_stream"] = np.array(ys, np.float32)
            R.savgol_destroy(f)

    cases2d = [(1, 1, 1, 0, 0), (2, 2, 2, 1, 0), (7, 7, 3, 0, 0), (2, 1, 2, 0, 1), (3, 4, 4, 1, 1),
               (7, 7, 3, 2, 0), (16, 16, 6, 0, 0), (5, 3, 6, 2, 2)]
    out["cases2d"] = np.array(cases2d, np.int64)
    for ci, (nx, ny, o, dx, dy) in enumerate(cases2d):
        cfg = O.Savgol2DConfig(nx, ny, o, dx, dy, 0.5, 2.0)
        f = R.savgol2d_create(C.byref(cfg))
        assert f, cases2d[ci]
        area = f.contents.window_area
        out[f"d{ci}_W"] = np.ctypeslib.as_array(f.contents.weights, shape=(area,)).copy()
        out[f"d{ci}_scale"] = np.float32(f.contents.scale)
        rows, cols = 2 * ny + 1 + int(rng.integers(0, 12)), 2 * nx + 1 + int(rng.integers(0, 12))
        img = rng.standard_normal((rows, cols)).astype(np.float32)
        out[f"d{ci}_img"] = img
        for b in range(3):
            y = np.full((rows, cols), -3.0, np.float32)
            assert R.savgol2d_apply(f, img.ctypes.data, rows, cols, cols, y.ctypes.data, cols, b) == 0
            out[f"d{ci}_apply_b{b}"] = y
        R.savgol2d_destroy(f)
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
