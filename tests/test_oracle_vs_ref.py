"""Differential test: oracle/savgol_oracle.c vs the UNMODIFIED reference compiled from
/root/reference (oracle/_ref/libsavgol_ref.so).  Skipped where the .so is absent."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_all_weight_tables_bit_exact():
    R = O.ref()
    cnt = 0
    for n in range(1, 33):
        for m in range(0, 11):
            for d in range(0, 5):
                cfg = O.make_config(n, m, d, 0.5, 0)
                f = R.savgol_create(C.byref(cfg))
                ok = (m < 2 * n + 1) and d <= m
                assert bool(f) == ok, (n, m, d)
                if not f:
                    with pytest.raises(ValueError):
                        O.Filter1D(n, m, d, 0.5)
                    continue
                of = O.Filter1D(n, m, d, 0.5)
                cw = np.ctypeslib.as_array(f.contents.center_weights)
                ew = np.ctypeslib.as_array(f.contents.edge_weights).reshape(32, 65)
                assert np.array_equal(bits(cw), bits(of.center)), (n, m, d)
                assert np.array_equal(bits(ew), bits(of.edge)), (n, m, d)
                assert np.float32(1.0) / np.float32(f.contents.dt_scale) == np.float32(of.dt_inv)
                R.savgol_destroy(f)
                cnt += 1
    assert cnt == 1341


@pytest.mark.parametrize("n,m,d", [(1, 0, 0), (3, 2, 2), (12, 4, 0), (16, 3, 1), (32, 4, 2), (10, 2, 1)])
def test_apply_random_bit_exact(n, m, d):
    R = O.ref()
    rng = np.random.default_rng(n * 100 + m * 10 + d)
    for L in (2 * n + 1, 2 * n + 2, 4 * n + 3, 1000):
        x = rng.standard_normal(L).astype(np.float32)
        for b in range(4):
            cfg = O.make_config(n, m, d, 0.7, b)
            f = R.savgol_create(C.byref(cfg))
            y = np.zeros(L, np.float32)
            assert R.savgol_apply(f, x.ctypes.data, y.ctypes.data, L) == 0
            yo = O.Filter1D(n, m, d, 0.7, b).apply(x)
            assert np.array_equal(bits(y), bits(yo)), (n, m, d, L, b)
            R.savgol_destroy(f)


def test_q6_padded_modes_equal_valid_over_numpy_pad():
    # the identity the 2^32-sample config relies on.
    rng = np.random.default_rng(6)
    for n, L in ((5, 11), (16, 300), (32, 4096)):
        x = rng.standard_normal(L).astype(np.float32)
        for mode, pad in ((1, "symmetric"), (2, "wrap"), (3, "edge")):
            f = O.Filter1D(n, 3, 1, 1.0, mode)
            y = f.apply(x)
            yv = f.apply_valid(np.pad(x, n, pad))
            assert np.array_equal(bits(y), bits(yv))


def test_2d_random_bit_exact():
    R = O.ref()
    rng = np.random.default_rng(2)
    for nx, ny, o, dx, dy in [(2, 2, 2, 0, 0), (7, 7, 3, 0, 0), (3, 5, 4, 1, 2), (16, 16, 6, 0, 0)]:
        cfg = O.Savgol2DConfig(nx, ny, o, dx, dy, 1.5, 0.5)
        f = R.savgol2d_create(C.byref(cfg))
        of = O.Filter2D(nx, ny, o, dx, dy, 1.5, 0.5)
        W = np.ctypeslib.as_array(f.contents.weights, shape=(f.contents.window_area,))
        assert np.array_equal(bits(W), bits(of.W.ravel()))
        img = rng.standard_normal((2 * ny + 9, 2 * nx + 14)).astype(np.float32)
        for b in range(3):
            y = np.full(img.shape, 5.0, np.float32)
            yo = np.full(img.shape, 5.0, np.float32)
            assert R.savgol2d_apply(f, img.ctypes.data, img.shape[0], img.shape[1], img.shape[1],
                                    y.ctypes.data, img.shape[1], b) == 0
            of.apply(img, b, yo)
            assert np.array_equal(bits(y), bits(yo)), (nx, ny, o, dx, dy, b)
        R.savgol2d_destroy(f)


def test_harness_rows_equals_loop():
    R = O.ref()
    L = O.lib()
    rng = np.random.default_rng(3)
    x = rng.standard_normal((37, 257)).astype(np.float32)
    cfg = O.make_config(16, 3, 1, 1.0, 1)
    f = R.savgol_create(C.byref(cfg))
    y = np.zeros_like(x)
    rc = L.sgh_apply_rows(O.fnptr(R, "savgol_apply"), C.cast(f, C.c_void_p), O._fp(x), O._fp(y),
                          37, 257, 257, 257, 4)
    assert rc == 0
    yo = O.Filter1D(16, 3, 1, 1.0, 1).apply(x)
    assert np.array_equal(bits(y), bits(yo))
    R.savgol_destroy(f)
