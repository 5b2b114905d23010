"""Pins oracle/savgol_oracle.c (the CPU restatement) against committed golden vectors:
the reference repo's MATLAB known-answer vector and outputs of the unmodified reference
(tests/golden/make_golden.py).  Bit-exact unless stated."""
import os

import numpy as np
import pytest


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "ref_outputs.npz"))


def test_matlab_known_answer(oracle, golden_dir):
    # savgolComparison.m:2,5,7-9 -> window 13 (n=6), order 3, smoothing, polynomial edges.
    z = np.load(os.path.join(golden_dir, "matlab_n6_m3.npz"))
    y = oracle.Filter1D(6, 3, 0, 1.0, "polynomial").apply(z["raw"])
    # the expected vector is printed with 6 decimals at magnitude <= 39
    assert np.max(np.abs(y - z["expected"])) < 1e-5


def test_weights_bit_exact(oracle, G):
    for ci, (n, m, d, dt) in enumerate(G["cases"]):
        f = oracle.Filter1D(int(n), int(m), int(d), float(dt))
        assert np.array_equal(bits(f.center), bits(G[f"c{ci}_center"])), (n, m, d)
        assert np.array_equal(bits(f.edge), bits(G[f"c{ci}_edge"])), (n, m, d)


@pytest.mark.parametrize("b", [0, 1, 2, 3])
def test_apply_modes_bit_exact(oracle, G, b):
    for ci, (n, m, d, dt) in enumerate(G["cases"]):
        f = oracle.Filter1D(int(n), int(m), int(d), float(dt), b)
        y = f.apply(G[f"c{ci}_x"])
        assert np.array_equal(bits(y), bits(G[f"c{ci}_apply_b{b}"])), (ci, b)


def test_valid_strided_stream_bit_exact(oracle, G):
    for ci, (n, m, d, dt) in enumerate(G["cases"]):
        f = oracle.Filter1D(int(n), int(m), int(d), float(dt))
        x = G[f"c{ci}_x"]
        assert np.array_equal(bits(f.apply_valid(x)), bits(G[f"c{ci}_valid"]))
        L = x.size
        rin = np.zeros(L * 3, np.float32); rin[1::3] = x
        rout = np.full(L * 3, -7.0, np.float32)
        assert f.apply_strided(rin, 12, 4, rout, 12, 4, L) == 0
        assert np.array_equal(bits(rout), bits(G[f"c{ci}_strided"]))
        ys = f.stream_run(x)
        assert ys.size == L
        assert np.array_equal(bits(ys), bits(G[f"c{ci}_stream"]))


def test_2d_bit_exact(oracle, G):
    for ci, (nx, ny, o, dx, dy) in enumerate(G["cases2d"]):
        f = oracle.Filter2D(int(nx), int(ny), int(o), int(dx), int(dy), 0.5, 2.0)
        assert np.array_equal(bits(f.W.ravel()), bits(G[f"d{ci}_W"])), ci
        assert np.float32(f.scale) == G[f"d{ci}_scale"]
        img = G[f"d{ci}_img"]
        for b in range(3):
            out = np.full(img.shape, -3.0, np.float32)
            f.apply(img, b, out)
            assert np.array_equal(bits(out), bits(G[f"d{ci}_apply_b{b}"])), (ci, b)


def test_error_returns(oracle):
    # ref: src/savgolFilter.c:751-755, 833-835 -- too-short input
    f = oracle.Filter1D(5, 2)
    with pytest.raises(ValueError):
        f.apply(np.zeros(10, np.float32))
    assert f.apply_valid(np.zeros(10, np.float32)).size == 0
    for bad in [(0, 2, 0), (2, 10, 0), (5, 2, 3), (33, 2, 0), (5, 4, 5)]:
        with pytest.raises(ValueError):
            oracle.Filter1D(*bad)
    with pytest.raises(ValueError):
        oracle.Filter1D(5, 2, 0, 0.0)


def test_q1_leading_edge_sign(oracle):
    # ramp 3i+1, n=5 m=2 d=1 -> leading edge -3, rest +3 (reference behaviour).
    x = (3.0 * np.arange(40) + 1.0).astype(np.float32)
    y = oracle.Filter1D(5, 2, 1, 1.0).apply(x)
    assert np.allclose(y[:5], -3.0, atol=1e-3) and np.allclose(y[5:], 3.0, atol=1e-3)
