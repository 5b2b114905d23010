"""Seeded random sweep of the 1D / 2D / stream C ABI against the oracle: random filter parameters, boundary
modes, lengths (around every kernel-internal threshold: window size, 128/256/512-sample packing classes,
1024-sample segments), row counts, pitches and offsets.  Fast arithmetic within the north_star tolerance,
`exact` arithmetic bit-identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import savgol_b200 as sg  # noqa: E402
from tolerance import l1_gain_1d, l1_gain_2d, parity_tol  # noqa: E402

MODES = ["polynomial", "reflect", "periodic", "constant"]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(autouse=True)
def _fast_mode():
    sg.set_exact(False)
    yield
    sg.set_exact(False)


def _random_filter(rng):
    n = int(rng.choice([1, 2, 3, 5, 8, 12, 16, 17, 24, 32]))
    m = int(rng.integers(1, min(6, 2 * n) + 1))
    d = int(rng.integers(0, min(m, 3) + 1))
    dt = float(rng.choice([1.0, 0.5, 0.01, 2.0]))
    return n, m, d, dt


@pytest.mark.parametrize("seed", range(6))
def test_batch_random_sweep(oracle, seed):
    rng = np.random.default_rng(9000 + seed)
    for _ in range(25):
        n, m, d, dt = _random_filter(rng)
        ws = 2 * n + 1
        base = int(rng.choice([ws, 64, 128, 129, 256, 300, 512, 513, 1024, 1025, 2048, 3000, 4096]))
        L = max(ws, base + int(rng.integers(-3, 4)))
        rows = int(rng.choice([1, 2, 3, 5, 8, 9, 31, 64]))
        pitch = L + int(rng.choice([0, 0, 1, 3, 4]))
        off = int(rng.choice([0, 0, 1, 2]))
        mode = MODES[int(rng.integers(4))]
        big = rng.standard_normal((rows, pitch + off)).astype(np.float32)
        x = big[:, off:off + L]
        dbig = torch.from_numpy(big).cuda()
        dx = dbig[:, off:off + L]
        o = oracle.Filter1D(n, m, d, dt, mode)
        f = sg.SavgolFilter(n, m, d, dt, mode)
        ref = o.apply(np.ascontiguousarray(x))
        out = torch.full((rows, pitch + off), 5.0, device="cuda")
        f.apply(dx, out=out[:, off:off + L])
        got = out[:, off:off + L].cpu().numpy()
        tol = parity_tol(x, 1.0 / dt ** d, l1_gain_1d(o))
        assert np.max(np.abs(got - ref)) <= tol, (seed, n, m, d, dt, mode, L, rows, pitch, off)
        assert torch.all(out[:, :off] == 5.0) and torch.all(out[:, off + L:] == 5.0)
        sg.set_exact(True)
        ge = f.apply(dx).cpu().numpy()
        sg.set_exact(False)
        assert np.array_equal(bits(ge), bits(ref)), (seed, n, m, d, dt, mode, L, rows, pitch, off)
        f.close()


@pytest.mark.parametrize("seed", range(3))
def test_stream_random_sweep(oracle, seed):
    rng = np.random.default_rng(7000 + seed)
    for _ in range(8):
        n, m, d, dt = _random_filter(rng)
        C_ = int(rng.choice([1, 2, 7, 33, 130]))
        chunks = [int(rng.choice([1, 2, 5, 2 * n, 2 * n + 1, 64, 128, 200, 512, 1024, 1500])) for _ in range(int(rng.integers(2, 7)))]
        if sum(chunks) < 2 * n + 1:
            chunks.append(2 * n + 1)
        total = sum(chunks)
        sig = rng.standard_normal((C_, total)).astype(np.float32)
        o = oracle.Filter1D(n, m, d, dt)
        want = np.stack([o.stream_run(r) for r in sig])
        d_sig = torch.from_numpy(sig).cuda()
        for exact in (False, True):
            sg.set_exact(exact)
            s = sg.SavgolMCStream(C_, n, m, d, dt)
            outs, pos = [], 0
            for K in chunks:
                oo, k = s.push(d_sig[:, pos:pos + K].contiguous())
                outs.append(oo[:, :k].cpu().numpy())
                pos += K
            oo, k = s.flush(d_sig)
            outs.append(oo[:, :k].cpu().numpy())
            s.close()
            got = np.concatenate(outs, axis=1)
            sg.set_exact(False)
            assert got.shape == want.shape, (seed, n, chunks)
            if exact:
                assert np.array_equal(bits(got), bits(want)), (seed, n, m, d, chunks, C_)
            else:
                assert np.max(np.abs(got - want)) <= parity_tol(sig, 1.0 / dt ** d, l1_gain_1d(o)), (seed, n, m, d, chunks, C_)


@pytest.mark.parametrize("seed", range(3))
def test_2d_random_sweep(oracle, seed):
    rng = np.random.default_rng(5000 + seed)
    for _ in range(8):
        nx, ny = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        if rng.random() < 0.5:
            ny = nx
        order = int(rng.integers(1, min(6, 2 * min(nx, ny)) + 1))
        dx = int(rng.integers(0, min(order, 2) + 1))
        dy = int(rng.integers(0, min(order - dx, 2) + 1))
        rows = int(rng.integers(2 * ny + 1, 2 * ny + 400))
        cols = int(rng.choice([2 * nx + 1 + int(rng.integers(0, 8)), 64, 128, 131, 256, 260, 300, 512]))
        cols = max(cols, 2 * nx + 1)
        images = int(rng.choice([1, 1, 2, 3]))
        x = rng.standard_normal((images, rows, cols)).astype(np.float32)
        d_x = torch.from_numpy(x).cuda()
        o = oracle.Filter2D(nx, ny, order, dx, dy)
        f = sg.Savgol2DFilter(nx, ny, order, dx, dy)
        b = ["valid", "constant", "reflect"][int(rng.integers(3))]
        out = torch.full(x.shape, -3.0, device="cuda")
        f.apply(d_x, b, out=out)
        got = out.cpu().numpy()
        tol = parity_tol(x, o.scale, l1_gain_2d(o))
        for i in range(images):
            ref = np.full((rows, cols), -3.0, np.float32)
            o.apply(x[i], b, ref)
            assert np.max(np.abs(got[i] - ref)) <= tol, (seed, nx, ny, order, dx, dy, rows, cols, b, i)
