"""GPU parity tests of the multi-channel chunked stream vs the reference's per-channel stream
semantics (oracle: sgo_stream_run == savgol_stream_push_full for every sample, then flush)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import savgol_b200 as sg  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(autouse=True)
def _fast_mode():
    sg.set_exact(False)
    yield
    sg.set_exact(False)


def run_stream(channels, n, m, d, dt, sig, chunks, device=True):
    s = sg.SavgolMCStream(channels, n, m, d, dt)
    outs, pos, counts = [], 0, []
    for K in chunks:
        piece = np.ascontiguousarray(sig[:, pos:pos + K])
        pos += K
        arg = torch.from_numpy(piece).cuda() if device else piece
        o, k = s.push(arg)
        counts.append(k)
        o = o.cpu().numpy() if device else o
        outs.append(o[:, :k])
    like = torch.empty(1, device="cuda") if device else np.empty(1, np.float32)
    o, k = s.flush(like)
    o = o.cpu().numpy() if device else o
    outs.append(o[:, :k])
    assert s.latency == n and s.samples_received == pos
    return np.concatenate(outs, axis=1), counts, s


@pytest.mark.parametrize("n,m,d,dt", [(10, 2, 1, 0.1), (5, 3, 0, 1.0), (1, 1, 0, 1.0), (16, 4, 2, 0.5), (32, 4, 2, 1.0), (3, 2, 1, 1.0)])
def test_chunked_stream_equals_reference_stream(oracle, n, m, d, dt):
    rng = np.random.default_rng(40 + n)
    C = 37
    ws = 2 * n + 1
    patterns = [[1024, 1024, 1024], [ws, 1, 1, 5, 300, 1024, 7], [3, 2, max(1, ws - 6), 4, 2048, 100], [5000], [ws - 1, 1, 64]]
    o = oracle.Filter1D(n, m, d, dt)
    for chunks in patterns:
        total = sum(chunks)
        sig = rng.standard_normal((C, total)).astype(np.float32)
        want = np.stack([o.stream_run(r) for r in sig])
        got, counts, s = run_stream(C, n, m, d, dt, sig, chunks)
        assert got.shape == want.shape == (C, total)         # total outputs == total inputs (ref test :277-304)
        # per-call output counts follow the reference: 0 while filling, received-n at the first fill, K afterwards
        recv = 0
        for K, k in zip(chunks, counts):
            before, recv = recv, recv + K
            assert k == (0 if recv < ws else (recv - n if before < ws else K)), (chunks, K, k)
        tol = 1e-6 * float(np.abs(sig).max()) / dt ** d
        assert np.max(np.abs(got - want)) <= tol, (chunks, float(np.max(np.abs(got - want))), tol)
        sg.set_exact(True)
        gote, _, _ = run_stream(C, n, m, d, dt, sig, chunks)
        sg.set_exact(False)
        assert np.array_equal(bits(gote), bits(want)), chunks


def test_host_chunks_and_reset(oracle):
    rng = np.random.default_rng(2)
    sig = rng.standard_normal((9, 700)).astype(np.float32)
    o = oracle.Filter1D(10, 2, 1, 1.0)
    want = np.stack([o.stream_run(r) for r in sig])
    got, _, s = run_stream(9, 10, 2, 1, 1.0, sig, [300, 400], device=False)
    assert np.max(np.abs(got - want)) <= 1e-6 * float(np.abs(sig).max())
    s.reset()
    assert s.samples_received == 0 and s.samples_output == 0
    out, k = s.push(sig[:, :5].copy())
    assert k == 0
    out, k = s.flush(np.empty(1, np.float32))
    assert k == 0     # flush before the window ever filled (ref: src/savgol_stream.c:238-241)


def test_config5_shape_channels(oracle):
    # BASELINE config 5 at reduced channel count: 4096 channels x 1024-sample chunks, n10 m2 d1
    rng = np.random.default_rng(4)
    C, K = 4096, 1024
    sig = rng.standard_normal((C, 3 * K)).astype(np.float32)
    got, counts, _ = run_stream(C, 10, 2, 1, 1.0, sig, [K, K, K])
    assert counts == [K - 10, K, K]
    o = oracle.Filter1D(10, 2, 1, 1.0)
    pick = [0, 1, 17, 2047, 4095]
    want = np.stack([o.stream_run(sig[c]) for c in pick])
    assert np.max(np.abs(got[pick] - want)) <= 1e-6 * float(np.abs(sig).max())
    # stream == batch with polynomial edges to ~2e-7 (ref: test_savgol_stream.c:140-189)
    f = sg.SavgolFilter(10, 2, 1, 1.0, "polynomial")
    yb = f.apply(torch.from_numpy(sig[pick]).cuda()).cpu().numpy()
    assert np.max(np.abs(yb - got[pick])) <= 1e-6 * float(np.abs(sig).max())


def test_checkpoint_resume_is_bit_identical():
    # save after a few chunks, resume in a fresh stream: same outputs as the uninterrupted stream
    rng = np.random.default_rng(9)
    C_, n = 53, 7
    sig = rng.standard_normal((C_, 3000)).astype(np.float32)
    d = torch.from_numpy(sig).cuda()
    a = sg.SavgolMCStream(C_, n, 3, 1, 0.5)
    outs = []
    for lo, hi in ((0, 5), (5, 700), (700, 1500)):
        o, k = a.push(d[:, lo:hi].contiguous())
        outs.append(o[:, :k].cpu().numpy())
    blob = a.save()
    rest_a = [a.push(d[:, 1500:2200].contiguous()), a.push(d[:, 2200:].contiguous())]
    fa, ka = a.flush(d)
    b = sg.SavgolMCStream(C_, n, 3, 1, 0.5)
    b.restore(blob)
    assert b.samples_received == 1500 and b.samples_output == 1500 - n
    rest_b = [b.push(d[:, 1500:2200].contiguous()), b.push(d[:, 2200:].contiguous())]
    fb, kb = b.flush(d)
    for (oa, ka_), (ob, kb_) in zip(rest_a, rest_b):
        assert ka_ == kb_ and torch.equal(oa[:, :ka_], ob[:, :kb_])
    assert ka == kb == n and torch.equal(fa[:, :n], fb[:, :n])
    # a blob from another configuration is refused
    c = sg.SavgolMCStream(C_, n + 1, 3, 1, 0.5)
    with pytest.raises(ValueError):
        c.restore(blob)
