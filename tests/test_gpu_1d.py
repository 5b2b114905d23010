"""GPU parity tests of the 1D path: CUDA kernels called through the C ABI vs the oracle.

Two bars (north_star): default (FMA) arithmetic within  max|d| <= 1e-6 * max|x| / dt^d  of the
reference; `exact` arithmetic (reference summation order, unfused) bit-identical."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import savgol_b200 as sg  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


from tolerance import parity_tol  # noqa: E402


def tol(x, dt, d):
    # north_star: max |delta| <= 1e-6 * max|x| * (1/dt^d)   (tests/tolerance.py)
    return parity_tol(x, 1.0 / (dt ** d))


@pytest.fixture(autouse=True)
def _fast_mode():
    sg.set_exact(False)
    yield
    sg.set_exact(False)


MODES = ["polynomial", "reflect", "periodic", "constant"]
CASES = [(1, 1, 0, 1.0), (2, 2, 1, 0.5), (3, 2, 2, 1.0), (5, 3, 0, 1.0), (6, 3, 0, 1.0), (7, 5, 3, 1.0), (10, 2, 1, 0.1),
         (12, 4, 0, 1.0), (13, 4, 1, 1.0), (16, 3, 1, 0.01), (21, 6, 2, 1.0), (27, 3, 0, 1.0), (30, 5, 4, 2.0),
         (31, 4, 1, 1.0), (32, 4, 2, 0.25)]


@pytest.mark.parametrize("n,m,d,dt", CASES)
def test_single_signal_all_modes_tolerance_and_exact(oracle, n, m, d, dt):
    rng = np.random.default_rng(1000 * n + 10 * m + d)
    for L in (2 * n + 1, 2 * n + 2, 100 + 3 * n, 4096, 4097, 4096 + n, 9000, 12289):
        x = rng.standard_normal(L).astype(np.float32)
        xd = torch.from_numpy(x).cuda()
        for mode in MODES:
            ref = oracle.Filter1D(n, m, d, dt, mode).apply(x)
            f = sg.SavgolFilter(n, m, d, dt, mode)
            y = f.apply(xd).cpu().numpy()
            err = np.max(np.abs(y - ref))
            assert err <= tol(x, dt, d), (n, m, d, L, mode, err, tol(x, dt, d))
            sg.set_exact(True)
            ye = f.apply(xd).cpu().numpy()
            sg.set_exact(False)
            assert np.array_equal(bits(ye), bits(ref)), (n, m, d, L, mode, int(np.argmax(bits(ye) != bits(ref))))
            f.close()


def test_config2_shape_batch_reflect(oracle):
    # BASELINE config 2 at reduced batch: signals x 4096, n16 m3 d1 REFLECT
    rng = np.random.default_rng(1)
    rows, L = 512, 4096
    t = np.arange(L, dtype=np.float32)
    x = (rng.standard_normal((rows, L)) + np.sin(0.01 * t)[None, :] * rng.uniform(0.5, 2, (rows, 1))).astype(np.float32)
    f = sg.SavgolFilter(16, 3, 1, 1.0, "reflect")
    ref = oracle.Filter1D(16, 3, 1, 1.0, "reflect").apply(x)
    y = f.apply(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.max(np.abs(y - ref)) <= tol(x, 1.0, 1)
    sg.set_exact(True)
    ye = f.apply(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(bits(ye), bits(ref))


def test_config1_shape_polynomial_1m(oracle):
    # BASELINE config 1: one 1,000,000-sample signal, n12 m4 d0, polynomial edges
    rng = np.random.default_rng(0)
    L = 1_000_000
    x = (np.sin(1e-3 * np.arange(L)) + 0.1 * (rng.random(L) - 0.5)).astype(np.float32)
    f = sg.SavgolFilter(12, 4, 0, 1.0, "polynomial")
    ref = oracle.Filter1D(12, 4, 0, 1.0, "polynomial").apply(x)
    y = f.apply(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.max(np.abs(y - ref)) <= tol(x, 1.0, 0)
    sg.set_exact(True)
    assert np.array_equal(bits(f.apply(torch.from_numpy(x).cuda()).cpu().numpy()), bits(ref))


def test_matlab_known_answer_on_gpu(golden_dir):
    import os
    z = np.load(os.path.join(golden_dir, "matlab_n6_m3.npz"))
    f = sg.SavgolFilter(6, 3, 0, 1.0, "polynomial")
    y = f.apply(torch.from_numpy(z["raw"]).cuda()).cpu().numpy()
    assert np.max(np.abs(y - z["expected"])) < 1e-5
    yh = f.apply(z["raw"].copy())  # host pointer path
    assert np.array_equal(bits(yh), bits(y))


def test_golden_reference_outputs_exact(golden_dir):
    import os
    G = np.load(os.path.join(golden_dir, "ref_outputs.npz"))
    sg.set_exact(True)
    for ci, (n, m, d, dt) in enumerate(G["cases"]):
        x = torch.from_numpy(G[f"c{ci}_x"]).cuda()
        for b in range(4):
            f = sg.SavgolFilter(int(n), int(m), int(d), float(dt), b)
            assert np.array_equal(bits(f.apply(x).cpu().numpy()), bits(G[f"c{ci}_apply_b{b}"])), (ci, b)
            if b == 0:
                assert np.array_equal(bits(f.apply_valid(x).cpu().numpy()), bits(G[f"c{ci}_valid"])), ci
            f.close()


def test_pitched_batch_and_short_rows(oracle):
    rng = np.random.default_rng(5)
    f = sg.SavgolFilter(4, 2, 0, 1.0, "constant")
    o = oracle.Filter1D(4, 2, 0, 1.0, "constant")
    big = torch.from_numpy(rng.standard_normal((37, 301)).astype(np.float32)).cuda()
    view = big[:, 3:250]  # pitch 301, length 247, misaligned start
    out = torch.zeros_like(big)
    f.apply(view, out=out[:, 5:252])
    ref = o.apply(view.cpu().numpy().copy())
    assert np.max(np.abs(out[:, 5:252].cpu().numpy() - ref)) <= tol(view.cpu().numpy(), 1.0, 0)
    assert torch.all(out[:, :5] == 0) and torch.all(out[:, 252:] == 0)


def test_error_returns_like_reference(capfd):
    f = sg.SavgolFilter(5, 2)
    lib = sg.lib()
    x = torch.zeros(100, device="cuda")
    assert lib.savgol_apply(f.handle, x.data_ptr(), x.data_ptr(), 10) == -1      # ref: src/savgolFilter.c:751-755
    assert lib.savgol_apply(f.handle, None, x.data_ptr(), 100) == -1              # ref: :746-749
    assert lib.savgol_apply_valid(f.handle, x.data_ptr(), 10, x.data_ptr()) == 0  # ref: :833-835
    assert lib.savgol_apply_strided(f.handle, x.data_ptr(), 4, 0, x.data_ptr(), 4, 0, 10) == -1
    err = capfd.readouterr().err
    assert "NULL pointer" in err and "window size" in err


def test_in_place_gives_out_of_place_result(oracle):
    # DESIGN.md "in-place": alias-safe, equals the out-of-place result
    rng = np.random.default_rng(9)
    for L in (500, 4096, 20000):
        x = rng.standard_normal(L).astype(np.float32)
        for mode in ("polynomial", "reflect"):
            f = sg.SavgolFilter(8, 3, 0, 1.0, mode)
            ref = oracle.Filter1D(8, 3, 0, 1.0, mode).apply(x)
            xd = torch.from_numpy(x).cuda()
            f.apply(xd, out=xd)
            assert np.max(np.abs(xd.cpu().numpy() - ref)) <= tol(x, 1.0, 0), (L, mode)


def test_strided_struct_field(oracle):
    # ref: test/iterative/test_savgol.c:245-294 -- 12-byte records, float field at offset 4
    rng = np.random.default_rng(11)
    for L in (11, 360, 5000):
        x = rng.standard_normal(L).astype(np.float32)
        rec = np.zeros(3 * L, np.float32); rec[1::3] = x
        f = sg.SavgolFilter(5, 2, 1, 0.5, "reflect")      # boundary is ignored by strided (Q3)
        o = oracle.Filter1D(5, 2, 1, 0.5, "polynomial")
        ref = o.apply(x)
        # device records
        rin = torch.from_numpy(rec).cuda()
        rout = torch.full((3 * L,), -7.0, device="cuda")
        assert f.apply_strided(rin.data_ptr(), 12, 4, rout.data_ptr(), 12, 4, L) == 0
        got = rout.cpu().numpy()
        assert np.max(np.abs(got[1::3] - ref)) <= tol(x, 0.5, 1)
        assert np.all(got[0::3] == -7.0) and np.all(got[2::3] == -7.0)
        # host records
        hout = np.full(3 * L, -7.0, np.float32)
        assert f.apply_strided(rec.ctypes.data, 12, 4, hout.ctypes.data, 12, 4, L) == 0
        assert np.array_equal(bits(hout), bits(got))
        # exact
        sg.set_exact(True)
        rout.fill_(-7.0)
        assert f.apply_strided(rin.data_ptr(), 12, 4, rout.data_ptr(), 12, 4, L) == 0
        assert np.array_equal(bits(rout.cpu().numpy()[1::3]), bits(ref))
        sg.set_exact(False)


def test_halo_slices_reassemble_full_signal(oracle):
    # the per-GPU piece of a partitioned long signal (BASELINE config 3 shape: n32 m4 d2 periodic)
    rng = np.random.default_rng(3)
    n, L, parts = 32, 40000, 5
    x = rng.standard_normal(L).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    for mode in MODES:
        f = sg.SavgolFilter(n, 4, 2, 1.0, mode)
        ref = oracle.Filter1D(n, 4, 2, 1.0, mode).apply(x)
        sg.set_exact(True)
        cuts = [0, 7000, 16001, 16100, 30000, L]
        out = torch.empty_like(xd)
        for p in range(parts):
            a, b = cuts[p], cuts[p + 1]
            if mode == "periodic":
                left = xd[a - n:a] if a > 0 else xd[L - n:]
                right = xd[b:b + n] if b < L else xd[:n]
            else:
                left = xd[a - n:a] if a > 0 else None
                right = xd[b:b + n] if b < L else None
            f.apply_halo(xd[a:b], left.contiguous() if left is not None else None,
                         right.contiguous() if right is not None else None, out=out[a:b])
        sg.set_exact(False)
        assert np.array_equal(bits(out.cpu().numpy()), bits(ref)), mode


def test_host_pointer_path_chunks(oracle):
    rng = np.random.default_rng(4)
    f = sg.SavgolFilter(16, 3, 1, 1.0, "reflect")
    o = oracle.Filter1D(16, 3, 1, 1.0, "reflect")
    x = rng.standard_normal((300, 4096)).astype(np.float32)
    y = f.apply(x)
    assert isinstance(y, np.ndarray)
    assert np.max(np.abs(y - o.apply(x))) <= tol(x, 1.0, 1)
    # pinned host memory
    xp = torch.from_numpy(x).pin_memory()
    yp = torch.empty_like(xp).pin_memory()
    f.apply(xp, out=yp)
    assert np.array_equal(bits(yp.numpy()), bits(y))


def test_signal_longer_than_2_31_samples_periodic(oracle):
    """BASELINE config 3 scale: the reference's padded modes overflow `int` here (SIGFPE at 2^32,
    SURVEY.md Q5).  Checked through the Q6 identity (periodic == VALID over the wrap-padded signal)
    on both ends, around the 2^31 index and on random interior windows."""
    free, _ = torch.cuda.mem_get_info()
    L = (1 << 31) + 12345
    if free < 2 * 4 * L + (2 << 30):
        pytest.skip("not enough device memory")
    n, m, d = 32, 4, 2
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    x = torch.empty(L, device="cuda", dtype=torch.float32)
    step = 1 << 28
    for a in range(0, L, step):
        x[a:a + step].normal_(generator=g)
    f = sg.SavgolFilter(n, m, d, 1.0, "periodic")
    y = f.apply(x)
    o = oracle.Filter1D(n, m, d, 1.0, "periodic")
    W = 5000
    tol_ = 1e-6 * 6.0
    head = torch.cat([x[L - n:], x[:W + n]]).cpu().numpy()
    tail = torch.cat([x[L - W - n:], x[:n]]).cpu().numpy()
    assert np.max(np.abs(o.apply_valid(head) - y[:W].cpu().numpy())) <= tol_
    assert np.max(np.abs(o.apply_valid(tail) - y[L - W:].cpu().numpy())) <= tol_
    rng = np.random.default_rng(0)
    for c in [1 << 31, (1 << 31) - 1000, 1 << 30] + [int(v) for v in rng.integers(W, L - 2 * W, 6)]:
        seg = x[c - n:c + W + n].cpu().numpy()
        assert np.max(np.abs(o.apply_valid(seg) - y[c:c + W].cpu().numpy())) <= tol_, c


@pytest.mark.parametrize("n,m,d,dt", [(1, 1, 0, 1.0), (4, 2, 1, 0.5), (12, 4, 0, 1.0), (16, 3, 1, 1.0), (25, 4, 2, 1.0), (32, 5, 0, 1.0)])
def test_short_row_batches_share_warps(oracle, n, m, d, dt):
    # rows of <= 512 samples run in the packed kernel (2 ... 32 signals per warp, sg1d_packed.cuh):
    # every packing width, ragged row counts, aligned / odd pitches, offset views
    rng = np.random.default_rng(77 + n)
    ws = 2 * n + 1
    lengths = sorted({ws, ws + 1, max(ws, 31), max(ws, 32), max(ws, 33), max(ws, 64), 65 + n, 100 + n, 128, 129, 200, 256, 257, 360, 511, 512})
    for L in lengths:
        for rows in (2, 3, 9, 64, 333):
            pitch = L + (0 if rows % 2 == 0 else 3)
            big = torch.from_numpy(rng.standard_normal((rows, pitch + 1)).astype(np.float32)).cuda()
            off = 1 if rows == 9 else 0
            x = big[:, off:off + L]
            xh = x.cpu().numpy().copy()
            mode = MODES[(L + rows) % 4]
            f = sg.SavgolFilter(n, m, d, dt, mode)
            ref = oracle.Filter1D(n, m, d, dt, mode).apply(xh)
            out = torch.full((rows, pitch + 1), 7.0, device="cuda")
            f.apply(x, out=out[:, off:off + L])
            y = out[:, off:off + L].cpu().numpy()
            assert np.max(np.abs(y - ref)) <= tol(xh, dt, d), (n, L, rows, mode, float(np.max(np.abs(y - ref))))
            assert torch.all(out[:, :off] == 7.0) and torch.all(out[:, off + L:] == 7.0), (n, L, rows)
            f.close()


@pytest.mark.parametrize("n,m,d", [(16, 3, 1), (3, 2, 0), (32, 4, 2)])
def test_misaligned_rows_and_short_tails(oracle, n, m, d):
    """Rows that are not 16-byte aligned are cut on a per-row phase (segments start up to 3 outputs before the row),
    and rows that end 1..32 outputs behind a full segment hand those to the segment before (sg1d_kernel.cuh).
    Every phase of input and output, every boundary mode, lengths around the tail thresholds; both flavours."""
    rng = np.random.default_rng(77 + n)
    lengths = [1025, 1027, 1056, 1057, 2049, 2080, 2081, 3071, 4095, 4096, 4097, 4099, 4127, 4130]
    for L in lengths:
        rows = 5
        for off_in, off_out, extra in [(0, 0, 0), (1, 1, 0), (2, 2, 1), (3, 3, 2), (1, 2, 3), (3, 0, 1), (0, 1, 0)]:
            pitch = L + 3 + extra                          # odd pitches: the phase changes from row to row
            mode = MODES[(L + off_in + extra) % 4]
            big = rng.standard_normal((rows, pitch)).astype(np.float32)
            x = big[:, off_in:off_in + L]
            dbig = torch.from_numpy(big).cuda()
            o = oracle.Filter1D(n, m, d, 1.0, mode)
            f = sg.SavgolFilter(n, m, d, 1.0, mode)
            ref = o.apply(np.ascontiguousarray(x))
            out = torch.full((rows, pitch), 7.0, device="cuda")
            f.apply(dbig[:, off_in:off_in + L], out=out[:, off_out:off_out + L])
            got = out[:, off_out:off_out + L].cpu().numpy()
            assert np.max(np.abs(got - ref)) <= parity_tol(x, 1.0), (L, off_in, off_out, pitch, mode)
            assert torch.all(out[:, :off_out] == 7.0) and torch.all(out[:, off_out + L:] == 7.0), (L, off_in, off_out)
            sg.set_exact(True)
            out.fill_(7.0)
            f.apply(dbig[:, off_in:off_in + L], out=out[:, off_out:off_out + L])
            sg.set_exact(False)
            assert np.array_equal(bits(out[:, off_out:off_out + L].cpu().numpy()), bits(ref)), (L, off_in, off_out, pitch, mode)
            assert torch.all(out[:, :off_out] == 7.0) and torch.all(out[:, off_out + L:] == 7.0)
            v = f.apply_valid(dbig[0, off_in:off_in + L])
            sg.set_exact(True)
            ve = f.apply_valid(dbig[0, off_in:off_in + L]).cpu().numpy()
            sg.set_exact(False)
            rv = o.apply_valid(np.ascontiguousarray(x[0]))
            assert np.array_equal(bits(ve), bits(rv)) and np.max(np.abs(v.cpu().numpy() - rv)) <= parity_tol(x, 1.0)
            f.close()
    # the tail outputs use the operation order of the main loop: the interior of the fast result does not depend on
    # where the row ends (4097 = four segments + a tail of 1; 5000 = five segments)
    f = sg.SavgolFilter(n, m, d, 1.0, "reflect")
    x = torch.from_numpy(rng.standard_normal(5000).astype(np.float32)).cuda()
    long = f.apply(x).cpu().numpy()
    for L in (4097, 4100, 4128):
        for off in (0, 1, 2, 3):
            short = f.apply(x[off:off + L]).cpu().numpy()
            assert np.array_equal(bits(short[n:L - n]), bits(long[off + n:off + L - n])), (L, off)


@pytest.mark.parametrize("n,m,d", [(2, 2, 0), (10, 2, 1), (16, 3, 1), (32, 5, 0)])
def test_short_misaligned_rows_in_the_packed_kernel(oracle, n, m, d):
    """Short rows (several per warp, sg1d_packed.cuh) that are not 16-byte aligned: their slot starts on the row's
    phase (up to 3 outputs early) so that the 16-byte copies and stores stay aligned; the pad elements of a slot come
    from a per-CTA table.  Lengths around every packing class, every phase of input and output, every mode."""
    rng = np.random.default_rng(500 + n)
    ws = 2 * n + 1
    for L in sorted({ws, ws + 1, 29, 30, 33, 48, 61, 62, 63, 64, 65, 100, 125, 126, 127, 128, 129, 250, 253, 254, 255, 256, 360, 509, 510, 511, 512}):
        if L < ws:
            continue
        rows = 37
        for off_in, off_out, extra in [(0, 0, 0), (1, 1, 0), (2, 2, 1), (3, 3, 2), (1, 3, 1), (0, 2, 3)]:
            pitch = L + 3 + extra
            mode = MODES[(L + off_in + extra) % 4]
            big = rng.standard_normal((rows, pitch)).astype(np.float32)
            x = big[:, off_in:off_in + L]
            dbig = torch.from_numpy(big).cuda()
            o = oracle.Filter1D(n, m, d, 1.0, mode)
            f = sg.SavgolFilter(n, m, d, 1.0, mode)
            ref = o.apply(np.ascontiguousarray(x))
            out = torch.full((rows, pitch), 7.0, device="cuda")
            f.apply(dbig[:, off_in:off_in + L], out=out[:, off_out:off_out + L])
            got = out[:, off_out:off_out + L].cpu().numpy()
            assert np.max(np.abs(got - ref)) <= parity_tol(x, 1.0), (L, off_in, off_out, pitch, mode)
            assert torch.all(out[:, :off_out] == 7.0) and torch.all(out[:, off_out + L:] == 7.0), (L, off_in, off_out)
            f.close()
    # the multichannel stream with short misaligned chunks (carried state = left halo of the slot)
    C_, K = 41, 77
    sig = rng.standard_normal((C_, 6 * K + 3)).astype(np.float32)
    dsig = torch.from_numpy(sig).cuda()
    o = oracle.Filter1D(n, m, d, 1.0)
    want = np.stack([o.stream_run(r[1:1 + 6 * K]) for r in sig])
    st = sg.SavgolMCStream(C_, n, m, d, 1.0)
    parts = []
    for k in range(6):
        out, cnt = st.push(dsig[:, 1 + k * K:1 + (k + 1) * K])
        parts.append(out[:, :cnt].cpu().numpy())
    out, cnt = st.flush(dsig)
    parts.append(out[:, :cnt].cpu().numpy())
    got = np.concatenate(parts, axis=1)
    assert got.shape == want.shape and np.max(np.abs(got - want)) <= parity_tol(sig, 1.0)


@pytest.mark.parametrize("n,m,d", [(10, 2, 1), (32, 4, 2)])
def test_stream_chunks_with_short_tails_and_misaligned_views(oracle, n, m, d):
    """Multichannel stream chunks whose length ends just behind a full segment (tail folded into the last segment,
    carried state written from the staged samples) and chunks that are misaligned views: concatenated outputs equal the
    scalar stream of the oracle."""
    rng = np.random.default_rng(900 + n)
    C_ = 19
    for chunks, off in (((1025, 1040, 1056, 1057, 2080), 0), ((1030, 1031, 2049, 1024), 1), ((4097, 1056), 3)):
        total = sum(chunks)
        sig = rng.standard_normal((C_, total + 8)).astype(np.float32)
        dsig = torch.from_numpy(sig).cuda()
        o = oracle.Filter1D(n, m, d, 1.0)
        want = np.stack([o.stream_run(r[off:off + total]) for r in sig])
        st = sg.SavgolMCStream(C_, n, m, d, 1.0)
        parts, pos = [], off
        for K in chunks:
            out, cnt = st.push(dsig[:, pos:pos + K])
            parts.append(out[:, :cnt].cpu().numpy())
            pos += K
        out, cnt = st.flush(dsig)
        parts.append(out[:, :cnt].cpu().numpy())
        got = np.concatenate(parts, axis=1)
        assert got.shape == want.shape, (chunks, off)
        assert np.max(np.abs(got - want)) <= parity_tol(sig, 1.0), (chunks, off)
        st.close()


def test_halo_slices_fast_flavour_misaligned_with_tails(oracle):
    """Default arithmetic: slices of a long signal with explicit halos, cut so that the slices are misaligned and end
    just behind a full segment (per-row phase + tail with lhalo / rhalo), reassemble the whole-signal result bit for
    bit and match the oracle."""
    rng = np.random.default_rng(13)
    L = 30011
    x = rng.standard_normal(L).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    for n, m, d in ((16, 3, 1), (5, 2, 0), (32, 4, 2)):
        for mode in MODES:
            f = sg.SavgolFilter(n, m, d, 1.0, mode)
            ref = oracle.Filter1D(n, m, d, 1.0, mode).apply(x)
            whole = f.apply(xd)
            cuts = [0, 4097, 4097 + 1031, 4097 + 1031 + 2061, 12001, 12001 + 1056, 20003, L]
            out = torch.full_like(xd, 7.0)
            for a, b in zip(cuts[:-1], cuts[1:]):
                if mode == "periodic":
                    left = xd[a - n:a] if a > 0 else xd[L - n:]
                    right = xd[b:b + n] if b < L else xd[:n]
                else:
                    left = xd[a - n:a] if a > 0 else None
                    right = xd[b:b + n] if b < L else None
                f.apply_halo(xd[a:b], left.contiguous() if left is not None else None,
                             right.contiguous() if right is not None else None, out=out[a:b])
            assert np.max(np.abs(out.cpu().numpy() - ref)) <= parity_tol(x, 1.0), (n, mode)
            assert torch.equal(out.view(torch.int32), whole.view(torch.int32)), (n, mode)
            f.close()
