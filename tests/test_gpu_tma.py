"""The bulk-tensor (TMA) 1D kernels vs the cp.async kernels and the oracle.

Both kernel families run the same FFMA2 chains in the same order, so for the default flavour their outputs
must be BIT-IDENTICAL; the cp.async family is itself pinned to the oracle elsewhere (tests/test_gpu_1d.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
import savgol_b200 as sg  # noqa: E402


def bits(t):
    return t.view(torch.int32)


@pytest.fixture(autouse=True)
def _restore():
    sg.set_exact(False)
    sg.lib().savgol_b200_set_tma(1)
    yield
    sg.lib().savgol_b200_set_tma(1)


def both(call):
    lib = sg.lib()
    c0 = lib.savgol_b200_tma_launch_count()
    lib.savgol_b200_set_tma(2)   # wherever the layout allows, whatever the size heuristic says
    a = call()
    used = lib.savgol_b200_tma_launch_count() - c0
    lib.savgol_b200_set_tma(0)
    c1 = lib.savgol_b200_tma_launch_count()
    b = call()
    assert lib.savgol_b200_tma_launch_count() == c1
    lib.savgol_b200_set_tma(1)
    return a, b, used


@pytest.mark.parametrize("n,m,d", [(1, 1, 0), (3, 2, 1), (7, 3, 0), (10, 2, 1), (12, 4, 0), (16, 3, 1), (17, 4, 2), (25, 5, 3), (31, 4, 1), (32, 4, 2)])
@pytest.mark.parametrize("mode", ["polynomial", "reflect", "periodic", "constant"])
def test_tma_equals_cp_async_all_modes(oracle, n, m, d, mode):
    g = torch.Generator(device="cuda").manual_seed(100 * n + d)
    f = sg.SavgolFilter(n, m, d, 1.0, mode)
    of = oracle.Filter1D(n, m, d, 1.0, mode)
    for rows, L, pitch in ((1, 1024, 1024), (3, 1024, 1028), (2, 2048, 2048), (5, 4096, 4096), (2, 4100, 4100), (3, 5000, 5004),
                           (1, 1024 + 32, 1056), (2, 1024 + 31 + n, 1024 + 64 + 4 * ((n + 3) // 4)), (1, 12289, 12292), (1, 100000, 100000)):
        buf = torch.randn(rows, pitch, device="cuda", generator=g)
        x = buf[:, :L]
        outbuf = [torch.full((rows, pitch), 7.0, device="cuda") for _ in range(2)]
        k = [0]

        def call():
            o = outbuf[k[0]]
            k[0] += 1
            f.apply(x, out=o[:, :L])
            return o
        a, b, used = both(call)
        assert used >= 1, (n, mode, rows, L, pitch)
        assert torch.equal(bits(a), bits(b)), (n, mode, rows, L, pitch)          # including the untouched pitch padding
        ref = of.apply(x.cpu().numpy()) if rows > 1 else of.apply(x[0].cpu().numpy())[None, :]
        err = float(np.max(np.abs(a[:, :L].cpu().numpy() - ref)))
        l1 = float(np.sum(np.abs(np.asarray(of.center))))
        assert err <= 1e-6 * max(1.0, l1) * float(x.abs().max()), (n, mode, rows, L, err)


def test_tma_misaligned_or_short_rows_take_the_other_kernels():
    lib = sg.lib()
    lib.savgol_b200_set_tma(2)
    f = sg.SavgolFilter(5, 2, 0, 1.0, "reflect")
    c0 = lib.savgol_b200_tma_launch_count()
    f.apply(torch.randn(4, 1000, device="cuda"))                    # shorter than a segment
    f.apply(torch.randn(4, 2049, device="cuda"))                    # odd pitch
    f.apply(torch.randn(4100, device="cuda")[1:])                   # misaligned base
    assert lib.savgol_b200_tma_launch_count() == c0


def test_tma_valid_and_halo_slices(oracle):
    n = 8
    f = sg.SavgolFilter(n, 3, 1, 1.0, "periodic")
    of = oracle.Filter1D(n, 3, 1, 1.0, "periodic")
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(3 * 4096, device="cuda", generator=g)
    whole = f.apply(x)
    ref = of.apply(x.cpu().numpy())
    assert float(np.max(np.abs(whole.cpu().numpy() - ref))) <= 1e-6 * float(x.abs().max())
    # the signal cut into three slices, halos = the neighbours' samples (periodic wrap at the ends)
    parts = []
    for s in range(3):
        sl = x[s * 4096:(s + 1) * 4096]
        left = x[(s * 4096 - n) % x.numel():][:n] if s else x[-n:]
        right = x[((s + 1) * 4096) % x.numel():][:n]
        parts.append(f.apply_halo(sl, left.clone(), right.clone()))
    assert torch.equal(bits(torch.cat(parts)), bits(whole))
    # VALID (n % 4 == 0 keeps the shifted base 16-byte aligned -> TMA kernel with explicit halos)
    sg.lib().savgol_b200_set_tma(2)
    c0 = sg.lib().savgol_b200_tma_launch_count()
    v = f.apply_valid(x)
    assert sg.lib().savgol_b200_tma_launch_count() > c0
    assert torch.equal(bits(v), bits(whole[n:-n])) or float((v - whole[n:-n]).abs().max()) == 0.0


def test_tma_stream_chunks(oracle):
    n, C, K = 10, 64, 1024
    s = sg.SavgolMCStream(C, n, 2, 1, 1.0)
    g = torch.Generator(device="cuda").manual_seed(9)
    sig = torch.randn(C, 4 * K, device="cuda", generator=g)
    lib = sg.lib()
    outs = {}
    for on in (2, 0):
        lib.savgol_b200_set_tma(on)
        s.reset()
        c0 = lib.savgol_b200_tma_launch_count()
        got = []
        for c in range(4):
            # rows of the output buffer must be 16-byte aligned for the bulk-tensor store: pitch K + n rounded up to 4
            o, k = s.push(sig[:, c * K:(c + 1) * K].contiguous(), out=torch.empty(C, (K + n + 3) & ~3, device="cuda"))
            got.append(o[:, :k].clone())
        o, k = s.flush(torch.empty(1, device="cuda"))
        got.append(o[:, :k].clone())
        outs[on] = torch.cat(got, dim=1)
        assert (lib.savgol_b200_tma_launch_count() > c0) == bool(on)
    lib.savgol_b200_set_tma(1)
    assert torch.equal(bits(outs[2]), bits(outs[0]))
    want = np.stack([oracle.Filter1D(n, 2, 1, 1.0).stream_run(r) for r in sig[:8].cpu().numpy()])
    assert float(np.max(np.abs(outs[2][:8].cpu().numpy() - want))) <= 1e-6 * float(sig.abs().max())
