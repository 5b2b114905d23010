"""Host-buffer staging (pipelines, in place, VALID, stream chunks) and the single-process multi-GPU C ABI.

The staging chunk is shrunk to 1 MiB in a child process (SAVGOL_B200_CHUNK_MIB is read once per process) so that
modest arrays are cut into many pieces / channel blocks."""
import ctypes as C
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import savgol_b200 as sg  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def run_child(code):
    env = dict(os.environ, SAVGOL_B200_CHUNK_MIB="1")
    p = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r)\n" % ROOT + textwrap.dedent(code)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return p.stdout


def test_host_long_signal_in_place_and_valid_many_pieces():
    # with input == output the D2H of one piece used to overwrite the next piece's left halo
    out = run_child("""
        import numpy as np, savgol_b200 as sg
        from oracle import oracle as O
        rng = np.random.default_rng(3)
        L = 5 * (1 << 18) + 12345                      # > 5 staging pieces of 1 MiB
        for n, m, d, mode in [(16, 3, 1, "reflect"), (32, 4, 2, "periodic"), (7, 3, 0, "polynomial"), (5, 2, 0, "constant")]:
            x = rng.standard_normal(L).astype(np.float32)
            f = sg.SavgolFilter(n, m, d, 1.0, mode)
            sg.set_exact(True)
            ref = O.Filter1D(n, m, d, 1.0, mode).apply(x)
            y = f.apply(x)                                # out of place, host
            assert np.array_equal(y.view(np.uint32), ref.view(np.uint32)), ("oop", mode)
            z = x.copy()
            f.apply(z, out=z)                             # in place, host: must equal the out-of-place result
            assert np.array_equal(z.view(np.uint32), ref.view(np.uint32)), ("in place", mode)
            w = np.concatenate([x, np.zeros(64, np.float32)])
            f.apply(w[:L], out=w[8:L + 8])                # partial overlap
            assert np.array_equal(w[8:L + 8].view(np.uint32), ref.view(np.uint32)), ("overlap", mode)
            v = f.apply_valid(x)                          # VALID through the pipeline: interior of the polynomial result
            rv = O.Filter1D(n, m, d, 1.0, mode).apply_valid(x)
            assert v.shape == rv.shape and np.array_equal(v.view(np.uint32), rv.view(np.uint32)), ("valid", mode)
            sg.set_exact(False)
            f.close()
        # batch of rows, in place and with a pitch
        xb = rng.standard_normal((3000, 1100)).astype(np.float32)
        f = sg.SavgolFilter(10, 2, 1, 1.0, "reflect")
        ref = O.Filter1D(10, 2, 1, 1.0, "reflect").apply(np.ascontiguousarray(xb[:, :1024]))
        zb = xb.copy()
        f.apply(zb[:, :1024], out=zb[:, :1024])
        assert np.max(np.abs(zb[:, :1024] - ref)) <= 1e-6 * np.abs(xb).max()
        assert np.array_equal(zb[:, 1024:], xb[:, 1024:])
        print("ok")
    """)
    assert "ok" in out


def test_mcstream_host_chunks_in_channel_blocks_equal_device_chunks():
    out = run_child("""
        import numpy as np, torch, savgol_b200 as sg
        n, Cn, K = 10, 700, 1024                       # 700 channels x (1024 + 21 -> 1048 floats) = 2.9 MB -> 3 channel blocks
        rng = np.random.default_rng(4)
        sig = rng.standard_normal((Cn, 5 * K + 7)).astype(np.float32)
        cuts = [0, 5, 2 * n + 1, K + 30, 2 * K + 30, 3 * K + 30, 5 * K + 7]     # incl. a chunk shorter than the window and the first fill
        res = {}
        for where in ("host", "device"):
            s = sg.SavgolMCStream(Cn, n, 2, 1, 1.0)
            got = []
            for a, b in zip(cuts[:-1], cuts[1:]):
                ch = np.ascontiguousarray(sig[:, a:b])
                if where == "device":
                    o, k = s.push(torch.from_numpy(ch).cuda())
                    got.append(o[:, :k].cpu().numpy())
                else:
                    o, k = s.push(ch)
                    got.append(o[:, :k].copy())
            o, k = s.flush(np.empty(1, np.float32) if where == "host" else torch.empty(1, device="cuda"))
            got.append(o[:, :k].copy() if where == "host" else o[:, :k].cpu().numpy())
            res[where] = np.concatenate(got, axis=1)
            assert s.samples_received == 5 * K + 7 and s.samples_output == 5 * K + 7
        assert res["host"].shape == sig.shape
        assert np.array_equal(res["host"].view(np.uint32), res["device"].view(np.uint32))
        print("ok")
    """)
    assert "ok" in out


def test_mcstream_restore_reads_exactly_one_checkpoint_and_push_checks_the_pitch(capfd):
    lib = sg.lib()
    s = sg.SavgolMCStream(5, 4, 2, 0, 1.0)
    x = torch.randn(5, 64, device="cuda")
    s.push(x)
    blob = s.save()
    big = blob + b"\xff" * 4096                       # the caller's buffer is larger than the checkpoint
    buf = C.create_string_buffer(big, len(big))
    assert lib.savgol_mcstream_restore(s._h, buf, len(big)) == 0
    o1, k1 = s.push(x)
    s2 = sg.SavgolMCStream(5, 4, 2, 0, 1.0)
    s2.restore(blob)
    o2, k2 = s2.push(x)
    assert k1 == k2 and torch.equal(o1[:, :k1], o2[:, :k2])
    assert lib.savgol_mcstream_restore(s._h, buf, len(blob) - 1) == -1          # truncated
    # single channel: the output row must still hold chunk_len + half_window samples (the first fill emits that many)
    s1 = sg.SavgolMCStream(1, 4, 2, 0, 1.0)
    xin = torch.randn(1, 32, device="cuda")
    small = torch.empty(1, 32, device="cuda")
    assert lib.savgol_mcstream_push(s1._h, xin.data_ptr(), 32, 32, small.data_ptr(), 32) == -1
    assert "out_pitch" in capfd.readouterr().err


def test_batch_multi_and_slices_on_a_device_list(oracle):
    lib = sg.lib()
    ndev = torch.cuda.device_count()
    devs = [i % ndev for i in range(max(3, ndev))]
    darr = (C.c_int * len(devs))(*devs)
    rng = np.random.default_rng(8)
    # (a) host batch sharded by rows
    f = sg.SavgolFilter(16, 3, 1, 1.0, "reflect")
    x = rng.standard_normal((301, 4096)).astype(np.float32)
    y = np.empty_like(x)
    assert lib.savgol_apply_batch_multi(f.handle, x.ctypes.data, y.ctypes.data, 301, 4096, 4096, 4096, darr, len(devs)) == 0
    assert np.array_equal(bits(y), bits(f.apply(x)))
    # (b) one long host signal partitioned along its length, all modes, exact flavour bit-identical to the oracle
    L = (1 << 21) + 333
    xl = rng.standard_normal(L).astype(np.float32)
    for mode in ("periodic", "polynomial", "reflect", "constant"):
        g = sg.SavgolFilter(32, 4, 2, 1.0, mode)
        yl = np.empty_like(xl)
        sg.set_exact(True)
        assert lib.savgol_apply_batch_multi(g.handle, xl.ctypes.data, yl.ctypes.data, 1, L, L, L, darr, len(devs)) == 0
        sg.set_exact(False)
        assert np.array_equal(bits(yl), bits(oracle.Filter1D(32, 4, 2, 1.0, mode).apply(xl))), mode
        g.close()
    # (c) device-resident slices, halos read from the neighbour slices
    for mode in ("periodic", "polynomial", "reflect"):
        g = sg.SavgolFilter(12, 4, 1, 1.0, mode)
        lens = [70000, 4096, 123457][:len(devs)] + [5000] * (len(devs) - 3)
        sig = rng.standard_normal(sum(lens)).astype(np.float32)
        ins, outs, off = [], [], 0
        for d, ln in zip(devs, lens):
            ins.append(torch.from_numpy(sig[off:off + ln]).to(f"cuda:{d}"))
            outs.append(torch.empty(ln, device=f"cuda:{d}"))
            off += ln
        for d in set(devs):
            torch.cuda.synchronize(d)
        pi = (C.c_void_p * len(devs))(*[t.data_ptr() for t in ins])
        po = (C.c_void_p * len(devs))(*[t.data_ptr() for t in outs])
        pl = (C.c_size_t * len(devs))(*lens)
        sg.set_exact(True)
        assert lib.savgol_apply_slices(g.handle, pi, po, pl, darr, len(devs)) == 0
        sg.set_exact(False)
        got = np.concatenate([t.cpu().numpy() for t in outs])
        assert np.array_equal(bits(got), bits(oracle.Filter1D(12, 4, 1, 1.0, mode).apply(sig))), mode
        g.close()
    assert lib.savgol_apply_batch_multi(f.handle, x.ctypes.data, y.ctypes.data, 301, 4096, 4096, 4096, (C.c_int * 1)(99), 1) == -1
    f.close()


def test_tensor_on_a_device_that_is_not_current():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    f = sg.SavgolFilter(8, 3, 0, 1.0, "reflect")
    x = torch.randn(64, 2048, device="cuda:1")
    with torch.cuda.device(0):
        y = f.apply(x)
    torch.cuda.synchronize(1)
    ref = f.apply(x.cpu().numpy())
    assert np.array_equal(bits(y.cpu().numpy()), bits(ref))


def test_inplace_compat_mode_reproduces_the_reference_in_place_result(oracle):
    # the reference's savgol_apply(f, x, x, L) reads samples it has already overwritten; with
    # savgol_b200_set_inplace_compat(1) an exactly aliased call reproduces that result bit for bit (default: the
    # alias-safe out-of-place result, which differs by ~0.1 on random data)
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    R = oracle.ref()
    lib = sg.lib()
    rng = np.random.default_rng(17)
    try:
        for n, m, d, mode in [(5, 3, 0, "polynomial"), (16, 3, 1, "reflect"), (7, 2, 2, "periodic"), (12, 4, 0, "constant"), (2, 2, 1, "polynomial")]:
            cfg = oracle.make_config(n, m, d, 0.5, mode)
            rf = R.savgol_create(C.byref(cfg))
            f = sg.SavgolFilter(n, m, d, 0.5, mode)
            for L in (2 * n + 1, 2 * n + 2, 300, 5000):
                x = rng.standard_normal((3, L)).astype(np.float32)
                want = x.copy()
                for r in range(3):
                    assert R.savgol_apply(rf, want[r].ctypes.data, want[r].ctypes.data, L) == 0
                lib.savgol_b200_set_inplace_compat(1)
                h = x.copy()
                f.apply(h, out=h)                                      # host, in place
                dv = torch.from_numpy(x.copy()).cuda()
                f.apply(dv, out=dv)                                    # device, in place
                one = torch.from_numpy(x[0].copy()).cuda()
                f.apply(one, out=one)                                  # savgol_apply (single signal)
                lib.savgol_b200_set_inplace_compat(0)
                assert np.array_equal(bits(h), bits(want)), (n, mode, L, "host")
                assert np.array_equal(bits(dv.cpu().numpy()), bits(want)), (n, mode, L, "device")
                assert np.array_equal(bits(one.cpu().numpy()), bits(want[0])), (n, mode, L, "single")
                # default semantics: the out-of-place result
                dflt = torch.from_numpy(x.copy()).cuda()
                f.apply(dflt, out=dflt)
                assert torch.equal(dflt, f.apply(torch.from_numpy(x).cuda()))
            R.savgol_destroy(rf)
            f.close()
    finally:
        lib.savgol_b200_set_inplace_compat(0)


def test_pageable_host_buffers_take_the_bounce_path_and_match_pinned_bit_for_bit():
    """malloc / numpy memory cannot be DMA'd: it crosses pinned bounce buffers filled by the host copy pool
    (host_stage.cu).  Results must not depend on the kind of host memory, on mixing kinds, or on how many host
    threads are staging at once."""
    import threading

    lib = sg.lib()
    rng = np.random.default_rng(11)
    rows, L = 8192 + 3, 4096                                  # 128 MiB: nine 16 MiB bounce chunks, ragged last one
    x = rng.standard_normal((rows, L), dtype=np.float32)
    f = sg.SavgolFilter(16, 3, 1, 1.0, "reflect")
    xp = torch.empty(rows, L, dtype=torch.float32, pin_memory=True)
    xp.copy_(torch.from_numpy(x))
    yp = torch.empty(rows, L, dtype=torch.float32, pin_memory=True)
    assert lib.savgol_apply_batch(f.handle, xp.data_ptr(), yp.data_ptr(), rows, L, L, L) == 0
    want = yp.numpy().copy()
    yd = f.apply(torch.from_numpy(x).cuda()).cpu().numpy()    # device path
    assert np.array_equal(bits(yd), bits(want))

    y = np.full_like(x, np.nan)
    assert lib.savgol_apply_batch(f.handle, x.ctypes.data, y.ctypes.data, rows, L, L, L) == 0          # pageable -> pageable
    assert np.array_equal(bits(y), bits(want))
    y[:] = np.nan
    assert lib.savgol_apply_batch(f.handle, xp.data_ptr(), y.ctypes.data, rows, L, L, L) == 0          # pinned -> pageable
    assert np.array_equal(bits(y), bits(want))
    yp.zero_()
    assert lib.savgol_apply_batch(f.handle, x.ctypes.data, yp.data_ptr(), rows, L, L, L) == 0          # pageable -> pinned
    assert np.array_equal(bits(yp.numpy()), bits(want))
    z = x.copy()
    assert lib.savgol_apply_batch(f.handle, z.ctypes.data, z.ctypes.data, rows, L, L, L) == 0          # in place
    assert np.array_equal(bits(z), bits(want))
    # pitched rows: only the first 4000 samples of each row are filtered, the tail must stay untouched
    z = x.copy()
    f2 = sg.SavgolFilter(9, 2, 0, 1.0, "polynomial")
    assert lib.savgol_apply_batch(f2.handle, z.ctypes.data, z.ctypes.data, rows, 4000, L, L) == 0
    ref = f2.apply(torch.from_numpy(np.ascontiguousarray(x[:, :4000])).cuda()).cpu().numpy()
    assert np.array_equal(bits(z[:, :4000]), bits(ref)) and np.array_equal(z[:, 4000:], x[:, 4000:])

    # one long pageable signal (pieces with halos), in place too
    xs = rng.standard_normal(40 * (1 << 20) + 777, dtype=np.float32)    # 160 MiB: eleven pieces
    fs = sg.SavgolFilter(12, 4, 2, 0.5, "periodic")
    ws = fs.apply(torch.from_numpy(xs).cuda()).cpu().numpy()
    ys = fs.apply(xs)
    assert np.array_equal(bits(ys), bits(ws))
    zs = xs.copy()
    fs.apply(zs, out=zs)
    assert np.array_equal(bits(zs), bits(ws))

    # four host threads staging pageable batches at once share the copy pool
    parts = np.array_split(np.arange(rows), 4)
    outs = [np.empty((len(p), L), np.float32) for p in parts]
    ins = [np.ascontiguousarray(x[p]) for p in parts]
    rc = [None] * 4

    def work(i):
        rc[i] = lib.savgol_apply_batch(f.handle, ins[i].ctypes.data, outs[i].ctypes.data, len(parts[i]), L, L, L)

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert rc == [0, 0, 0, 0]
    assert np.array_equal(bits(np.concatenate(outs)), bits(want))

    # 2D images and the multichannel stream: pageable == pinned
    g = sg.Savgol2DFilter(3, 3, 3)
    imgs = rng.standard_normal((5, 1024, 1536), dtype=np.float32)
    ip = torch.empty(imgs.shape, dtype=torch.float32, pin_memory=True)
    ip.copy_(torch.from_numpy(imgs))
    op = torch.empty(imgs.shape, dtype=torch.float32, pin_memory=True)
    g.apply(ip, "reflect", out=op)
    o = g.apply(imgs, "reflect")
    assert np.array_equal(bits(o), bits(op.numpy()))
    o2 = imgs.copy()
    g.apply(o2, "reflect", out=o2)                            # in place, one image per slot
    assert np.array_equal(bits(o2), bits(o))

    ch, K = 3000, 2048                                        # 24 MiB per push: two channel blocks
    got = []
    for pinned in (False, True):
        st = sg.SavgolMCStream(ch, 8, 3)
        acc = []
        for k in range(2):
            blk = np.ascontiguousarray(x[:ch, k * K:(k + 1) * K])
            if pinned:
                t = torch.empty(ch, K, dtype=torch.float32, pin_memory=True)
                t.copy_(torch.from_numpy(blk))
                blk = t
            out, cnt = st.push(blk)
            acc.append(np.asarray(out[:, :cnt]).copy())
        out, cnt = st.flush(blk)
        acc.append(np.asarray(out[:, :cnt]).copy())
        got.append(np.concatenate(acc, axis=1))
        st.close()
    assert got[0].shape == (ch, 2 * K) and np.array_equal(bits(got[0]), bits(got[1]))


def test_pageable_views_with_odd_offsets_and_pitches_through_the_copy_pool():
    """The host copy pool moves rows of any width / pitch / alignment (streaming stores need an aligned body: head
    and tail bytes are copied separately); the padding of the output buffer must stay untouched."""
    lib = sg.lib()
    rng = np.random.default_rng(21)
    f = sg.SavgolFilter(9, 3, 0, 1.0, "reflect")
    for rows, L, pitch_in, off_in, pitch_out, off_out in ((3000, 4093, 4100, 3, 4101, 2), (700, 20001, 20001, 0, 20003, 1), (1, 3_000_001, 3_000_001, 0, 3_000_001, 0)):
        big_in = rng.standard_normal((rows, pitch_in + 4)).astype(np.float32)
        big_out = np.full((rows, pitch_out + 4), 7.0, np.float32)
        x = big_in[:, off_in:off_in + L]
        y = big_out[:, off_out:off_out + L]
        assert lib.savgol_apply_batch(f.handle, x.ctypes.data, y.ctypes.data, rows, L, big_in.shape[1], big_out.shape[1]) == 0
        want = f.apply(torch.from_numpy(np.ascontiguousarray(x)).cuda()).cpu().numpy()
        assert np.array_equal(bits(y), bits(want)), (rows, L)
        assert np.all(big_out[:, :off_out] == 7.0) and np.all(big_out[:, off_out + L:] == 7.0)
    # 2D: a pitched pageable image batch
    g = sg.Savgol2DFilter(4, 4, 3)
    big = rng.random((4, 1500, 1203)).astype(np.float32)
    outb = np.full_like(big, 7.0)
    imgs, outs = big[:, :, 1:1201], outb[:, :, 2:1202]
    assert lib.savgol2d_apply_batch(g.handle, imgs.ctypes.data, 1500, 1200, 1203, 1500 * 1203, outs.ctypes.data, 1203, 1500 * 1203, 4, 2) == 0
    want = g.apply(torch.from_numpy(np.ascontiguousarray(imgs)).cuda(), "reflect").cpu().numpy()
    assert np.array_equal(bits(outs), bits(want))
    assert np.all(outb[:, :, :2] == 7.0) and np.all(outb[:, :, 1202:] == 7.0)


def test_large_host_images_over_a_device_list_take_the_band_path_concurrently():
    """savgol2d_apply_batch_multi with one large image per device thread: every thread cuts its image into row bands
    through its own pipeline; same numbers as the device path."""
    lib = sg.lib()
    rng = np.random.default_rng(31)
    imgs = rng.random((3, 2048, 2304), dtype=np.float32)
    out = np.empty_like(imgs)
    f = sg.Savgol2DFilter(7, 7, 3)
    devs = (C.c_int * 3)(0, 0, 0)
    assert lib.savgol2d_apply_batch_multi(f.handle, imgs.ctypes.data, 2048, 2304, 2304, 2048 * 2304, out.ctypes.data, 2304, 2048 * 2304, 3, 2, devs, 3) == 0
    want = f.apply(torch.from_numpy(imgs).cuda(), "reflect").cpu().numpy()
    assert np.array_equal(bits(out), bits(want))


@pytest.mark.timeout(300)
def test_host_copy_pool_under_contention():
    """Eight host threads hammer the pageable path (shared copy pool, leased pipelines) with calls of different
    sizes; every result must equal the single-threaded one."""
    import threading

    lib = sg.lib()
    rng = np.random.default_rng(41)
    f = sg.SavgolFilter(6, 2, 0, 1.0, "constant")
    shapes = [(300, 1100), (64, 4096), (2000, 777), (17, 70001), (1200, 1024), (5, 400000), (900, 2049), (128, 8192)]
    ins = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    want = [f.apply(torch.from_numpy(a).cuda()).cpu().numpy() for a in ins]
    bad = []

    def work(i):
        a = ins[i]
        out = np.empty_like(a)
        for _ in range(12):
            out.fill(np.nan)
            rc = lib.savgol_apply_batch(f.handle, a.ctypes.data, out.ctypes.data, a.shape[0], a.shape[1], a.shape[1], a.shape[1])
            if rc != 0 or not np.array_equal(bits(out), bits(want[i])):
                bad.append(i)
                return

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(shapes))]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not bad, bad
