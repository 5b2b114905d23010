"""CPU-side checks of the drop-in boundary (no GPU, no compute launches):
the C-ABI library loads, exports every symbol include/savgol_b200.h declares, keeps the
reference's struct layouts, builds weight tables bit-identical to the oracle's, and runs the
host-only parts of the API (creation/validation, the scalar stream shim) like the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import savgol_b200 as sg
from savgol_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "savgol_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(savgol[A-Za-z0-9_]*)\s*\(", hdr))
    declared -= {"savgol2d_valid_size", "savgol2d_num_terms"}  # static inline in the header
    lib = sg.lib()
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(_capi.PROTOTYPES), declared ^ set(_capi.PROTOTYPES)
    assert lib.savgol_b200_version() == 100


def test_struct_layouts_match_reference_abi():
    # SURVEY.md section 0 (probed from the reference headers with gcc)
    assert C.sizeof(_capi.SavgolConfig) == 12
    assert C.sizeof(_capi.SavgolFilterStruct) == 8600
    assert _capi.SavgolFilterStruct.window_size.offset == 12
    assert _capi.SavgolFilterStruct.dt_scale.offset == 16
    assert _capi.SavgolFilterStruct.center_weights.offset == 20
    assert _capi.SavgolFilterStruct.edge_weights.offset == 280
    assert C.sizeof(_capi.SavgolStreamStruct) == 296
    assert C.sizeof(_capi.Savgol2DConfig) == 16
    assert C.sizeof(_capi.Savgol2DFilterStruct) == 48


def test_compat_headers_compile_reference_style_code(tmp_path):
    # a translation unit written against the reference's header names must compile against ours
    src = tmp_path / "t.c"
    src.write_text('#include "savgolFilter.h"\n#include "savgol_stream.h"\n#include "savgol2d.h"\n'
                   "int main(void){ SavgolConfig c = SAVGOL_DERIV1(5,2,0.5f); SavgolStream s; Savgol2DConfig d;"
                   "(void)c;(void)s;(void)d; return (int)sizeof(SavgolFilter) - 8600; }\n")
    import subprocess
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_create_validation_like_reference(capfd):
    # ref: test/iterative/test_savgol.c:37-85
    f = sg.SavgolFilter(5, 2)
    assert f.window_size == 11
    f.close()
    sg.lib().savgol_destroy(None)
    for bad in [dict(half_window=0, poly_order=2), dict(half_window=2, poly_order=10),
                dict(half_window=5, poly_order=2, derivative=3), dict(half_window=33, poly_order=2),
                dict(half_window=5, poly_order=2, time_step=0.0), dict(half_window=5, poly_order=4, derivative=5)]:
        with pytest.raises(ValueError):
            sg.SavgolFilter(**bad)
    assert "savgol:" in capfd.readouterr().err


def test_all_weight_tables_bit_identical_to_oracle(oracle):
    cnt = 0
    for n in range(1, 33):
        for m in range(0, 2 * n + 1):
            if 2 * n + m + 1 >= 76:
                continue
            for d in range(0, min(m, 4) + 1):
                if m > 12 and (n + m + d) % 5:   # thin out the high-order tail
                    continue
                f = sg.SavgolFilter(n, m, d, 0.5)
                o = oracle.Filter1D(n, m, d, 0.5)
                assert np.array_equal(bits(f.center_weights), bits(o.center[: 2 * n + 1])), (n, m, d)
                assert np.array_equal(bits(f.edge_weights), bits(o.edge)), (n, m, d)
                assert np.float32(f.dt_scale) == np.float32(0.5) ** np.float32(d)
                f.close()
                cnt += 1
    assert cnt > 1500


def test_weight_properties_like_reference():
    # ref: test/iterative/test_savgol.c:91-140
    f = sg.SavgolFilter(5, 3)
    w = f.center_weights
    assert abs(w.sum() - 1.0) < 1e-5 and abs(w[0] - w[-1]) < 1e-6 and abs(w[1] - w[-2]) < 1e-6
    g = sg.SavgolFilter(5, 3, 1).center_weights
    assert abs(g[0] + g[-1]) < 1e-6 and abs(g[5]) < 1e-6


def test_2d_weights_bit_identical_to_oracle(oracle):
    for nx, ny, o, dx, dy in [(1, 1, 1, 0, 0), (2, 2, 2, 1, 0), (7, 7, 3, 0, 0), (2, 1, 2, 0, 1), (3, 4, 4, 1, 1),
                              (16, 16, 6, 0, 0), (5, 3, 6, 2, 2), (7, 7, 3, 2, 0)]:
        f = sg.Savgol2DFilter(nx, ny, o, dx, dy, 0.5, 2.0)
        of = oracle.Filter2D(nx, ny, o, dx, dy, 0.5, 2.0)
        assert np.array_equal(bits(f.weights), bits(of.W)), (nx, ny, o, dx, dy)
        assert np.float32(f.scale) == np.float32(of.scale)
        f.close()
    # ref: test/iterative/test_savgol2d.c:27-71
    for bad in [(0, 2, 2, 0, 0), (2, 2, 2, 2, 1), (1, 1, 6, 0, 0), (17, 2, 2, 0, 0), (2, 2, 7, 0, 0)]:
        with pytest.raises(ValueError):
            sg.Savgol2DFilter(*bad)


def test_scalar_stream_shim_bit_identical_to_oracle(oracle):
    # ref: test/iterative/test_savgol_stream.c:140-189 (stream == batch), here against the stream oracle
    rng = np.random.default_rng(12345)
    for n, m, d in [(5, 3, 0), (10, 2, 1), (2, 2, 2), (32, 4, 2)]:
        x = rng.standard_normal(300).astype(np.float32)
        s = sg.SavgolStream(n, m, d, 0.5)
        assert s.latency == n and not s.ready
        ys = []
        for i, v in enumerate(x):
            out = s.push_full(float(v))
            assert (len(out) > 0) == (i >= 2 * n)
            ys.extend(out)
        ys.extend(s.flush())
        assert len(ys) == x.size == s.samples_output and s.samples_received == x.size
        ref = oracle.Filter1D(n, m, d, 0.5).stream_run(x)
        assert np.array_equal(bits(np.array(ys, np.float32)), bits(ref))
        # plain push: no leading edge, first valid output on sample 2n
        s.reset()
        got = [s.push(float(v)) for v in x[: 2 * n + 3]]
        assert [g[1] for g in got] == [False] * (2 * n) + [True] * 3
        assert np.float32(got[2 * n][0]) == ref[n]
        s.close()


def test_stream_flush_leading_matches_the_compiled_reference(oracle):
    # ref: src/savgol_stream.c:254-275 -- re-emits the n leading-edge outputs from the CURRENT window contents (plain
    # push skips them); compared call by call with the unmodified reference, including the counters and the
    # max_count / not-yet-full / NULL conventions
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    R = oracle.ref()
    rng = np.random.default_rng(7)
    for n, m, d in [(5, 3, 0), (10, 2, 1), (3, 2, 2)]:
        cfg = oracle.make_config(n, m, d, 0.5, 0)
        rs = R.savgol_stream_create(C.byref(cfg))
        s = sg.SavgolStream(n, m, d, 0.5)
        rbuf = (C.c_float * 40)()
        x = rng.standard_normal(2 * n + 9).astype(np.float32)
        for i, v in enumerate(x):
            if i < 2 * n + 1:   # window not full yet: nothing to re-emit, counters untouched
                assert R.savgol_stream_flush_leading(rs, rbuf, 40) == 0 and s.flush_leading() == []
            ok = C.c_bool(False)
            R.savgol_stream_push(rs, float(v), C.byref(ok))
            s.push(float(v))
            if i >= 2 * n:
                k = R.savgol_stream_flush_leading(rs, rbuf, 40)
                mine = s.flush_leading()
                assert k == n == len(mine)
                assert np.array_equal(bits(np.array(mine, np.float32)), bits(np.array(rbuf[:k], np.float32))), (n, m, d, i)
                assert R.savgol_stream_samples_output(rs) == s.samples_output
        # max_count clamps, non-positive max_count and NULL return 0 (not -1, unlike flush)
        lib = sg.lib()
        small = (C.c_float * 2)()
        assert lib.savgol_stream_flush_leading(s._h, small, 2) == R.savgol_stream_flush_leading(rs, rbuf, 2) == 2
        assert np.array_equal(bits(np.array(small[:2], np.float32)), bits(np.array(rbuf[:2], np.float32)))
        assert lib.savgol_stream_flush_leading(s._h, small, 0) == R.savgol_stream_flush_leading(rs, rbuf, 0) == 0
        assert lib.savgol_stream_flush_leading(None, small, 2) == 0
        R.savgol_stream_destroy(rs)
        s.close()


def test_apply_fails_loudly_without_gpu(capfd):
    if sg.device_ok():
        pytest.skip("GPU present")
    f = sg.SavgolFilter(5, 2)
    x = np.zeros(100, np.float32)
    with pytest.raises(RuntimeError):
        f.apply(x)
    assert "no CPU fallback" in capfd.readouterr().err
    assert f.apply_valid(x).size == 0


def test_reference_export_tool_on_our_library_emits_identical_header(tmp_path):
    """The reference's coefficient export CLI (src/savgol_export.c, reads f->center_weights /
    f->edge_weights / f->window_size directly) compiled unmodified against libsavgol_b200 must print
    the same header as when built against the reference library -- struct layout + weight parity."""
    import subprocess
    ours = os.path.join(ROOT, "oracle", "_ref", "savgol_export_b200")
    ref = os.path.join(ROOT, "oracle", "_ref", "savgol_export_ref")
    if not (os.path.exists(ours) and os.path.exists(ref)):
        pytest.skip("drop-in binaries not built (make -C oracle dropin)")
    for n, m, d in [(12, 4, 0), (16, 3, 1), (32, 4, 2), (5, 2, 2)]:
        a, b = tmp_path / "a.h", tmp_path / "b.h"
        for exe, out in ((ours, a), (ref, b)):
            subprocess.run([exe, "-n", str(n), "-m", str(m), "-d", str(d), "-o", str(out)], check=True,
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        strip = lambda p: [l for l in open(p).read().splitlines() if "enerated" not in l]
        assert strip(a) == strip(b), (n, m, d)


def test_2d_fast_path_covers_the_configuration_space():
    # host-only: every valid 2D filter (square and rectangular windows up to 33x33, orders 0..6, derivatives
    # up to 2+2) must get a separable plan within the acceptance bound -- a rejected plan would silently send
    # that filter to the literal window kernel (14x slower)
    import ctypes as C
    import savgol_b200 as sg
    lib = sg.lib()
    total, rejected, worst, ranks = 0, [], 0.0, {}
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    os.dup2(devnull, 2)          # invalid windows print the reference's "weight computation failed" line
    try:
        for n in range(1, 17):
            for order in range(0, 7):
                for dx in range(0, 3):
                    for dy in range(0, 3):
                        if dx + dy > order:
                            continue
                        for nx, ny in ((n, n), (n, max(1, n // 2))):
                            cfg = sg._capi.Savgol2DConfig(nx, ny, order, dx, dy, 1.0, 1.0)
                            if not lib.savgol2d_config_valid(C.byref(cfg)):
                                continue
                            h = lib.savgol2d_create(C.byref(cfg))
                            if not h:
                                continue
                            r, e = C.c_int(), C.c_float()
                            assert lib.savgol2d_b200_plan(h, C.byref(r), C.byref(e)) == 0
                            total += 1
                            ranks[r.value] = ranks.get(r.value, 0) + 1
                            if r.value == 0:
                                rejected.append((nx, ny, order, dx, dy))
                            else:
                                worst = max(worst, e.value)
                            lib.savgol2d_destroy(h)
    finally:
        os.dup2(saved, 2)
        os.close(devnull)
        os.close(saved)
    assert total > 1000
    assert len(rejected) <= total // 100, rejected[:20]
    assert worst <= 4e-7, worst
    assert set(ranks) <= {0, 1, 2, 3, 4}


def test_2d_additive_plans():
    # host-only: which filters the planner hands to the additive kernel (sg2d_add.cu): exactly the rank-2 surfaces
    # W(y,x) = u(x) + v(y) with a square window -- every order-2/3 smoothing filter (and the even/even derivatives where
    # they are rank 2) -- reproduced by u + v to within the same acceptance bound as the factorisations
    import ctypes as C
    lib = sg.lib()
    for n in range(1, 17):
        for order in (2, 3):
            if (2 * n + 1) ** 2 < (order + 1) * (order + 2) // 2:
                continue                       # fewer window points than polynomial terms (3x3, order 3)
            f = sg.Savgol2DFilter(n, n, order)
            r, e = C.c_int(), C.c_float()
            assert lib.savgol2d_b200_plan(f.handle, C.byref(r), C.byref(e)) == 0
            assert lib.savgol2d_b200_plan_kind(f.handle) == 2 and r.value == 2 and e.value <= 4e-7, (n, order, r.value, e.value)
            # the planner's claim, checked here from the public weight table: W - (row 0 + column 0 - corner) == 0
            W = f.weights.astype(np.float64)
            assert np.max(np.abs(W - (W[n:n + 1, :] + W[:, n:n + 1] - W[n, n]))) <= 2e-7 * np.abs(W).max() + 1e-9
            f.close()
    for args, kind in (((7, 7, 1), 1), ((7, 7, 0), 1), ((7, 7, 4), 1), ((7, 5, 3), 1), ((7, 7, 3, 1, 0), 1), ((7, 7, 3, 1, 1), 1)):
        f = sg.Savgol2DFilter(*args)
        assert lib.savgol2d_b200_plan_kind(f.handle) == kind, args
        f.close()
    assert lib.savgol2d_b200_plan_kind(None) == -1


def test_host_copy_pool_moves_rows_of_any_shape_and_alignment():
    """The pageable-memory path of the host staging (csrc/host_stage.cu) rests on a multi-threaded 2D copy with
    streaming stores; it needs no GPU.  Random widths / pitches / byte offsets, sizes below and above the point where
    the pool takes over, concurrent callers."""
    import threading

    import savgol_b200 as sg
    lib = sg.lib()
    rng = np.random.default_rng(8)
    cases = [(1, 1), (3, 17), (1, 300_001), (700, 4093), (64, 70_001), (5, 3_000_001), (2000, 1024), (1, 16 << 20)]
    for rows, width in cases:
        for so, do in ((0, 0), (1, 3), (5, 2)):
            sp, dp = width + int(rng.integers(0, 40)), width + int(rng.integers(0, 40))
            src = rng.integers(0, 256, rows * sp + 64, dtype=np.uint8)
            dst = np.full(rows * dp + 64, 0xAB, np.uint8)
            n = lib.savgol_b200_host_copy2d(dst.ctypes.data + do, dp, src.ctypes.data + so, sp, width, rows)
            assert n >= 1
            want = np.full_like(dst, 0xAB)
            for r in range(rows):
                want[do + r * dp:do + r * dp + width] = src[so + r * sp:so + r * sp + width]
            assert np.array_equal(dst, want), (rows, width, so, do)
    # concurrent callers share the pool
    srcs = [rng.integers(0, 256, 6_000_000 + 1000 * i, dtype=np.uint8) for i in range(6)]
    dsts = [np.zeros_like(a) for a in srcs]

    def work(i):
        for _ in range(10):
            dsts[i][:] = 0
            lib.savgol_b200_host_copy2d(dsts[i].ctypes.data, srcs[i].size, srcs[i].ctypes.data, srcs[i].size, srcs[i].size, 1)
            assert np.array_equal(dsts[i], srcs[i])

    th = [threading.Thread(target=work, args=(i,)) for i in range(6)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert all(np.array_equal(d, s_) for d, s_ in zip(dsts, srcs))
    assert lib.savgol_b200_host_copy2d(None, 0, None, 0, 0, 0) >= 1


def test_dispatch_of_1d_launches_is_what_the_design_says():
    """savgol_b200_plan_1d: the host logic that picks the kernel family and the work decomposition (DESIGN.md 4.1 /
    4.2), checked without a GPU."""
    import ctypes as C

    import savgol_b200 as sg
    lib = sg.lib()

    def plan(n, rows, L, pitch=None, off=0, stream=0, exact=0, poly=0):
        fam, g, ph, tl, spr = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
        v = lib.savgol_b200_plan_1d(n, stream, exact, rows, L, pitch or L, off, poly, C.byref(fam), C.byref(g), C.byref(ph), C.byref(tl), C.byref(spr))
        assert v >= 0
        return fam.value, g.value, ph.value, tl.value, spr.value

    GENERIC, PACKED, TMA = 0, 1, 2
    # config 2 / config 5 shapes: bulk-tensor kernels, four / one segment(s) per row
    assert plan(16, 65536, 4096) == (TMA, 0, 0, 0, 4)
    assert plan(10, 1 << 20, 1024, stream=1) == (TMA, 0, 0, 0, 1)
    # wide windows and small launches keep the cp.async kernel (the measured rule of tma_eligible)
    assert plan(32, 65536, 4096)[0] == GENERIC and plan(16, 16, 4096)[0] == GENERIC
    assert plan(16, 65536, 4096, exact=1) == (GENERIC, 0, 0, 0, 4)
    # misaligned rows: per-row phase, segments counted for the worst phase, short tails folded
    assert plan(16, 65520, 4097) == (GENERIC, 0, 1, 1, 4)             # 4097 + 31 = 4 x 1024 + 32
    assert plan(16, 65552, 4095) == (GENERIC, 0, 1, 1, 4)
    assert plan(16, 1000, 4096, off=1) == (GENERIC, 0, 1, 1, 4)       # aligned pitch, misaligned base
    assert plan(16, 53687, 5001) == (GENERIC, 0, 1, 0, 5)
    assert plan(16, 1, 4097) == (GENERIC, 0, 0, 1, 4)                 # a single row only needs an aligned base
    assert plan(16, 65536, 4100) == (GENERIC, 0, 0, 1, 4)             # aligned rows with a tail of 4: generic + tail beats 5 TMA segments
    assert plan(16, 65536, 4096, pitch=4100) == (TMA, 0, 0, 0, 4)
    assert plan(16, 65520, 4097, exact=1) == (GENERIC, 0, 0, 0, 5)    # the exact flavours keep the plain decomposition
    # short rows: lanes per row by length (+3 when the rows need a phase), wider slots when shared memory runs out
    assert plan(16, 1000, 512) == (PACKED, 16, 0, 0, 1)
    assert plan(16, 1000, 360) == (PACKED, 16, 0, 0, 1)
    assert plan(16, 1000, 256) == (PACKED, 8, 0, 0, 1)
    assert plan(16, 1000, 250) == (PACKED, 8, 1, 0, 1)
    assert plan(16, 1000, 254) == (PACKED, 16, 1, 0, 1)               # 254 + 3 > 256
    assert plan(16, 1000, 64) == (PACKED, 2, 0, 0, 1)
    assert plan(16, 1000, 33) == (PACKED, 2, 1, 0, 1)
    assert plan(2, 1000, 20) == (PACKED, 1, 0, 0, 1)
    assert plan(16, 1000, 510, off=2)[0] == GENERIC                   # 510 + 3 > 512: one generic segment on a phase
    assert plan(16, 1, 300)[0] == GENERIC                             # a single short row is not a batch
    assert plan(16, 1000, 513) == (GENERIC, 0, 1, 0, 1)
    assert plan(16, 1000, 1000) == (GENERIC, 0, 0, 0, 1)
    assert plan(16, 1000, 1001) == (GENERIC, 0, 1, 1, 1)              # 1001 + 31 = 1024 + 8: one segment + tail
    assert plan(16, 1000, 1040) == (GENERIC, 0, 0, 1, 1)
    assert lib.savgol_b200_plan_1d(0, 0, 0, 1, 100, 100, 0, 0, None, None, None, None, None) == -1
    assert lib.savgol_b200_plan_1d(16, 0, 0, 1, 100, 50, 0, 0, None, None, None, None, None) == -1


def test_staging_chunk_policy():
    """Host-pointer calls: 64 MiB chunks (16 MiB through bounce buffers), shrunk for calls that would fit one or two
    chunks so that upload, kernel and download overlap; small calls stay whole."""
    import os

    import savgol_b200 as sg
    if any(os.environ.get(k) for k in ("SAVGOL_B200_CHUNK_MIB", "SAVGOL_B200_BOUNCE_MIB", "SAVGOL_B200_FIXED_CHUNK")):
        pytest.skip("staging chunk overridden by the environment")
    chunk = sg.lib().savgol_b200_staging_chunk
    MiB = 1 << 18   # floats
    assert chunk(1 << 30, 0) == 64 * MiB and chunk(1 << 30, 1) == 16 * MiB       # large calls: the configured chunk
    assert chunk(64 * MiB, 0) == 8 * MiB and chunk(64 * MiB, 1) == 16 * MiB      # 64 MiB: eight chunks pinned, four pageable
    assert chunk(16 * MiB, 0) == 2 * MiB and chunk(16 * MiB, 1) == 4 * MiB
    assert chunk(8 * MiB, 0) == 2 * MiB                                            # never below 2 MiB
    assert chunk(1_000_000, 0) == 64 * MiB and chunk(1_000_000, 1) == 16 * MiB    # under 8 MiB: one chunk (config 1)
    assert chunk(0, 0) == 64 * MiB


def test_gradient_and_hessian_take_one_launch_for_half_windows_up_to_8():
    """savgol2d_b200_wrapper_plan: which wrapper configurations have a multi-output kernel (csrc/sg2d_multi.cu) --
    half-windows <= 8 and 1..3 (gradient) / 1..2 (Hessian) separable factors per component."""
    import savgol_b200 as sg
    plan = sg.lib().savgol2d_b200_wrapper_plan
    for hw in range(1, 9):
        for order in range(1, min(5, 2 * hw) + 1):
            assert plan(hw, hw, order, 0) == 1, (hw, order)
            if order >= 2:
                assert plan(hw, hw, order, 1) == 1, (hw, order)
    assert plan(3, 2, 3, 0) == 1 and plan(2, 7, 2, 1) == 1          # rectangular windows
    assert plan(9, 9, 3, 0) == 0 and plan(12, 12, 2, 1) == 0        # wider windows: per-component launches
    assert plan(9, 3, 2, 0) == 0
    assert plan(2, 2, 1, 1) == -1                                   # a Hessian needs order >= 2 (ref: src/savgol2d.c:507-510)
    assert plan(2, 2, 9, 0) == -1                                   # invalid configuration
