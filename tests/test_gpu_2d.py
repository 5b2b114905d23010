"""GPU parity tests of the 2D path (savgol2d_* through the C ABI) vs the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
import savgol_b200 as sg  # noqa: E402
from tolerance import l1_gain_2d, parity_tol  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(autouse=True)
def _fast_mode():
    sg.set_exact(False)
    yield
    sg.set_exact(False)


CASES = [(1, 1, 1, 0, 0), (2, 2, 2, 0, 0), (2, 2, 2, 1, 0), (2, 1, 2, 0, 1), (3, 4, 4, 1, 1), (7, 7, 3, 0, 0), (7, 7, 3, 2, 0),
         (5, 3, 6, 2, 2), (16, 16, 6, 0, 0), (4, 4, 3, 0, 0), (8, 8, 2, 0, 2), (12, 12, 5, 1, 0),
         # additive surfaces (order 2/3 smoothing): 4 columns per lane up to 17x17, 2 columns per lane above
         (1, 1, 2, 0, 0), (3, 3, 3, 0, 0), (8, 8, 3, 0, 0), (9, 9, 3, 0, 0), (12, 12, 2, 0, 0), (16, 16, 3, 0, 0)]


@pytest.mark.parametrize("nx,ny,order,dx,dy", CASES)
def test_apply_all_boundaries(oracle, nx, ny, order, dx, dy):
    rng = np.random.default_rng(nx * 1000 + ny * 100 + order * 10 + dx + dy)
    hx, hy = 0.5, 2.0
    o = oracle.Filter2D(nx, ny, order, dx, dy, hx, hy)
    f = sg.Savgol2DFilter(nx, ny, order, dx, dy, hx, hy)
    for rows, cols in ((2 * ny + 1, 2 * nx + 1), (2 * ny + 2, 2 * nx + 5), (97, 131), (260, 300)):
        img = rng.standard_normal((rows, cols)).astype(np.float32)
        d_img = torch.from_numpy(img).cuda()
        tol = parity_tol(img, o.scale, l1_gain_2d(o))
        for b in ("valid", "constant", "reflect"):
            ref = np.full(img.shape, -3.0, np.float32)
            o.apply(img, b, ref)
            out = torch.full(img.shape, -3.0, device="cuda")
            f.apply(d_img, b, out=out)
            got = out.cpu().numpy()
            assert np.max(np.abs(got - ref)) <= tol, (rows, cols, b, float(np.max(np.abs(got - ref))), tol)
            if b == "valid":   # border untouched (ref: include/iterative/savgol2d.h:108-112)
                assert np.all(got[:ny] == -3.0) and np.all(got[:, :nx] == -3.0)
            sg.set_exact(True)
            out.fill_(-3.0)
            f.apply(d_img, b, out=out)
            sg.set_exact(False)
            assert np.array_equal(bits(out.cpu().numpy()), bits(ref)), (rows, cols, b)
        v = f.apply_valid(d_img).cpu().numpy()
        assert np.max(np.abs(v - o.apply_valid(img))) <= tol


def test_golden_reference_outputs_exact(golden_dir):
    G = np.load(os.path.join(golden_dir, "ref_outputs.npz"))
    sg.set_exact(True)
    for ci, (nx, ny, o, dx, dy) in enumerate(G["cases2d"]):
        f = sg.Savgol2DFilter(int(nx), int(ny), int(o), int(dx), int(dy), 0.5, 2.0)
        img = torch.from_numpy(G[f"d{ci}_img"]).cuda()
        for b in range(3):
            out = torch.full(img.shape, -3.0, device="cuda")
            f.apply(img, b, out=out)
            assert np.array_equal(bits(out.cpu().numpy()), bits(G[f"d{ci}_apply_b{b}"])), (ci, b)


def test_config4_window_batch_constant(oracle):
    # BASELINE config 4 filter (15x15, order 3, constant) on a small batch of images, strided views
    rng = np.random.default_rng(3)
    f = sg.Savgol2DFilter(7, 7, 3)
    o = oracle.Filter2D(7, 7, 3)
    big = torch.from_numpy(rng.random((3, 200, 340)).astype(np.float32)).cuda()
    view = big[:, 10:190, 20:320]          # in_stride 340, image pitch 200*340
    out = torch.zeros(3, 180, 300, device="cuda")
    f.apply(view, "constant", out=out)
    for i in range(3):
        ref = o.apply(view[i].cpu().numpy().copy(), "constant")
        assert np.max(np.abs(out[i].cpu().numpy() - ref)) <= 1e-6
    # a view whose rows are not 16-byte aligned takes the same kernel through its per-chunk staging path
    view2 = big[:, 10:190, 21:321]
    out2 = torch.zeros(3, 180, 300, device="cuda")
    f.apply(view2, "constant", out=out2)
    for i in range(3):
        ref = o.apply(view2[i].cpu().numpy().copy(), "constant")
        assert np.max(np.abs(out2[i].cpu().numpy() - ref)) <= 1e-6
    # host path == device path
    h = f.apply(view[0].cpu().numpy().copy(), "constant")
    assert np.array_equal(bits(h), bits(out[0].cpu().numpy()))


def test_reference_properties():
    # ref: test/iterative/test_savgol2d.c:126-356 -- constant / plane preserved, derivatives of polynomials
    yy, xx = np.mgrid[0:40, 0:50].astype(np.float32)
    const = np.full((20, 20), 42.0, np.float32)
    f = sg.Savgol2DFilter(2, 2, 2)
    assert np.allclose(f.apply(torch.from_numpy(const).cuda(), "constant").cpu().numpy(), 42.0, atol=1e-3)
    plane = (2 * xx + 3 * yy).astype(np.float32)
    v = f.apply_valid(torch.from_numpy(plane).cuda()).cpu().numpy()
    assert np.allclose(v, plane[2:-2, 2:-2], atol=1e-2)
    for (dx, dy, img, want) in [(1, 0, 5 * xx, 5.0), (0, 1, 7 * yy, 7.0), (2, 0, xx * xx, 2.0), (0, 2, 3 * yy * yy, 6.0),
                                (1, 1, 4 * xx * yy, 4.0)]:
        g = sg.Savgol2DFilter(3, 3, 3, dx, dy).apply_valid(torch.from_numpy(img.astype(np.float32)).cuda()).cpu().numpy()
        assert np.allclose(g, want, atol=2e-2), (dx, dy)
    # rectangular window (ref :508-543)
    r = sg.Savgol2DFilter(2, 1, 2).apply(torch.from_numpy(const).cuda(), "constant").cpu().numpy()
    assert np.allclose(r, 42.0, atol=1e-3)


def test_wrappers_vs_oracle_composition(oracle):
    # ref: src/savgol2d.c:462-618 -- gradient / hessian / laplacian are compositions of single filters
    rng = np.random.default_rng(8)
    img = rng.standard_normal((64, 80)).astype(np.float32)
    d = torch.from_numpy(img).cuda()
    gx, gy = sg.gradient(d, 3, 3, 3, 0.5, 2.0, "constant")
    hxx, hxy, hyy = sg.hessian(d, 3, 3, 3, 0.5, 2.0, "reflect")
    lap = sg.laplacian(d, 3, 3, 3, 0.5, 2.0, "constant")

    def ref(dx, dy, b):
        return oracle.Filter2D(3, 3, 3, dx, dy, 0.5, 2.0).apply(img, b)
    def bound(*derivs):
        # tests/tolerance.py: 1e-6 * max|x| * (output scale) * (L1 gain); a sum of filters adds their bounds
        return sum(parity_tol(img, oracle.Filter2D(3, 3, 3, dx, dy, 0.5, 2.0).scale, l1_gain_2d(oracle.Filter2D(3, 3, 3, dx, dy, 0.5, 2.0)))
                   for dx, dy in derivs)
    assert np.max(np.abs(gx.cpu().numpy() - ref(1, 0, "constant"))) <= bound((1, 0))
    assert np.max(np.abs(gy.cpu().numpy() - ref(0, 1, "constant"))) <= bound((0, 1))
    assert np.max(np.abs(hxx.cpu().numpy() - ref(2, 0, "reflect"))) <= bound((2, 0))
    assert np.max(np.abs(hxy.cpu().numpy() - ref(1, 1, "reflect"))) <= bound((1, 1))
    assert np.max(np.abs(hyy.cpu().numpy() - ref(0, 2, "reflect"))) <= bound((0, 2))
    want = ref(2, 0, "constant") + ref(0, 2, "constant")
    assert np.max(np.abs(lap.cpu().numpy() - want)) <= bound((2, 0), (0, 2))
    # host-pointer wrappers give the same numbers (same kernel, same table)
    lap_h = sg.laplacian(img.copy(), 3, 3, 3, 0.5, 2.0, "constant")
    assert np.array_equal(lap_h, lap.cpu().numpy())
    with pytest.raises(RuntimeError):
        sg.laplacian(d, 3, 3, 1)


def test_errors_like_reference():
    f = sg.Savgol2DFilter(3, 3, 2)
    lib = sg.lib()
    x = torch.zeros(5, 5, device="cuda")
    assert lib.savgol2d_apply_valid(f.handle, x.data_ptr(), 5, 5, 5, x.data_ptr(), 5) == -1   # ref: src/savgol2d.c:371
    assert lib.savgol2d_apply(f.handle, None, 5, 5, 5, x.data_ptr(), 5, 1) == -1
    assert lib.savgol2d_apply(f.handle, x.data_ptr(), 5, 5, 5, x.data_ptr(), 5, 0) == -1


def test_laplacian_is_one_fused_pass(oracle):
    # W_lap = Wxx/dx^2 + Wyy/dy^2 is one polynomial weight table -> one separable launch
    rng = np.random.default_rng(12)
    img = rng.standard_normal((300, 400)).astype(np.float32)
    d = torch.from_numpy(img).cuda()
    c0 = sg.launch_count()
    lap = sg.laplacian(d, 7, 7, 3, 1.0, 1.0, "constant")
    torch.cuda.synchronize()
    assert sg.launch_count() - c0 == 1
    want = oracle.Filter2D(7, 7, 3, 2, 0).apply(img, "constant") + oracle.Filter2D(7, 7, 3, 0, 2).apply(img, "constant")
    assert np.max(np.abs(lap.cpu().numpy() - want)) <= 1e-6 * float(np.abs(img).max())
    # the exact flavour keeps the reference's composition (two filters + add) bit for bit
    sg.set_exact(True)
    lap_e = sg.laplacian(d, 7, 7, 3, 1.0, 1.0, "constant")
    sg.set_exact(False)
    assert np.array_equal(bits(lap_e.cpu().numpy()), bits(want))


@pytest.mark.parametrize("n,ny,order,dx,dy", [(7, 7, 3, 0, 0), (3, 3, 2, 1, 0), (8, 8, 4, 0, 2), (5, 5, 5, 1, 1), (7, 2, 3, 0, 1), (2, 9, 4, 1, 0)])
def test_streaming_kernel_bands_strips_and_edges(oracle, n, ny, order, dx, dy):
    # shapes that exercise the separable streaming kernel's work decomposition: many bands with an odd
    # last band (the padded extra row must not be read past the image), one / two / many strips, narrow
    # images whose single strip has both x edges, a batch with image pitch
    rng = np.random.default_rng(500 + n)
    o = oracle.Filter2D(n, ny, order, dx, dy)     # (n, ny) = (half_window_x, half_window_y): rectangular windows too
    f = sg.Savgol2DFilter(n, ny, order, dx, dy)
    for images, rows, cols in ((1, 1101, 1024), (1, 333, 64), (2, 97, 256), (3, 150, 132), (1, 2 * ny + 2, 516)):
        x = rng.standard_normal((images, rows, cols)).astype(np.float32)
        d = torch.from_numpy(x).cuda()
        tol = parity_tol(x, o.scale, l1_gain_2d(o))
        for b in ("valid", "constant", "reflect"):
            out = torch.full(x.shape, -3.0, device="cuda")
            f.apply(d, b, out=out)
            got = out.cpu().numpy()
            for i in range(images):
                ref = np.full((rows, cols), -3.0, np.float32)
                o.apply(x[i], b, ref)
                err = float(np.max(np.abs(got[i] - ref)))
                assert err <= tol, (images, rows, cols, b, i, err, tol)


@pytest.mark.parametrize("boundary", ["constant", "reflect"])
def test_row_bands_reassemble_whole_image(oracle, boundary):
    # savgol2d_apply_band: bands of one image (the per-GPU pieces of a row-sharded image) == whole-image rows, bit for bit
    rng = np.random.default_rng(77)
    img = torch.from_numpy(rng.standard_normal((700, 520)).astype(np.float32)).cuda()
    for nx, ny, order in ((7, 7, 3), (3, 5, 4)):
        f = sg.Savgol2DFilter(nx, ny, order)
        for exact in (False, True):
            sg.set_exact(exact)
            whole = f.apply(img, boundary)
            out = torch.empty_like(img)
            cuts = [0, ny, 233, 240, 611, 700]
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                top = ny if lo > 0 else 0
                bottom = ny if hi < 700 else 0
                f.apply_band(img[lo - top: hi + bottom], top, bottom, boundary, out=out[lo:hi], image_row0=lo - top)
            sg.set_exact(False)
            assert torch.equal(out, whole), (nx, ny, order, exact, boundary)
        f.close()
    lib = sg.lib()
    f = sg.Savgol2DFilter(2, 2, 2)
    assert lib.savgol2d_apply_band(f.handle, img.data_ptr(), 50, 520, 520, img.data_ptr() + 4 * 520 * 100, 520, 1, 1, 0) == -1   # halo != ny
    assert lib.savgol2d_apply_band(f.handle, img.data_ptr(), 50, 520, 520, img.data_ptr() + 4 * 520 * 100, 520, 0, 2, 2) == -1   # VALID


def test_wrapper_components_run_concurrently_and_equal_the_single_filters():
    # savgol2d_gradient / _hessian launch their components concurrently (caller's stream + side streams, forked and
    # joined); every component must equal the single filter's output bit for bit, on the caller's stream order
    g = torch.Generator(device="cuda").manual_seed(21)
    img = torch.rand(1500, 1100, device="cuda", generator=g)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        img2 = img * 2.0                                   # produced on the caller's stream right before the call
        gx, gy = sg.gradient(img2, 7, 7, 3, 0.5, 2.0, "reflect")
        hxx, hxy, hyy = sg.hessian(img2, 5, 5, 4, 1.0, 1.0, "constant")
        after = gx + gy                                    # consumed on the caller's stream right after
    s.synchronize()
    for got, (nx, o, dx, dy, ddx, ddy, b) in ((gx, (7, 3, 1, 0, 0.5, 2.0, "reflect")), (gy, (7, 3, 0, 1, 0.5, 2.0, "reflect")),
                                              (hxx, (5, 4, 2, 0, 1.0, 1.0, "constant")), (hxy, (5, 4, 1, 1, 1.0, 1.0, "constant")),
                                              (hyy, (5, 4, 0, 2, 1.0, 1.0, "constant"))):
        f = sg.Savgol2DFilter(nx, nx, o, dx, dy, ddx, ddy)
        assert torch.equal(got, f.apply(img2, b)), (nx, o, dx, dy)
        f.close()
    assert torch.equal(after, gx + gy)
    # aliased outputs fall back to the sequential composition and still work
    buf = torch.rand(300, 400, device="cuda")
    lib = sg.lib()
    assert lib.savgol2d_gradient(3, 3, 2, buf.data_ptr(), 300, 400, 400, buf.data_ptr(), buf.data_ptr(), 1.0, 1.0, 1) == 0


def test_gradient_and_hessian_run_as_one_multi_output_launch():
    """Half-windows <= 8: the components share ONE launch of the multi-output kernel (sg2d_multi.cu) -- the image is staged
    once, each component has its own accumulator ring.  Same factors, same operation order per component as the
    single filters: bit-identical results.  ref: src/savgol2d.c:462-558 (one savgol2d_apply per component)."""
    import os
    if os.environ.get("SAVGOL_B200_WRAP_FUSED") == "0" or os.environ.get("SAVGOL_B200_WRAP_SEQ") == "1":
        pytest.skip("multi-output launches switched off by the environment")
    lib = sg.lib()
    g = torch.Generator(device="cuda").manual_seed(33)
    for rows, cols in ((257, 1031), (1024, 1024), (64, 130)):
        img = torch.rand(rows, cols, device="cuda", generator=g) - 0.5
        for hw, order, b in ((1, 2, "constant"), (2, 2, "reflect"), (2, 3, "constant"), (3, 3, "reflect"), (3, 5, "constant"),
                             (4, 4, "reflect"), (4, 5, "constant"), (2, 4, "reflect"), (5, 3, "reflect"), (7, 3, "constant"), (8, 5, "reflect")):
            c0 = sg.launch_count()
            gx, gy = sg.gradient(img, hw, hw, order, 0.5, 2.0, b)
            c1 = sg.launch_count()
            assert c1 - c0 == 1, (hw, order, c1 - c0)
            comps = [(gx, 1, 0), (gy, 0, 1)]
            if order >= 2:
                hxx, hxy, hyy = sg.hessian(img, hw, hw, order, 0.5, 2.0, b)
                assert sg.launch_count() - c1 == 1, (hw, order)
                comps += [(hxx, 2, 0), (hxy, 1, 1), (hyy, 0, 2)]
            for got, dx, dy in comps:
                f = sg.Savgol2DFilter(hw, hw, order, dx, dy, 0.5, 2.0)
                assert torch.equal(got, f.apply(img, b)), (rows, cols, hw, order, dx, dy, b)
                f.close()
    # rectangular windows and partial requests (only some components asked for)
    img = torch.rand(300, 500, device="cuda", generator=g)
    gx = torch.empty_like(img)
    c0 = sg.launch_count()
    assert lib.savgol2d_gradient(3, 2, 3, img.data_ptr(), 300, 500, 500, gx.data_ptr(), None, 1.0, 1.0, 2) == 0
    assert sg.launch_count() - c0 == 1
    f = sg.Savgol2DFilter(3, 2, 3, 1, 0)
    assert torch.equal(gx, f.apply(img, "reflect"))
    hxx, hyy = torch.empty_like(img), torch.empty_like(img)
    c0 = sg.launch_count()
    assert lib.savgol2d_hessian(3, 2, 3, img.data_ptr(), 300, 500, 500, hxx.data_ptr(), None, hyy.data_ptr(), 1.0, 1.0, 1) == 0
    assert sg.launch_count() - c0 == 1
    assert torch.equal(hxx, sg.Savgol2DFilter(3, 2, 3, 2, 0).apply(img, "constant"))
    assert torch.equal(hyy, sg.Savgol2DFilter(3, 2, 3, 0, 2).apply(img, "constant"))
    # outputs with different 16-byte phases keep the per-component launches, with the same numbers; aliased ones run in sequence
    flat = torch.zeros(2 * 300 * 500 + 8, device="cuda")
    o0, o1 = flat[:150000].view(300, 500), flat[150001:300001].view(300, 500)
    c0 = sg.launch_count()
    assert lib.savgol2d_gradient(2, 2, 2, img.data_ptr(), 300, 500, 500, o0.data_ptr(), o1.data_ptr(), 1.0, 1.0, 1) == 0
    assert sg.launch_count() - c0 == 2
    assert torch.equal(o0, sg.Savgol2DFilter(2, 2, 2, 1, 0).apply(img, "constant"))
    assert torch.equal(o1, sg.Savgol2DFilter(2, 2, 2, 0, 1).apply(img, "constant"))
    assert lib.savgol2d_gradient(2, 2, 2, img.data_ptr(), 300, 500, 500, o0.data_ptr(), o0.data_ptr(), 1.0, 1.0, 1) == 0
    assert torch.equal(o0, sg.Savgol2DFilter(2, 2, 2, 0, 1).apply(img, "constant"))       # gy written last


def test_host_image_wrappers_upload_once_and_match_the_device_path():
    """Host images (the reference's calling convention): one upload, the device path, one download per component."""
    lib = sg.lib()
    rng = np.random.default_rng(12)
    img = rng.random((300, 517)).astype(np.float32)
    d = torch.from_numpy(img).cuda()
    for hw, order, b in ((2, 2, "constant"), (3, 3, "reflect"), (9, 3, "constant")):
        gx, gy = sg.gradient(img, hw, hw, order, 0.5, 2.0, b)
        dgx, dgy = sg.gradient(d, hw, hw, order, 0.5, 2.0, b)
        assert np.array_equal(bits(gx), bits(dgx.cpu().numpy())) and np.array_equal(bits(gy), bits(dgy.cpu().numpy())), (hw, order)
        hs = sg.hessian(img, hw, hw, order, 0.5, 2.0, b)
        ds = sg.hessian(d, hw, hw, order, 0.5, 2.0, b)
        for a_, b_ in zip(hs, ds):
            assert np.array_equal(bits(a_), bits(b_.cpu().numpy())), (hw, order)
    # pitched host buffers; VALID leaves the border of the outputs untouched (ref: src/savgol2d.c:374-393)
    big = rng.random((300, 600)).astype(np.float32)
    gx = np.full((300, 600), 7.0, np.float32)
    gy = np.full((300, 600), 7.0, np.float32)
    assert lib.savgol2d_gradient(3, 2, 3, big.ctypes.data, 300, 517, 600, gx.ctypes.data, gy.ctypes.data, 1.0, 1.0, 0) == 0
    fx, fy = sg.Savgol2DFilter(3, 2, 3, 1, 0), sg.Savgol2DFilter(3, 2, 3, 0, 1)
    view = torch.from_numpy(np.ascontiguousarray(big[:, :517])).cuda()
    wx, wy = fx.apply_valid(view).cpu().numpy(), fy.apply_valid(view).cpu().numpy()
    assert np.array_equal(bits(gx[2:298, 3:514]), bits(wx)) and np.array_equal(bits(gy[2:298, 3:514]), bits(wy))
    for g_ in (gx, gy):
        assert np.all(g_[:2] == 7.0) and np.all(g_[298:] == 7.0) and np.all(g_[:, :3] == 7.0) and np.all(g_[:, 514:] == 7.0)
    # overlapping host buffers keep the reference's component-by-component order
    z = img.copy()
    assert lib.savgol2d_gradient(2, 2, 2, z.ctypes.data, 300, 517, 517, z.ctypes.data, gy[:, :517].copy().ctypes.data, 1.0, 1.0, 1) == 0


def test_single_large_host_image_travels_in_row_bands_and_matches_the_device_result():
    """One large host image (the reference's call) is cut into row bands with ny halo rows so that upload, kernel and
    download overlap; the result is bit-identical to filtering the whole image on the device."""
    rng = np.random.default_rng(5)
    for shape in ((2300, 2051), (2, 2048, 2048)):
        img = rng.random(shape, dtype=np.float32)
        dimg = torch.from_numpy(img).cuda()
        pinned = torch.empty(shape, dtype=torch.float32, pin_memory=True)
        pinned.copy_(torch.from_numpy(img))
        for hw, order, dx, dy, b in ((7, 3, 0, 0, "reflect"), (7, 3, 0, 0, "constant"), (2, 2, 1, 0, "reflect"), (12, 4, 0, 0, "constant"), (3, 5, 0, 1, "reflect")):
            f = sg.Savgol2DFilter(hw, hw, order, dx, dy)
            want = f.apply(dimg, b).cpu().numpy()
            got = f.apply(img, b)                                   # pageable host image
            assert np.array_equal(bits(got), bits(want)), (shape, hw, order, b)
            outp = torch.empty(shape, dtype=torch.float32, pin_memory=True)
            f.apply(pinned, b, out=outp)
            assert np.array_equal(bits(outp.numpy()), bits(want)), (shape, hw, order, b, "pinned")
            f.close()
    # in place keeps the image-per-slot path
    f = sg.Savgol2DFilter(7, 7, 3)
    img = rng.random((2300, 2051), dtype=np.float32)
    want = f.apply(torch.from_numpy(img).cuda(), "reflect").cpu().numpy()
    z = img.copy()
    f.apply(z, "reflect", out=z)
    assert np.array_equal(bits(z), bits(want))


def test_wrappers_on_degenerate_image_shapes():
    """Tiny and skinny images: narrower than a strip, narrower than 4 columns (the separable kernels need 4), a single
    window.  Device and host images; every component equals its single filter."""
    g = torch.Generator(device="cuda").manual_seed(44)
    for rows, cols in ((3, 5), (5, 3), (7, 4), (3, 3), (1, 9), (9, 1), (40, 2), (2, 40), (33, 129)):
        img = torch.rand(rows, cols, device="cuda", generator=g)
        himg = img.cpu().numpy()
        for hw, order, b in ((1, 2, "constant"), (1, 2, "reflect"), (2, 3, "constant")):
            gx, gy = sg.gradient(img, hw, hw, order, 1.0, 1.0, b)
            hs = sg.hessian(img, hw, hw, order, 1.0, 1.0, b)
            hgx, hgy = sg.gradient(himg, hw, hw, order, 1.0, 1.0, b)
            for got, hgot, (dx, dy) in ((gx, hgx, (1, 0)), (gy, hgy, (0, 1)), (hs[0], None, (2, 0)), (hs[1], None, (1, 1)), (hs[2], None, (0, 2))):
                f = sg.Savgol2DFilter(hw, hw, order, dx, dy)
                want = f.apply(img, b)
                assert torch.equal(got, want), (rows, cols, hw, order, b, dx, dy)
                if hgot is not None:
                    assert np.array_equal(bits(hgot), bits(want.cpu().numpy())), (rows, cols, hw, "host")
                f.close()
