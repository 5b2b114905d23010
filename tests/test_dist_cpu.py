"""world_size-2/3 gloo tests (CPU) of the N>1 host logic: batch sharding and the halo exchange of a
partitioned long signal.  The GPU kernel is not available here, so the oracle stands in for the
per-slice compute (VALID over [left | slice | right], boundary rule at true ends) -- which is
exactly the contract of savgol_apply_halo that tests/test_gpu_1d.py checks on the device."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slice_with_oracle(O, n, m, d, mode, xs, left, right):
    """What savgol_apply_halo computes for one slice, restated with the oracle."""
    f = O.Filter1D(n, m, d, 1.0, mode)
    L = xs.size
    if left is not None and right is not None:
        return f.apply_valid(np.concatenate([left, xs, right]))
    full = f.apply(np.concatenate([left, xs]) if left is not None else (np.concatenate([xs, right]) if right is not None else xs))
    if left is not None:       # true right end: boundary rule applies there
        return full[n:]
    if right is not None:      # true left end
        return full[:L]
    return full


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import savgol_b200 as sg
    from savgol_b200 import dist as sgd
    from oracle import oracle as O
    n, m, d = 8, 3, 1
    rng = np.random.default_rng(7)
    L = 1003
    x = rng.standard_normal(L).astype(np.float32)
    a, b = sgd.shard_range(L, rank, world)
    xs = torch.from_numpy(x[a:b].copy())
    left, right = sgd.exchange_halos(xs, n, periodic=(mode == "periodic"), group=None)
    ln = left.numpy() if left is not None else None
    rn = right.numpy() if right is not None else None
    # halos are exactly the neighbouring samples
    if ln is not None:
        assert np.array_equal(ln, x[(np.arange(a - n, a)) % L])
    if rn is not None:
        assert np.array_equal(rn, x[(np.arange(b, b + n)) % L])
    if mode != "periodic":
        assert (ln is None) == (rank == 0) and (rn is None) == (rank == world - 1)
    y = _slice_with_oracle(O, n, m, d, mode if mode != "periodic" else "periodic", x[a:b], ln, rn)
    ref = O.Filter1D(n, m, d, 1.0, mode).apply(x)[a:b]
    ok = np.array_equal(y.view(np.uint32), ref.view(np.uint32))
    # batch sharding: blocks are disjoint, contiguous and cover everything
    ranges = [sgd.shard_range(65536 + 5, r, world) for r in range(world)]
    ok = ok and ranges[0][0] == 0 and ranges[-1][1] == 65541 and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(int(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,mode", [(2, "periodic"), (2, "reflect"), (3, "polynomial"), (2, "constant")])
def test_partitioned_signal_halo_exchange_gloo(world, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1


class _OracleBandFilter:
    """Stand-in for Savgol2DFilter.apply_band on CPU: the oracle over [halo | band | halo], cropped -- the
    contract tests/test_gpu_2d.py checks for savgol2d_apply_band on the device."""

    def __init__(self, O, nx, ny, order):
        self.o = O.Filter2D(nx, ny, order)

        class _Cfg:
            half_window_y = ny
        self.config = _Cfg()

    def apply_band(self, buf, top, bottom, boundary, out=None, image_row0=0):
        self.row0_seen = image_row0
        full = self.o.apply(buf.numpy().copy(), boundary)
        return torch.from_numpy(full[top: full.shape[0] - bottom].copy())


def _band_worker(rank, world, port, boundary, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from savgol_b200 import dist as sgd
    from oracle import oracle as O
    nx, ny, order = 3, 4, 3
    rng = np.random.default_rng(11)
    img = rng.standard_normal((97, 40)).astype(np.float32)
    a, b = sgd.shard_range(img.shape[0], rank, world)
    f = _OracleBandFilter(O, nx, ny, order)
    y = sgd.apply_image_bands(f, torch.from_numpy(img[a:b].copy()), boundary).numpy()
    ref = O.Filter2D(nx, ny, order).apply(img, boundary)[a:b]
    ok = y.shape == ref.shape and np.array_equal(y.view(np.uint32), ref.view(np.uint32))
    ok = ok and f.row0_seen == a - (ny if rank > 0 else 0)   # the band's position in the image reaches the kernel
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(int(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,boundary", [(2, "constant"), (3, "reflect")])
def test_image_row_bands_halo_exchange_gloo(world, boundary):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_band_worker, args=(r, world, port, boundary, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs)
    assert q.get(timeout=10) == 1
