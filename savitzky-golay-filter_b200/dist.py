"""Multi-GPU plumbing: one process per GPU, torch.distributed for the (tiny) exchanges.

Batches of independent signals / images / channels shard across ranks with NO collective: every
rank filters its own contiguous block (`shard_range`).  Only a single very long signal is
partitioned into contiguous slices, and then the only communication is the n-sample halo each
slice needs from its ring neighbours (`exchange_halos`, one all_gather of 2n floats per rank --
latency bound, NVLink bandwidth is irrelevant).  The halos are input data, so the exchange can be
issued before/while anything else runs; `apply_partitioned` then calls savgol_apply_halo on the
local slice.  Works with the nccl backend (CUDA tensors) and with gloo (CPU tensors; used by the
CPU test-suite with the oracle as the compute stand-in).
"""
from __future__ import annotations


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [begin, end) of `n_units` independent units owned by `rank`."""
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def exchange_halos(x_local, n: int, periodic: bool, group=None):
    """Returns (left, right): the n samples preceding / following this rank's slice of a signal
    partitioned in rank order, or None at a true signal end (non-periodic).  One all_gather of a
    2n-sample strip [first n | last n] per rank."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if x_local.shape[0] < n:
        raise ValueError("every slice must hold at least half_window samples")
    strip = torch.cat([x_local[:n], x_local[-n:]]).contiguous()
    if world == 1:
        return (strip[n:].clone(), strip[:n].clone()) if periodic else (None, None)
    gathered = torch.empty(world * 2 * n, dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(gathered, strip, group=group)
    prev, nxt = (rank - 1) % world, (rank + 1) % world
    left = gathered[prev * 2 * n + n: prev * 2 * n + 2 * n]
    right = gathered[nxt * 2 * n: nxt * 2 * n + n]
    if not periodic:
        if rank == 0:
            left = None
        if rank == world - 1:
            right = None
    return left, right


def apply_partitioned(filt, x_local, out=None, group=None):
    """Filters this rank's slice of one long signal partitioned over the group (config 3)."""
    periodic = int(filt.config.boundary) == 2
    left, right = exchange_halos(x_local, filt.half_window, periodic, group)
    return filt.apply_halo(x_local, left, right, out=out)


class PeerRing:
    """Ring neighbours' slices of a partitioned signal, mapped into this process (CUDA IPC).

    Set up once per resident buffer: every rank exports (IPC handle, offset, length) of its slice, the
    tuples travel through one all_gather_object, each rank opens its two neighbours.  After that a
    step needs NO communication call at all: `apply` hands savgol_apply_halo pointers into the
    neighbours' HBM and the kernel reads the 2n halo samples over NVLink while staging the slice.
    The caller orders producer and consumer (neighbours' samples complete before `apply`)."""

    def __init__(self, x_local, half_window: int, periodic: bool, group=None):
        import ctypes as C
        import torch.distributed as dist
        from . import lib
        from ._capi import IPC_HANDLE_BYTES

        self._lib = lib()
        self.n = int(half_window)
        self.x = x_local
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        if x_local.shape[0] < self.n:
            raise ValueError("every slice must hold at least half_window samples")
        self._mapped = []          # (mapped pointer, offset) to close
        self.left_ptr = self.right_ptr = None
        if world == 1:
            if periodic:            # the signal wraps onto itself
                self.left_ptr = x_local.data_ptr() + 4 * (x_local.shape[0] - self.n)
                self.right_ptr = x_local.data_ptr()
            return
        handle = (C.c_ubyte * IPC_HANDLE_BYTES)()
        off = C.c_size_t(0)
        if self._lib.savgol_b200_ipc_export(C.c_void_p(x_local.data_ptr()), handle, C.byref(off)) != 0:
            raise RuntimeError("savgol_b200_ipc_export failed")
        infos = [None] * world
        dist.all_gather_object(infos, (bytes(handle), int(off.value), int(x_local.shape[0])), group=group)
        opened = {}

        def base_of(r):
            if r not in opened:
                h, o, _len = infos[r]
                p = self._lib.savgol_b200_ipc_open(C.create_string_buffer(h, IPC_HANDLE_BYTES), C.c_size_t(o))
                if not p:
                    raise RuntimeError(f"savgol_b200_ipc_open failed for rank {r}")
                opened[r] = int(p)
                self._mapped.append((int(p), o))
            return opened[r]

        prev, nxt = (rank - 1) % world, (rank + 1) % world
        if periodic or rank > 0:
            self.left_ptr = base_of(prev) + 4 * (infos[prev][2] - self.n)
        if periodic or rank < world - 1:
            self.right_ptr = base_of(nxt)

    def apply(self, filt, out):
        """Filters the local slice; halos are read from the neighbours' memory inside the kernel."""
        import ctypes as C
        rc = self._lib.savgol_apply_halo(filt.handle, C.c_void_p(self.x.data_ptr()), C.c_void_p(out.data_ptr()), self.x.shape[0],
                                         C.c_void_p(self.left_ptr) if self.left_ptr else None,
                                         C.c_void_p(self.right_ptr) if self.right_ptr else None)
        if rc != 0:
            raise RuntimeError("savgol_apply_halo failed (see stderr)")
        return out

    def close(self):
        import ctypes as C
        for p, o in self._mapped:
            self._lib.savgol_b200_ipc_close(C.c_void_p(p), C.c_size_t(o))
        self._mapped = []


def exchange_halo_rows(band, ny: int, group=None):
    """Row-band sharding of one image: returns (top, bottom) = the ny image rows above / below this rank's
    band (bands in rank order), or None at the image border.  One all_gather of a [2*ny, cols] strip."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if band.shape[0] < ny:
        raise ValueError("every band must hold at least half_window_y rows")
    if world == 1:
        return None, None
    strip = torch.cat([band[:ny], band[-ny:]]).contiguous()
    flat = torch.empty(world * strip.numel(), dtype=band.dtype, device=band.device)
    dist.all_gather_into_tensor(flat, strip.reshape(-1), group=group)
    gathered = flat.view((world,) + tuple(strip.shape))
    top = gathered[rank - 1, ny:] if rank > 0 else None
    bottom = gathered[rank + 1, :ny] if rank < world - 1 else None
    return top, bottom


def apply_image_bands(filt, band, boundary="constant", out=None, group=None, row_begin=None):
    """Filters this rank's horizontal band of ONE image sharded by rows over the group: halo rows from the
    ring neighbours, then savgol2d_apply_band_at.  The result equals the same rows of the whole-image filter,
    bit for bit.  ``row_begin`` = image row of band[0] (default: the sum of the lower ranks' band heights)."""
    import torch
    import torch.distributed as dist

    ny = int(filt.config.half_window_y)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if row_begin is None:
        row_begin = 0
        if world > 1:
            heights = [None] * world
            dist.all_gather_object(heights, int(band.shape[0]), group=group)
            row_begin = sum(heights[:rank])
    top, bottom = exchange_halo_rows(band, ny, group)
    parts = ([top] if top is not None else []) + [band] + ([bottom] if bottom is not None else [])
    buf = torch.cat(parts).contiguous() if len(parts) > 1 else band
    t = ny if top is not None else 0
    return filt.apply_band(buf, t, ny if bottom is not None else 0, boundary, out=out, image_row0=row_begin - t)
