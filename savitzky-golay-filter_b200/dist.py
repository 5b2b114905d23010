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
