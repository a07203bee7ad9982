"""Stream sharding across the GPUs of one box.

Streams are independent decode units with per-stream state in HBM, so all frames of a
stream go to one GPU and there is no data-path collective.  The only cross-rank
traffic is the timing reduction of the benchmark (max over ranks).
"""
from __future__ import annotations


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of `n_units` streams owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %d/%d" % (rank, world))
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of(unit: int, n_units: int, world: int) -> int:
    """Rank that owns stream `unit` under shard_range."""
    if not (0 <= unit < n_units):
        raise ValueError("unit out of range")
    base, extra = divmod(n_units, world)
    split = extra * (base + 1)
    if unit < split:
        return unit // (base + 1)
    return extra + (unit - split) // base


def aggregate(units_local: float, seconds_local: float, device=None) -> tuple[float, float]:
    """(sum of units over ranks, max of seconds over ranks) -- the whole-job throughput is their ratio.
    Works on any initialised torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(units_local), float(seconds_local)
    u = torch.tensor([units_local], dtype=torch.float64, device=device)
    t = torch.tensor([seconds_local], dtype=torch.float64, device=device)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(u.item()), float(t.item())
