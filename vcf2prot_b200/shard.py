"""Sample sharding across the GPUs of one box: haplotypes are independent, every rank owns a
contiguous sample range and its own engine; the only cross-rank traffic is timing/accounting (no data-path collective).
"""
from __future__ import annotations

from typing import Sequence, Tuple


def sample_range(rank: int, world: int, n_samples: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of samples owned by `rank`; sizes differ by at most one."""
    if not (0 <= rank < world) or n_samples < 0:
        raise ValueError("bad rank/world/n_samples")
    base, extra = divmod(n_samples, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balanced_ranges(weights: Sequence[int], world: int) -> list:
    """Contiguous ranges split by cumulative weight (output bytes per sample) instead of count -- for skewed cohorts
    (split by sum of output bytes if sample-count ranges are > 3 % imbalanced)."""
    total = sum(weights)
    out, lo, acc = [], 0, 0
    for r in range(world):
        target = total * (r + 1) / world
        hi = lo
        while hi < len(weights) and (acc + weights[hi] <= target or r == world - 1):
            acc += weights[hi]
            hi += 1
        out.append((lo, hi))
        lo = hi
    out[-1] = (out[-1][0], len(weights))
    return out


def max_over_ranks(value: float, device=None) -> float:
    """Device time of the slowest rank (every multi-GPU number is the max over ranks)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: int, device=None) -> int:
    """Units processed by all ranks (numerator of the whole-job throughput)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(value)
    t = torch.tensor([value], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
