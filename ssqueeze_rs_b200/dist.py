"""Host-side plumbing of the N>1 path: one process per GPU, channels
sharded in contiguous blocks, NO data-path collective -- torch.distributed is
used only for the start barrier, the max-over-ranks step time and the sum of
the per-rank unit counts.  Works with the `nccl` backend on GPUs and with
`gloo` on CPU (tests/test_dist_gloo.py)."""
from __future__ import annotations


def rank_channel_block(channels: int, rank: int, world: int):
    """Strong-scaling partition: rank r owns channels [r*C/W, (r+1)*C/W)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return rank * channels // world, (rank + 1) * channels // world


def max_over_ranks(value: float, device=None) -> float:
    """The job's step time is the slowest rank's (device-timed) step time."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def job_throughput(units_this_rank: float, ms_this_rank: float, device=None) -> tuple[float, float]:
    """(units of ALL ranks) / (max-over-ranks time); returns (units per second, ms)."""
    ms = max_over_ranks(ms_this_rank, device)
    units = sum_over_ranks(units_this_rank, device)
    return units / (ms * 1e-3), ms
