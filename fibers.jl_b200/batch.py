"""Host-side helpers for the one-process-per-GPU driver (bench.py under torchrun).

The reconstruction path shards by subject / z-slab with NO data-path collective; the only things
ranks exchange are a barrier and scalar reductions (max of the timed region, voxel counts), which
work on any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import os


def rank_info():
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def reduce_max(value: float, dist=None, device=None) -> float:
    """max over ranks of a python float (identity when not distributed)."""
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def whole_job_throughput(units_this_rank: float, seconds_this_rank: float, dist=None, device=None) -> float:
    """Whole-job units/s = (units summed over ranks) / (max over ranks of the timed region)."""
    total = reduce_sum(units_this_rank, dist, device)
    worst = reduce_max(seconds_this_rank, dist, device)
    return total / worst
