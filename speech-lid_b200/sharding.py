"""Utterance sharding across the GPUs of one box and the global-CMVN statistics exchange.

The reference has no collective on this path (its only parallelism is DDP over the model,
ref: ccml/trainer.py:358-380,426-437).  The front-end shards by utterance -- no halo, no data exchange --
and global CMVN needs exactly one all-reduce of ``[sum_d, sumsq_d, count]`` (2*n_out+1 doubles).
"""
from __future__ import annotations

from typing import List, Sequence

import torch


def lpt_partition(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time greedy: utterance indices per rank, balanced by sample count.
    Deterministic (ties broken by index), so every rank computes the same partition locally."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(lengths[i])
    for s in shards:
        s.sort()
    return shards


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the [2*n_out+1] fp64 statistics vector over ranks (NCCL on GPUs, gloo on CPU tests).  In place."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def finalize_stats(stats: torch.Tensor):
    """stats [2D+1] fp64 -> (mean, unbiased std), the numbers the apply pass uses."""
    D = (stats.numel() - 1) // 2
    n = stats[2 * D]
    mean = stats[:D] / n
    var = (stats[D:2 * D] - stats[:D] * mean) / (n - 1.0)
    return mean, var.clamp_min(0.0).sqrt()
