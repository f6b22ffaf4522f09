"""Multi-GPU work distribution for the file-level codec: images are independent, so a batch is sharded across ranks
with no collective on the data path; torch.distributed is used only to agree on the timing
(barrier + max over ranks) and to gather per-rank statistics.  Works on any backend (NCCL on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence


def shard_by_size(sizes: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of item indices to `world` per-GPU queues (balanced by bytes)."""
    order = sorted(range(len(sizes)), key=lambda i: -sizes[i])
    loads = [0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: loads[k])
        shards[r].append(i)
        loads[r] += sizes[i]
    for s in shards:
        s.sort()
    return shards


def reduce_job_throughput(local_units: float, local_seconds: float, dist=None, device=None):
    """Whole-job throughput = units processed by all ranks / max over ranks of the elapsed time."""
    import torch
    t = torch.tensor([float(local_units), float(local_seconds)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        units = t[0:1].clone()
        secs = t[1:2].clone()
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
        dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        return float(units[0]) / float(secs[0]), float(units[0]), float(secs[0])
    return float(t[0]) / float(t[1]), float(t[0]), float(t[1])
