"""Multi-GPU layout of the inference path: independent replicas.

Clips are independent units: rank r owns clips [r*B/W, (r+1)*B/W) and a full copy of the
weights; there is NO data-path collective.  The only communication is the timing protocol of bench.py: a barrier
before/after the timed region and a MAX over ranks of the device time.  (The reference runs inference in a single
process, inference/predict.py:15; its only multi-GPU code is accelerate's DDP for training.)
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `global_batch` clips; the first (global_batch % world) ranks get one extra."""
    base, extra = divmod(global_batch, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value_ms: float, device=None) -> float:
    """Whole-job time of a replicated step = the slowest rank's device time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value_ms)
    t = torch.tensor([value_ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(device=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)
