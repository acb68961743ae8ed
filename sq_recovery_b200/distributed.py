"""Multi-GPU plumbing for the loss path: one process per GPU, batch sharded by sample.

Every sample's loss and its 12 gradients depend only on that sample's parameters (and its own depth image), so
the kernels need no data-path collective.  What crosses ranks is:
  * the scalar loss for logging        -> all-reduce of (sum of per-sample losses, count)
  * IoU's batch-wide counters          -> all-reduce of two int64 (torch/classes.py:437-439)
  * the CNN's parameter gradients      -> DistributedDataParallel (NCCL over NVLink), outside this package
The reference has none of this (single ``cuda:0``, torch/train.py:13).  Works with the nccl backend on GPUs and with
gloo on CPU tensors (used by the world_size-2 CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of ``total`` samples owned by ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def global_mean(local_mean: torch.Tensor, local_count: int, group=None) -> torch.Tensor:
    """Mean over the global batch from per-rank means of possibly different shard sizes (exact for equal shards,
    weighted otherwise).  Differentiable w.r.t. nothing: for logging only."""
    buf = torch.stack([local_mean.detach().double() * local_count,
                       torch.tensor(float(local_count), dtype=torch.float64, device=local_mean.device)])
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[0] / buf[1]


def global_iou(inter: torch.Tensor, union: torch.Tensor, group=None) -> torch.Tensor:
    """Batch-wide IoU (sum of intersections / sum of unions over all ranks), like IoUAccuracy(reduce=True)."""
    buf = torch.stack([inter.sum(), union.sum()]).to(torch.int64)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf[0] / buf[1]


def scale_local_grad(grad: torch.Tensor, local_count: int, global_count: int) -> torch.Tensor:
    """d(global mean loss)/d(local params) from d(local mean loss)/d(local params)."""
    return grad * (local_count / global_count)
