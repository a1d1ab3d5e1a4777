"""The only collective the path needs: one small all-reduce of loss scalars / per-patch statistics.

Reference: misc.all_reduce_mean (/root/reference/Point-MAE_SA3D/util/misc.py:345-353; called 3-4 times per
step at engine_pretrain_Classifier_SVM.py:297-305, each with a host->device tensor creation, an NCCL call
and an .item() sync) and SmoothedValue.synchronize_between_processes (util/misc.py:41-52).  Here the step's
scalars travel as ONE fp32 vector enqueued on the compute stream; no host sync is forced.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def all_reduce_mean(x):
    """Same contract as the reference: python float / 0-d tensor in, cross-rank mean out (float)."""
    if not is_dist():
        return x
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.as_tensor(x, dtype=torch.float32, device=dev).clone()
    dist.all_reduce(t)
    t /= dist.get_world_size()
    return t.item()


def all_reduce_stats(stats: torch.Tensor, async_op: bool = False):
    """In-place SUM all-reduce of a loss-statistics vector [sum, sum_sq, count, min, max, ...]
    (ops.loss_stats).  min / max slots are reduced with MIN / MAX in the same call sequence.
    Returns the work handle(s) when async_op, else None.  Means are formed by the reader as sum / count."""
    if not is_dist():
        return None
    works = [dist.all_reduce(stats[:3], op=dist.ReduceOp.SUM, async_op=async_op),
             dist.all_reduce(stats[3:4], op=dist.ReduceOp.MIN, async_op=async_op),
             dist.all_reduce(stats[4:5], op=dist.ReduceOp.MAX, async_op=async_op)]
    return works if async_op else None


def all_reduce_scalars(values: Sequence[torch.Tensor]) -> torch.Tensor:
    """Fuse several 0-d device tensors into one vector, mean-all-reduce it once, return the vector (device)."""
    v = torch.stack([t.detach().float().reshape(()) for t in values])
    if is_dist():
        dist.all_reduce(v)
        v /= dist.get_world_size()
    return v
