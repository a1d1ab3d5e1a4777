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


class PeerInbox:
    """The inter-GPU inboxes of the per-step statistics all-reduce (include/gm3d.h: gm3d_step_reduce_t): every rank
    allocates `slots` inboxes of GM3D_INBOX_BYTES on its GPU, exports them as a CUDA IPC handle, and maps the
    inboxes of all other ranks (same node, NVLink / NVSwitch peer access).  `step_reduce(slot, head_ptr)` returns the
    descriptor a loss launch takes: its tail pushes this rank's {sum, sum_sq, count} into every inbox and sums what
    the peers pushed into its own -- the reference's per-step `all_reduce_mean` (util/misc.py:345-353) without an
    NCCL launch.  Collective set-up: every rank of `group` must construct it (and close() it) together."""

    def __init__(self, slots: int, group=None, timeout_us: int = 2_000_000):
        import ctypes

        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerInbox needs an initialised process group")
        self.lib = _lib.load()
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.MAX_PEERS:
            raise NotImplementedError(f"PeerInbox serves up to {_lib.MAX_PEERS} ranks of one node")
        self.slots, self.timeout_us = int(slots), int(timeout_us)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        # Every rank takes part in every exchange below even when its own set-up failed, and all ranks raise
        # together: a rank-local exception in front of a collective would leave the others waiting in it.
        ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        rc = self.lib.gm3d_peer_alloc(self.slots * _lib.INBOX_BYTES, ctypes.byref(ptr), handle)
        self._own = ptr.value if rc == 0 else None
        handles = [None] * self.world
        dist.all_gather_object(handles, (rc, handle.raw), group=group)
        self.ptrs, self._opened = [], []
        bad = [(r, c) for r, (c, _) in enumerate(handles) if c != 0]
        rc_open = 0
        if not bad:
            for r, (_, h) in enumerate(handles):
                if r == self.rank:
                    self.ptrs.append(self._own)
                    continue
                q = ctypes.c_void_p()
                rc_open = self.lib.gm3d_peer_open(h, ctypes.byref(q))
                if rc_open != 0:
                    break
                self.ptrs.append(q.value)
                self._opened.append(q.value)
        opened = [None] * self.world
        dist.all_gather_object(opened, rc_open, group=group)
        bad += [(r, c) for r, c in enumerate(opened) if c != 0]
        if bad:
            for q in self._opened:
                self.lib.gm3d_peer_close(q)
            dist.barrier(group=group)
            if self._own is not None:
                self.lib.gm3d_peer_free(self._own)
            self._own, self._opened, self.ptrs = None, [], []
            raise RuntimeError("PeerInbox: peer memory unavailable: " +
                               ", ".join(f"rank {r}: {_lib.strerror(c)}" for r, c in bad))
        self.epoch = torch.zeros((self.slots,), dtype=torch.int32, device=self.dev)   # launch counter per slot
        self.status = torch.zeros((1,), dtype=torch.int32, device=self.dev)           # != 0: a peer never arrived
        self.collected = torch.zeros((self.slots,), dtype=torch.int32, device=self.dev)  # launch count summed, per slot
        dist.barrier(group=group)  # every inbox is mapped everywhere before anybody pushes

    def step_reduce(self, slot: int, head_ptr: int, defer: bool = False, lag: bool = False):
        """Descriptor of step slot `slot`.  defer: the loss launch only pushes (gm3d_step_reduce_collect sums);
        lag: the collect may run beside the next replay's loss launches (it tracks `collected`)."""
        from . import _lib
        if not 0 <= slot < self.slots:
            raise IndexError(f"inbox slot {slot} out of range [0, {self.slots})")
        r = _lib.StepReduce()
        r.head, r.world, r.rank = head_ptr, self.world, self.rank
        for q in range(self.world):
            r.inbox[q] = self.ptrs[q] + slot * _lib.INBOX_BYTES
        r.epoch = self.epoch.data_ptr() + 4 * slot
        r.timeout_us = self.timeout_us
        r.defer = 1 if defer else 0
        r.collected = self.collected.data_ptr() + 4 * slot if lag else None
        r.status = self.status.data_ptr()
        return r

    def close(self) -> None:
        """Collective: unmap the peers' inboxes, then free the own one once nobody can push into it any more."""
        if self._own is None:
            return
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=self.group)
        for q in self._opened:
            self.lib.gm3d_peer_close(q)
        dist.barrier(group=self.group)
        self.lib.gm3d_peer_free(self._own)
        self._own, self._opened, self.ptrs = None, [], []
