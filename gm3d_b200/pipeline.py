"""One step of the grouping + reconstruction-loss path on preallocated buffers, captured as a CUDA graph.

    xyz (B,N,3) --group--> neighborhood (B,G,k,3), center (B,G,3)
    loss_pred (B,G) --hard mask--> mask (B,G), M ones per row --select--> patch_index (B*M)
    pred (B*M,k,3) vs neighborhood[patch_index] --Chamfer fwd--> dist/idx, per-patch loss (B*M), scalar loss
    --Chamfer bwd (mean reduction)--> grad_pred (B*M,k,3);   per-patch loss --stats--> (8,) vector
    [world_size > 1: one all-reduce of the stats vector]

This is the per-step sequence of the reference's pre-training loop restricted to the hot path
(/root/reference/Point-MAE_SA3D/engine_pretrain_Classifier_SVM.py:108-118,157-184,297-305).  The step issues
4 kernels through the C ABI; replayed as one graph it has no host work between them.  `HostStagedStep`
adds the host<->device copies from/to pinned memory (the end-to-end arm of bench.py).
"""
from __future__ import annotations

import os

from typing import Optional

import torch

from . import _lib
from .masking import mask_lengths

FUSED_LANES_MAX_B = int(os.environ.get("GM3D_FUSED_LANES_MAX_B", "148"))  # tuning aid
KERNELS_PER_STEP = 4  # unfused: fps, knn_group, hard_mask (+patch index), chamfer fused (fwd + bwd + loss reduction)
FUSED_MAX_N, FUSED_MAX_G = 2048, 1024  # gm3d_cloud_step_f32 serves these; larger clouds use the 4-kernel sequence


class GroupLossStep:
    def __init__(self, B: int, N: int, G: int, k: int, mask_ratio: float = 0.6, epoch: int = 199,
                 total_epoch: int = 400, ratio_cap: float = 0.8, norm: int = 2, device=None, seed: int = 0,
                 rand_offset: int = 0, fused: Optional[bool] = None):
        self.lib = _lib.load()
        can_fuse = N <= FUSED_MAX_N and G <= FUSED_MAX_G and k <= _lib.KNN_MAX_K
        if fused and not can_fuse:
            raise NotImplementedError(f"gm3d_cloud_step_f32 serves N <= {FUSED_MAX_N}, G <= {FUSED_MAX_G}")
        # default: one CTA per cloud pays off while the FPS chain (G dependent rounds, ~0.3 us each beside the workers)
        # is not what the CTA waits for.  Measured on B200 against the kernel sequence spread over forked streams:
        # C1, C2, C4, M2AE level 2 (G <= 128) fused; M2AE level 1 (G = 256: 78 us fused, 62 us as a sequence) and
        # level 0 (G = 512) as a sequence.
        worth = G <= 128
        if fused is None and os.environ.get("GM3D_STEP_FUSED") in ("0", "1"):  # tuning aid (A/B runs of the two paths)
            fused = os.environ["GM3D_STEP_FUSED"] == "1" and can_fuse
        self.fused = (can_fuse and worth) if fused is None else fused
        self.kernels_per_step = 1 if self.fused else KERNELS_PER_STEP
        self.dev = torch.device(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        if self.dev.type != "cuda":
            raise RuntimeError("GroupLossStep runs on CUDA only (gm3d_b200 has no CPU fallback)")
        self.B, self.N, self.G, self.k, self.norm = B, N, G, k, norm
        self.len_keep, self.len_loss = mask_lengths(G, mask_ratio, epoch, total_epoch, True, None, ratio_cap)
        self.M = G - self.len_keep
        self.P = B * self.M
        self.seed, self.rand_offset = seed, rand_offset
        d, f32, i32 = self.dev, torch.float32, torch.int32
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=d)  # noqa: E731
        # inputs: one arena [xyz | pred | loss_pred] so that a host-fed step needs ONE H2D copy
        def carve(arena, specs):
            views, o = [], 0
            for shape, dt in specs:
                n = int(torch.tensor(shape).prod().item()) * torch.empty((), dtype=dt).element_size()
                views.append(arena[o:o + n].view(dt).view(shape))
                o += (n + 255) & ~255
            return views

        def arena_bytes(specs):
            return sum(((int(torch.tensor(sh).prod().item()) * torch.empty((), dtype=dt).element_size()) + 255) & ~255
                       for sh, dt in specs)

        self._in_specs = [((B, N, 3), f32), ((self.P, k, 3), f32), ((B, G), f32)]
        self.in_arena = e((arena_bytes(self._in_specs),), torch.uint8)
        self.xyz, self.pred, self.loss_pred = carve(self.in_arena, self._in_specs)
        # results a host reads back every step: one arena [stats | per_patch | mask] = ONE D2H copy
        self._res_specs = [((_lib.LOSS_STATS_LEN,), f32), ((self.P,), f32), ((B, G), torch.uint8)]
        self.res_arena = e((arena_bytes(self._res_specs),), torch.uint8)
        self.stats, self.per_patch, self.mask = carve(self.res_arena, self._res_specs)
        self._carve = carve
        # outputs
        self.fps_idx = e((B, G), i32)
        self.center = e((B, G, 3), f32)
        self.neighborhood = e((B, G, k, 3), f32)
        self.patch_index = e((self.P,), i32)
        self.dist1, self.dist2 = e((self.P, k), f32), e((self.P, k), f32)
        self.idx1, self.idx2 = e((self.P, k), i32), e((self.P, k), i32)
        self.total = e((1,), f32)
        self.grad_pred = e((self.P, k, 3), f32)
        self.status = torch.zeros((1,), dtype=i32, device=d)
        ws = self.lib.gm3d_workspace_bytes(_lib.OP_GROUP, B, N, G, k)
        self.ws = e((ws,), torch.uint8) if ws else None
        self.cd_ws = torch.zeros((max(self.lib.gm3d_workspace_bytes(_lib.OP_CHAMFER_FWD, self.P, k, k, 0),
                                      self.lib.gm3d_workspace_bytes(_lib.OP_CLOUD_STEP, self.P, 0, 0, 0)),),
                                 dtype=torch.uint8, device=d)  # ticket must start at zero; the kernel re-zeroes it
        self.side = torch.cuda.Stream(d)
        self.graph: Optional[torch.cuda.CUDAGraph] = None

    # algorithmic HBM bytes of one step per cloud (SURVEY App. B formulas; DESIGN.md "Roofline accounting")
    def bytes_per_cloud(self):
        N, G, k, M = self.N, self.G, self.k, self.M
        return {
            "fps": 12 * N + 16 * G,
            "knn_group": 12 * N + 12 * G + 12 * G * k,           # int64 idx not requested by Group
            # fused fwd+bwd: read pred + target patches once, write dist1/2 + idx1/2, per-patch loss, grad_pred
            "chamfer_fused": 24 * M * k + 16 * M * k + 4 * M + 12 * M * k,
            "hard_mask": 5 * G + 4 * M,
            # the fused step reads xyz, loss_pred, pred once and writes every output once; the neighbourhood,
            # centres and mask never come back from HBM
            "cloud_step": 12 * N + 4 * G + 12 * M * k + (4 * G + 12 * G + 12 * G * k) + (G + 4 * M)
                          + (16 * M * k + 4 * M + 12 * M * k),
        }

    def enqueue(self, flags: int = 0) -> None:
        """Enqueue the kernels of one step on torch's current stream (+ one forked side stream).  `flags`:
        _lib.STEP_OVERLAP_* for the fused kernel (only between steps that share no buffer)."""
        L, p = self.lib, (lambda t: None if t is None else t.data_ptr())
        main = torch.cuda.current_stream(self.dev)
        st = main.cuda_stream
        B, N, G, k, P = self.B, self.N, self.G, self.k, self.P
        chk = _lib.check
        g = (1.0 if self.norm == 2 else 0.5) / (P * k)  # d mean / d dist (L1: the outer /2 folded in)
        if self.fused:  # the whole step in one launch, one CTA per cloud
            chk("gm3d_cloud_step_f32", L.gm3d_cloud_step_f32(
                p(self.xyz), B, N, G, k, p(self.fps_idx), p(self.center), None, p(self.neighborhood), None,
                p(self.loss_pred), self.len_keep, self.len_loss, None, self.seed, self.rand_offset, p(self.mask),
                p(self.patch_index), p(self.pred), g, g, self.norm, p(self.dist1), p(self.dist2), p(self.idx1),
                p(self.idx2), p(self.per_patch), p(self.total), p(self.stats), p(self.grad_pred), flags, p(self.cd_ws), st))
            return
        # the mask depends only on loss_pred: fork it onto a side stream so that (also inside a captured
        # graph) it runs concurrently with FPS, which occupies one SM per cloud and leaves the rest idle
        fork, join = torch.cuda.Event(), torch.cuda.Event()
        fork.record(main)
        self.side.wait_event(fork)
        chk("gm3d_hard_mask_f32", L.gm3d_hard_mask_f32(p(self.loss_pred), B, G, self.len_keep, self.len_loss, None,
                                                       self.seed, self.rand_offset, p(self.mask), p(self.patch_index),
                                                       self.side.cuda_stream))
        join.record(self.side)
        chk("gm3d_group_f32", L.gm3d_group_f32(p(self.xyz), B, N, G, k, p(self.fps_idx), p(self.center), None,
                                               p(self.neighborhood), None, p(self.ws), st))
        main.wait_event(join)
        chk("gm3d_chamfer_fused_f32", L.gm3d_chamfer_fused_f32(
            p(self.pred), p(self.neighborhood), p(self.patch_index), P, k, k, g, g, p(self.dist1), p(self.dist2),
            p(self.idx1), p(self.idx2), p(self.per_patch), p(self.total), p(self.stats), self.norm, p(self.grad_pred),
            None, p(self.cd_ws), st))

    def capture(self, extra=None) -> "GroupLossStep":
        """Capture one step (plus `extra()`, e.g. the stats all-reduce) into a CUDA graph."""
        with torch.cuda.device(self.dev):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):  # warm-up outside capture (cudaFuncSetAttribute, lazy module load)
                self.enqueue()
                if extra is not None:
                    extra()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.enqueue()
                if extra is not None:
                    extra()
            self.graph = g
        return self

    def run(self) -> None:
        if self.graph is not None:
            self.graph.replay()
        else:
            self.enqueue()


class StepRing:
    """A ring of GroupLossStep buffer sets replayed as ONE CUDA graph.  The steps share no buffer, so the fused
    kernels are chained with programmatic dependent launch: step i+1 starts filling SMs while step i drains
    (launch latency, the cold-cloud prologue and the last-CTA loss reduction of one step hide under the next).

    `reduce_stats`: when the process group has more than one rank, the [sum, sum_sq, count] heads of all the
    ring's statistics vectors are packed into `self.head` (n, 3) and SUM-all-reduced ONCE per replay -- the
    reference all-reduces its loss scalars only to log them (util/misc.py:345-353), so batching the ring's
    scalars into one collective changes no result and keeps NCCL out of the kernel-to-kernel chain."""

    def __init__(self, steps, reduce_stats: bool = False):
        self.steps = list(steps)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.reduce_stats = reduce_stats
        self.head = torch.zeros((len(self.steps), 3), dtype=torch.float32, device=self.steps[0].dev) if reduce_stats else None
        self._lanes = None

    def enqueue(self) -> None:
        n = len(self.steps)
        s0 = self.steps[0]
        nl = max(1, min(n, int(os.environ.get("GM3D_RING_LANES", "4"))))  # tuning aid; 1 = one stream
        if s0.fused and nl > 1 and 2 * s0.B <= FUSED_LANES_MAX_B and n >= 2 * nl:
            # Small batches (one CTA per cloud fills a fraction of the 2 x 148 CTA slots): several chains of
            # programmatic-dependent launches side by side, one per forked stream.
            main = torch.cuda.current_stream(s0.dev)
            if self._lanes is None or len(self._lanes) != nl:
                self._lanes = [torch.cuda.Stream(s0.dev) for _ in range(nl)]
            fork = torch.cuda.Event()
            fork.record(main)
            for li, lane in enumerate(self._lanes):
                lane.wait_event(fork)
                mine = self.steps[li::nl]
                with torch.cuda.stream(lane):
                    for i, s in enumerate(mine):
                        s.enqueue((_lib.STEP_OVERLAP_NEXT if i + 1 < len(mine) else 0) | (_lib.STEP_OVERLAP_PREV if i > 0 else 0))
                join = torch.cuda.Event()
                join.record(lane)
                main.wait_event(join)
        elif not s0.fused and nl > 1:
            # Kernel-sequence steps: the steps go round-robin to forked streams, so the FPS chain of one step -- G
            # dependent rounds, one CTA per cloud, issue slots half empty -- runs beside the kNN / Chamfer kernels of
            # another wherever SM resources allow (FPS CTAs of different steps pair up at N <= 2048; at N = 8192 the
            # 20 SMs a 128-cloud FPS leaves free, and the tails of every kernel, get used).  The steps share no buffer.
            main = torch.cuda.current_stream(s0.dev)
            if self._lanes is None or len(self._lanes) != nl:
                self._lanes = [torch.cuda.Stream(s0.dev) for _ in range(nl)]
            fork = torch.cuda.Event()
            fork.record(main)
            for li, lane in enumerate(self._lanes):
                lane.wait_event(fork)
                with torch.cuda.stream(lane):
                    for s in self.steps[li::nl]:
                        s.enqueue(0)
                join = torch.cuda.Event()
                join.record(lane)
                main.wait_event(join)
        else:
            for i, s in enumerate(self.steps):
                f = 0
                if s.fused and n > 1:
                    f = (_lib.STEP_OVERLAP_NEXT if i + 1 < n else 0) | (_lib.STEP_OVERLAP_PREV if i > 0 else 0)
                s.enqueue(f)
        if self.reduce_stats:
            import torch.distributed as dist
            torch.stack([s.stats[:3] for s in self.steps], out=self.head)
            dist.all_reduce(self.head)

    def capture(self) -> "StepRing":
        dev = self.steps[0].dev
        with torch.cuda.device(dev):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self.enqueue()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.enqueue()
            self.graph = g
        return self

    def run(self) -> None:
        if self.graph is not None:
            self.graph.replay()
        else:
            self.enqueue()


class HostStagedStep(GroupLossStep):
    """GroupLossStep fed from / drained to pinned host memory: the caller writes `h_xyz`, `h_pred`,
    `h_loss_pred` (views of one pinned arena), calls run(), and after a stream sync reads `h_stats` (loss
    scalars), `h_per_patch` (the (B,M) loss matrix the loss predictor is trained on) and `h_mask`.  One H2D
    copy, the step, one D2H copy -- all three part of the captured graph."""

    def __init__(self, *a, cloud_only: bool = False, **kw):
        """cloud_only: only the point clouds cross PCIe each step; `pred` and `loss_pred` stay where the reference
        produces them -- on the device, as decoder / loss-predictor outputs
        (engine_pretrain_Classifier_SVM.py:108-118,157-164)."""
        super().__init__(*a, **kw)
        self.cloud_only = cloud_only
        self.h_in = torch.empty(self.in_arena.shape, dtype=torch.uint8, pin_memory=True)
        self.h_res = torch.empty(self.res_arena.shape, dtype=torch.uint8, pin_memory=True)
        self.h_xyz, self.h_pred, self.h_loss_pred = self._carve(self.h_in, self._in_specs)
        self.h_stats, self.h_per_patch, self.h_mask = self._carve(self.h_res, self._res_specs)

    @property
    def h2d_bytes(self) -> int:
        return self.xyz.numel() * 4 if self.cloud_only else self.h_in.numel()

    @property
    def d2h_bytes(self) -> int:
        return self.h_res.numel()

    def enqueue(self, flags: int = 0) -> None:
        if self.cloud_only:
            self.xyz.copy_(self.h_xyz, non_blocking=True)
        else:
            self.in_arena.copy_(self.h_in, non_blocking=True)
        super().enqueue(flags)
        self.h_res.copy_(self.res_arena, non_blocking=True)


class HostStagedGroup:
    """Several HostStagedStep slots replayed as ONE graph on one stream: [H2D, step, D2H] x n.  A host loop
    that launches one graph per step is bound by Python / launch overhead long before PCIe; grouping a few
    steps per launch and rotating two or three groups over separate streams keeps the copy engines and the
    SMs busy at the same time.  After `done.synchronize()` every slot's pinned results are valid."""

    def __init__(self, steps):
        self.steps = list(steps)
        self.stream = torch.cuda.Stream(self.steps[0].dev)
        self.done = torch.cuda.Event()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        # zero-copy NumPy views of the pinned statistics (reading a loss costs no torch dispatch)
        self.h_stats_np = [s.h_stats.numpy() for s in self.steps]

    def capture(self) -> "HostStagedGroup":
        with torch.cuda.stream(self.stream):
            for s in self.steps:
                s.enqueue()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.stream):
            for s in self.steps:
                s.enqueue()
        self.graph = g
        return self

    def launch(self) -> None:
        with torch.cuda.stream(self.stream):
            self.graph.replay()
            self.done.record()

    def losses(self):
        self.done.synchronize()
        return [float(v[0]) for v in self.h_stats_np]
