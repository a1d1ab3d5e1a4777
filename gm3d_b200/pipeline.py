"""One step of the grouping + reconstruction-loss path on preallocated buffers, captured as a CUDA graph.

    xyz (B,N,3) --group--> neighborhood (B,G,k,3), center (B,G,3)
    loss_pred (B,G) --hard mask--> mask (B,G), M ones per row, patch_index (B*M)
    pred (B*M,k,3) vs neighborhood[patch_index] --Chamfer fwd + bwd of the mean--> dist/idx, per-patch loss (B*M),
    scalar loss, grad_pred (B*M,k,3), statistics (8,) [+ the cross-rank sum of their head]

This is the per-step sequence of the reference's pre-training loop restricted to the hot path
(/root/reference/Point-MAE_SA3D/engine_pretrain_Classifier_SVM.py:108-118,157-184,297-305).  In that loop
`loss_pred` comes out of the teacher network, which consumes the grouping, and `pred` out of the student, which
consumes the grouping AND the mask.  The step therefore runs as the three operator launches a training loop can
interleave with its networks ("dataflow" path, the default and what bench.py times):

    gm3d_group_f32 / gm3d_cloud_step_f32(pred = NULL)  ->  gm3d_hard_mask_f32  ->  gm3d_chamfer_fused_f32

The "single" path issues the same work as ONE launch (gm3d_cloud_step_f32 with `pred`); it serves callers whose
prediction does not depend on this step's grouping / mask and is reported by bench.py as `single_launch`.
`StepRing` replays a ring of independent steps as one graph: every step is the stream-ordered sequence mask -> group ->
loss, the steps go round-robin to forked streams, so the latency-bound loss launch of one step runs beside the sampling
chains of the others.  GM3D_NVTX=1 wraps the enqueue stages in NVTX ranges.  `HostStagedStep` adds the host<->device copies from/to
pinned memory (the end-to-end arm of bench.py).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional

import torch

from . import _lib
from .masking import mask_lengths

FUSED_LANES_MAX_B = int(os.environ.get("GM3D_FUSED_LANES_MAX_B", "148"))  # tuning aid
_NVTX = os.environ.get("GM3D_NVTX", "0") == "1"  # NVTX ranges around the enqueue stages (nsys / ncu --nvtx timelines)


class _nvtx:
    """`with _nvtx("gm3d.group"):` -- a named NVTX range when GM3D_NVTX=1, nothing otherwise."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False
FUSED_MAX_N, FUSED_MAX_G = 2048, 1024  # gm3d_cloud_step_f32 serves these; larger clouds use fps + knn_group


class GroupLossStep:
    def __init__(self, B: int, N: int, G: int, k: int, mask_ratio: float = 0.6, epoch: int = 199,
                 total_epoch: int = 400, ratio_cap: float = 0.8, norm: int = 2, device=None, seed: int = 0,
                 rand_offset: int = 0, fused: Optional[bool] = None, path: Optional[str] = None,
                 xyz: Optional[torch.Tensor] = None):
        """path: 'dataflow' (default: group -> mask -> Chamfer launches) or 'single' (one gm3d_cloud_step_f32 launch;
        `fused=True` is the same request, `fused=False` forces the dataflow path with separate fps / kNN kernels).
        xyz: use this (B,N,3) tensor as the input cloud instead of allocating one (a hierarchy level that groups the
        previous level's centres: M2AEStep)."""
        self.lib = _lib.load()
        can_fuse = N <= FUSED_MAX_N and G <= FUSED_MAX_G and k <= _lib.KNN_MAX_K
        if path is None:
            path = "single" if fused else "dataflow"
        if path not in ("dataflow", "single"):
            raise ValueError(f"path must be 'dataflow' or 'single', got {path!r}")
        if path == "single" and not can_fuse:
            raise NotImplementedError(f"gm3d_cloud_step_f32 serves N <= {FUSED_MAX_N}, G <= {FUSED_MAX_G}")
        self.path = path
        self.fused = path == "single"
        # Grouping by the per-cloud kernel (sampling and patch selection overlapped inside one CTA) pays off while the
        # FPS chain (G dependent rounds) is not what the CTA waits for; measured on B200: G <= 128 (C1, C2, C4, M2AE
        # level 2) per-cloud kernel, M2AE levels 0 / 1 and N > 2048 as fps + knn_group.
        self.group_per_cloud = can_fuse and G <= 128 and fused is not False
        # the mask and Chamfer launches can join a programmatic-dependent-launch chain (StepRing) for these shapes
        self.chainable = G <= 64 and 16 < k <= 32
        self.kernels_per_step = 1 if self.fused else (3 if self.group_per_cloud else 4)
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

        if xyz is None:
            self._in_specs = [((B, N, 3), f32), ((self.P, k, 3), f32), ((B, G), f32)]
            self.in_arena = e((arena_bytes(self._in_specs),), torch.uint8)
            self.xyz, self.pred, self.loss_pred = carve(self.in_arena, self._in_specs)
        else:
            if tuple(xyz.shape) != (B, N, 3) or xyz.dtype != f32 or not xyz.is_contiguous() or xyz.device != d:
                raise ValueError(f"xyz must be a contiguous float32 ({B},{N},3) tensor on {d}")
            self._in_specs = [((self.P, k, 3), f32), ((B, G), f32)]
            self.in_arena = e((arena_bytes(self._in_specs),), torch.uint8)
            self.pred, self.loss_pred = carve(self.in_arena, self._in_specs)
            self.xyz = xyz
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
            # Group.forward in one launch: the cloud once in, fps_idx + centres + neighbourhood out
            "group": 12 * N + 4 * G + 12 * G + 12 * G * k,
            # fused fwd+bwd: read pred + target patches once, write dist1/2 + idx1/2, per-patch loss, grad_pred
            "chamfer_fused": 24 * M * k + 16 * M * k + 4 * M + 12 * M * k,
            "hard_mask": 5 * G + 4 * M,
            # the single launch reads xyz, loss_pred, pred once and writes every output once; the neighbourhood,
            # centres and mask never come back from HBM
            "cloud_step": 12 * N + 4 * G + 12 * M * k + (4 * G + 12 * G + 12 * G * k) + (G + 4 * M)
                          + (16 * M * k + 4 * M + 12 * M * k),
        }

    def step_bytes_per_cloud(self) -> int:
        b = self.bytes_per_cloud()
        if self.fused:
            return b["cloud_step"]
        return (b["group"] if self.group_per_cloud else b["fps"] + b["knn_group"]) + b["hard_mask"] + b["chamfer_fused"]

    # ---- the three operator launches of the dataflow path (each on torch's CURRENT stream)
    def enqueue_group(self, flags: int = 0) -> None:
        L, p = self.lib, (lambda t: None if t is None else t.data_ptr())
        st = torch.cuda.current_stream(self.dev).cuda_stream
        B, N, G, k = self.B, self.N, self.G, self.k
        with _nvtx("gm3d.group"):
            if self.group_per_cloud:  # Group.forward by the per-cloud kernel; `flags` chain independent steps
                _lib.check("gm3d_cloud_step_f32", L.gm3d_cloud_step_f32(
                    p(self.xyz), B, N, G, k, p(self.fps_idx), p(self.center), None, p(self.neighborhood), None,
                    None, 0, 0, None, 0, 0, None, None, None, 0.0, 0.0, 2, None, None, None, None, None, None, None, None,
                    flags, None, None, st))
            else:
                _lib.check("gm3d_group_f32", L.gm3d_group_f32(p(self.xyz), B, N, G, k, p(self.fps_idx), p(self.center), None,
                                                              p(self.neighborhood), None, p(self.ws), st))

    def enqueue_mask(self, flags: int = 0) -> None:
        p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        st = torch.cuda.current_stream(self.dev).cuda_stream
        with _nvtx("gm3d.mask"):
            _lib.check("gm3d_hard_mask_f32", self.lib.gm3d_hard_mask_f32(
                p(self.loss_pred), self.B, self.G, self.len_keep, self.len_loss, None, self.seed, self.rand_offset,
                p(self.mask), p(self.patch_index), flags, st))

    def _gscale(self) -> float:
        return (1.0 if self.norm == 2 else 0.5) / (self.P * self.k)  # d mean / d dist (L1: the outer /2 folded in)

    def enqueue_loss(self, reduce: Optional[_lib.StepReduce] = None, flags: int = 0) -> None:
        p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        st = torch.cuda.current_stream(self.dev).cuda_stream
        g, k = self._gscale(), self.k
        with _nvtx("gm3d.loss"):
            _lib.check("gm3d_chamfer_fused_f32", self.lib.gm3d_chamfer_fused_f32(
                p(self.pred), p(self.neighborhood), p(self.patch_index), self.P, k, k, g, g, p(self.dist1), p(self.dist2),
                p(self.idx1), p(self.idx2), p(self.per_patch), p(self.total), p(self.stats), self.norm, p(self.grad_pred),
                None, ctypes.byref(reduce) if reduce is not None else None, flags, p(self.cd_ws), st))

    def enqueue(self, flags: int = 0, reduce: Optional[_lib.StepReduce] = None) -> None:
        """Enqueue the kernels of one step on torch's current stream (+ one forked side stream for the mask).
        `flags`: _lib.STEP_OVERLAP_* for the single-launch kernel (only between steps that share no buffer)."""
        L, p = self.lib, (lambda t: None if t is None else t.data_ptr())
        main = torch.cuda.current_stream(self.dev)
        if self.fused:  # the whole step in one launch, one CTA per cloud
            B, N, G, k, g = self.B, self.N, self.G, self.k, self._gscale()
            _lib.check("gm3d_cloud_step_f32", L.gm3d_cloud_step_f32(
                p(self.xyz), B, N, G, k, p(self.fps_idx), p(self.center), None, p(self.neighborhood), None,
                p(self.loss_pred), self.len_keep, self.len_loss, None, self.seed, self.rand_offset, p(self.mask),
                p(self.patch_index), p(self.pred), g, g, self.norm, p(self.dist1), p(self.dist2), p(self.idx1),
                p(self.idx2), p(self.per_patch), p(self.total), p(self.stats), p(self.grad_pred), flags,
                ctypes.byref(reduce) if reduce is not None else None, p(self.cd_ws), main.cuda_stream))
            return
        # the mask depends only on loss_pred: fork it onto a side stream so that (also inside a captured graph) it
        # runs beside the grouping
        fork, join = torch.cuda.Event(), torch.cuda.Event()
        fork.record(main)
        self.side.wait_event(fork)
        with torch.cuda.stream(self.side):
            self.enqueue_mask()
        join.record(self.side)
        self.enqueue_group(0)
        main.wait_event(join)
        self.enqueue_loss(reduce)

    def capture(self, extra=None, reduce: Optional[_lib.StepReduce] = None) -> "GroupLossStep":
        """Capture one step (plus `extra()`, e.g. a stats all-reduce) into a CUDA graph."""
        with torch.cuda.device(self.dev):
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):  # warm-up outside capture (cudaFuncSetAttribute, lazy module load)
                self.enqueue(0, reduce)
                if extra is not None:
                    extra()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.enqueue(0, reduce)
                if extra is not None:
                    extra()
            self.graph = g
        return self

    def run(self) -> None:
        if self.graph is not None:
            self.graph.replay()
        else:
            self.enqueue()


class M2AEStep:
    """One step of the Point-M2AE+GM3D hierarchy (BASELINE config[2];
    /root/reference/Point-M2AE_SA3D/cfgs/config_Point_M2AE.yaml:57-69: num_groups 512/256/64, group_sizes 16/8/8,
    mask_ratio 0.8): three chained Group levels -- level 0 groups the raw cloud, level l+1 groups the CENTRES of
    level l -- and per level the hard-patch mask (M_l = G_l - int(G_l (1 - ratio))) and Chamfer-L2 forward + backward
    of the masked patches against a prediction.  Level l+1's input cloud IS level l's `center` tensor (no copy).
    Quacks like a GroupLossStep for StepRing (kernel-sequence lanes) and HostStagedGroup."""

    fused = False
    group_per_cloud = False
    path = "dataflow"

    def __init__(self, B: int, N: int, groups=(512, 256, 64), sizes=(16, 8, 8), mask_ratio: float = 0.8, device=None,
                 seed: int = 0, rand_offset: int = 0, **kw):
        self.B, self.N = B, N
        self.levels: List[GroupLossStep] = []
        n, xyz = N, None
        for li, (g, k) in enumerate(zip(groups, sizes)):
            s = GroupLossStep(B, n, g, k, mask_ratio, device=device, seed=seed + li, rand_offset=rand_offset, xyz=xyz, **kw)
            self.levels.append(s)
            n, xyz = g, s.center
        self.dev, self.lib = self.levels[0].dev, self.levels[0].lib
        self.kernels_per_step = sum(s.kernels_per_step for s in self.levels)
        self.stats = self.levels[0].stats  # StepRing publishes level 0's statistics head
        self.graph: Optional[torch.cuda.CUDAGraph] = None

    @property
    def xyz(self):
        return self.levels[0].xyz

    def step_bytes_per_cloud(self) -> int:
        return sum(s.step_bytes_per_cloud() for s in self.levels)

    def flop_per_cloud(self) -> int:
        return sum(8 * ((s.G - 1) * s.N + s.G * s.N + s.M * s.k * s.k) for s in self.levels)

    def enqueue(self, flags: int = 0, reduce: Optional[_lib.StepReduce] = None) -> None:
        """masks (they need loss_pred only) on a forked stream; the three group launches in order (level l+1 reads
        level l's centres); then the three Chamfer launches."""
        main = torch.cuda.current_stream(self.dev)
        side = self.levels[0].side
        fork, join = torch.cuda.Event(), torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            for s in self.levels:
                s.enqueue_mask()
        join.record(side)
        for s in self.levels:
            s.enqueue_group(0)
        main.wait_event(join)
        for li, s in enumerate(self.levels):
            s.enqueue_loss(reduce if li == 0 else None)

    capture = GroupLossStep.capture
    run = GroupLossStep.run


class StepRing:
    """A ring of GroupLossStep buffer sets replayed as ONE CUDA graph.  The steps share no buffer, so consecutive
    steps overlap: the per-cloud launches chain with programmatic dependent launch (step i+1 fills SMs while step i
    drains), and on the dataflow path the mask / Chamfer launches of step i run on forked streams behind their own
    group launch, beside the group launch of step i+1.

    reduce -- what happens to the [sum, sum_sq, count] head of every step's loss statistics (row i of `self.head`,
    (n, 4) f32, written by the tail of the step's loss launch; column 3 = number of ranks summed):
      'none'  this rank's values;
      'peer'  per step over peer memory (`inbox`: dist.PeerInbox): the tail of every step's loss launch PUSHES this
              rank's values into every rank's inbox over NVLink; one small launch per replay, forked at the START of
              the graph, sums -- in rank order -- what all ranks pushed during the PREVIOUS replay while this replay
              computes.  The reference's per-step all_reduce_mean (util/misc.py:345-353) without an NCCL launch or a
              host sync, and no rank waits for another unless that one is a whole replay behind.  `head` therefore
              lags one replay; flush() after the last replay brings it up to date;
      'peer_sync'  as 'peer', but the sums of a replay are formed by a launch at the END of the same replay (no lag;
              every replay ends with one wait for the slowest rank);
      'peer_step'  as 'peer', but every loss launch itself waits for its step's pushes (the sum is in `head` the
              moment that launch completes; the ranks then run in lock-step at step granularity);
      'nccl'  this rank's values, then ONE NCCL SUM all-reduce of the whole (n, 4) tensor at the end of the graph
              (batched: the reference reduces these scalars only to log them).
    Every rank must replay the ring the same number of times."""

    def __init__(self, steps, reduce: str = "none", inbox=None, reduce_stats: Optional[bool] = None,
                 schedule: str = "lanes"):
        """schedule (dataflow path with the per-cloud group kernel): 'lanes' = steps round-robin on forked streams
        (default, fastest); 'chain' = one stream of launches chained by programmatic dependent launch."""
        self.steps: List[GroupLossStep] = list(steps)
        if schedule not in ("lanes", "chain"):
            raise ValueError(f"schedule must be 'lanes' or 'chain', got {schedule!r}")
        self.schedule = schedule
        if reduce_stats is not None:  # round-1 spelling
            reduce = "nccl" if reduce_stats else "none"
        if reduce not in ("none", "peer", "peer_sync", "peer_step", "nccl"):
            raise ValueError(f"reduce must be 'none', 'peer', 'peer_sync', 'peer_step' or 'nccl', got {reduce!r}")
        if reduce.startswith("peer") and inbox is None:
            raise ValueError("reduce='peer' needs a dist.PeerInbox with one slot per step")
        self.reduce = reduce
        self.inbox = inbox
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        dev = self.steps[0].dev
        self.head = torch.zeros((len(self.steps), 4), dtype=torch.float32, device=dev)
        self._red = []
        for i in range(len(self.steps)):
            if reduce.startswith("peer"):
                r = inbox.step_reduce(i, self.head[i].data_ptr(), defer=reduce != "peer_step", lag=reduce == "peer")
            else:
                r = _lib.StepReduce()
                r.head, r.world, r.rank = self.head[i].data_ptr(), 1, 0
            self._red.append(r)
        self._lanes = None
        self._side = None

    def _streams(self, n):
        dev = self.steps[0].dev
        if self._lanes is None or len(self._lanes) != n:
            self._lanes = [torch.cuda.Stream(dev) for _ in range(n)]
        return self._lanes

    @staticmethod
    def _chain_flags(i, n):
        return (_lib.STEP_OVERLAP_NEXT if i + 1 < n else 0) | (_lib.STEP_OVERLAP_PREV if i > 0 else 0)

    def _collect(self) -> None:
        s0 = self.steps[0]
        _lib.check("gm3d_step_reduce_collect", s0.lib.gm3d_step_reduce_collect(
            ctypes.byref(self._red[0]), len(self.steps), torch.cuda.current_stream(s0.dev).cuda_stream))

    def flush(self) -> None:
        """reduce='peer': sum whatever the ranks have pushed and `head` does not hold yet (call after the last
        replay, on every rank).  Enqueues on the current stream; no host sync."""
        if self.reduce == "peer":
            self._collect()
            self._collect()

    def enqueue(self) -> None:
        n = len(self.steps)
        s0 = self.steps[0]
        main = torch.cuda.current_stream(s0.dev)
        lag_join = None
        if self.reduce == "peer":  # the previous replay's sums, beside this replay's launches
            if self._side is None:
                self._side = torch.cuda.Stream(s0.dev)
            fork = torch.cuda.Event()
            fork.record(main)
            self._side.wait_event(fork)
            with torch.cuda.stream(self._side):
                self._collect()
                lag_join = torch.cuda.Event()
                lag_join.record(self._side)
        nl = max(1, min(n, int(os.environ.get("GM3D_RING_LANES", "4"))))  # tuning aid; 1 = one stream
        if s0.fused and nl > 1 and 2 * s0.B <= FUSED_LANES_MAX_B and n >= 2 * nl:
            # Single-launch steps of small batches (one CTA per cloud fills a fraction of the 2 x 148 CTA slots):
            # several chains of programmatic-dependent launches side by side, one per forked stream.
            lanes = self._streams(nl)
            fork = torch.cuda.Event()
            fork.record(main)
            for li, lane in enumerate(lanes):
                lane.wait_event(fork)
                mine = list(range(li, n, nl))
                with torch.cuda.stream(lane):
                    for j, i in enumerate(mine):
                        self.steps[i].enqueue(self._chain_flags(j, len(mine)), self._red[i])
                join = torch.cuda.Event()
                join.record(lane)
                main.wait_event(join)
        elif s0.fused:
            for i, s in enumerate(self.steps):
                s.enqueue(self._chain_flags(i, n) if n > 1 else 0, self._red[i])
        elif s0.group_per_cloud and self.schedule == "lanes":
            # Dataflow path: every step is the plain stream-ordered sequence mask -> group -> Chamfer; the steps go
            # round-robin to forked streams, so the latency-bound Chamfer launch of one step and the sampling chains
            # of the other steps' group launches share the SMs (measured on B200, C2, us per step: 2 / 3 / 4 / 6 / 8
            # streams -- see DESIGN.md 4.3; programmatic-dependent-launch chains of the three kernels: 26-31).
            nl = max(1, min(n, int(os.environ.get("GM3D_RING_LANES", "12" if 2 * s0.B > FUSED_LANES_MAX_B else "24"))))
            lanes = self._streams(nl)
            fork = torch.cuda.Event()
            fork.record(main)
            for li, lane in enumerate(lanes):
                lane.wait_event(fork)
                with torch.cuda.stream(lane):
                    for i in range(li, n, nl):
                        s = self.steps[i]
                        s.enqueue_mask()
                        s.enqueue_group(_lib.STEP_SHARED_SMS if n > 1 else 0)
                        s.enqueue_loss(self._red[i])
                join = torch.cuda.Event()
                join.record(lane)
                main.wait_event(join)
        elif s0.group_per_cloud and s0.chainable:
            # Dataflow path, per-cloud group kernel: ONE stream of launches chained by programmatic dependent launch,
            #     G0 M0 | G1 C0 M1 | G2 C1 M2 | ... | C(n-1)        (G group, M mask, C Chamfer fwd + bwd)
            # G and M read only the step's own inputs (OVERLAP_PREV: start beside the predecessor, wait for it before
            # retiring); C reads its step's grouping and mask (AFTER_PREV: scheduled early, waits for the predecessor --
            # and transitively for everything before it -- before touching memory).  The loss of step i thus runs
            # beside the sampling chains of steps i+1 and i+2, inside the idle issue slots of their CTAs.
            N_, P_, A_ = _lib.STEP_OVERLAP_NEXT, _lib.STEP_OVERLAP_PREV, _lib.STEP_AFTER_PREV
            seq = [("g", 0), ("m", 0)]
            for i in range(n - 1):
                seq += [("g", i + 1), ("c", i), ("m", i + 1)]
            seq.append(("c", n - 1))
            for j, (kind, i) in enumerate(seq):
                nxt = N_ if j + 1 < len(seq) else 0
                s = self.steps[i]
                if kind == "g":
                    s.enqueue_group(nxt | (P_ if j > 0 else 0))
                elif kind == "m":
                    s.enqueue_mask(nxt | (P_ if j > 0 else 0))
                else:
                    s.enqueue_loss(self._red[i], nxt | (A_ if j > 0 else 0))
        elif nl > 1:
            # Kernel-sequence steps (N > 2048 or G > 128): the steps go round-robin to forked streams, so the FPS chain
            # of one step -- G dependent rounds, one CTA per cloud, issue slots half empty -- runs beside the kNN /
            # Chamfer kernels of another wherever SM resources allow.  The steps share no buffer.
            lanes = self._streams(nl)
            fork = torch.cuda.Event()
            fork.record(main)
            for li, lane in enumerate(lanes):
                lane.wait_event(fork)
                with torch.cuda.stream(lane):
                    for i in range(li, n, nl):
                        self.steps[i].enqueue(0, self._red[i])
                join = torch.cuda.Event()
                join.record(lane)
                main.wait_event(join)
        else:
            for i, s in enumerate(self.steps):
                s.enqueue(0, self._red[i])
        if lag_join is not None:
            main.wait_event(lag_join)
        if self.reduce == "peer_sync":
            self._collect()
        elif self.reduce == "nccl":
            import torch.distributed as dist
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
        with _nvtx(f"gm3d.ring[{len(self.steps)}]"):
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

    def enqueue(self, flags: int = 0, reduce: Optional[_lib.StepReduce] = None) -> None:
        if self.cloud_only:
            self.xyz.copy_(self.h_xyz, non_blocking=True)
        else:
            self.in_arena.copy_(self.h_in, non_blocking=True)
        super().enqueue(flags, reduce)
        self.h_res.copy_(self.res_arena, non_blocking=True)


class HostStagedM2AE(M2AEStep):
    """M2AEStep fed from / drained to pinned host memory: per step ONE H2D copy per level (level 0: cloud + prediction
    + predicted losses; levels 1, 2: prediction + predicted losses -- their clouds are the previous level's centres)
    and ONE D2H copy per level (statistics, per-patch losses, mask)."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.h_in = [torch.empty(s.in_arena.shape, dtype=torch.uint8, pin_memory=True) for s in self.levels]
        self.h_res = [torch.empty(s.res_arena.shape, dtype=torch.uint8, pin_memory=True) for s in self.levels]
        self.h_views = [s._carve(h, s._in_specs) for s, h in zip(self.levels, self.h_in)]  # level 0: xyz, pred, loss_pred
        self.h_stats = self.levels[0]._carve(self.h_res[0], self.levels[0]._res_specs)[0]

    @property
    def h2d_bytes(self) -> int:
        return sum(h.numel() for h in self.h_in)

    @property
    def d2h_bytes(self) -> int:
        return sum(h.numel() for h in self.h_res)

    def enqueue(self, flags: int = 0, reduce: Optional[_lib.StepReduce] = None) -> None:
        for s, h in zip(self.levels, self.h_in):
            s.in_arena.copy_(h, non_blocking=True)
        super().enqueue(flags, reduce)
        for s, h in zip(self.levels, self.h_res):
            h.copy_(s.res_arena, non_blocking=True)


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
        """The steps' mean per-patch losses (statistics slot 5), read from pinned memory after the group's D2H copies."""
        self.done.synchronize()
        return [float(v[5]) for v in self.h_stats_np]
