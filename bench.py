#!/usr/bin/env python
"""Benchmark of the GM3D grouping + reconstruction-loss hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c1|c4|c5] [--impl native|reference]

A "step" is one pass of the path over one batch of B synthetic clouds per GPU:
Group (FPS -> kNN -> gather -> centre) -> hard-patch mask -> masked-patch select -> Chamfer-L2 forward with
the per-patch / scalar reduction -> Chamfer backward (grad w.r.t. the prediction) -> loss statistics
[-> one all-reduce of the statistics vector when N > 1].  Default workload: BASELINE config[1]
(Point-MAE+GM3D pre-train shape, B=128, N=1024, G=64, k=32, M=39 masked patches).

Prints ONE JSON line (see the field notes in DESIGN.md "Measurement").  `value` is device-resident
throughput (inputs already in HBM, each step replayed as one CUDA graph over a ring of input/output buffer
sets larger than L2); `e2e` feeds every step from pinned host memory and reads the loss back;
`--impl reference` times the CPU oracle (reference operator semantics, all host threads) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "clouds/s FPS+kNN-group+Chamfer fwd/bwd"
CONFIGS = {  # name: (B per GPU, N, G, k, mask_ratio, description)
    "c1": (8, 1024, 64, 32, 0.6, "Point-MAE Group+Chamfer-L2 B=8 N=1024 G=64 k=32"),
    "c2": (128, 1024, 64, 32, 0.6, "Point-MAE+GM3D pretrain B=128 N=1024 G=64 k=32 M=39 Chamfer-L2 fwd+bwd + hard-patch mask"),
    "c3l0": (128, 2048, 512, 16, 0.8, "Point-M2AE+GM3D level 0: B=128 N=2048 G=512 k=16 M=410"),
    "c3l1": (128, 512, 256, 8, 0.8, "Point-M2AE+GM3D level 1 (on level-0 centres): B=128 N=512 G=256 k=8 M=205"),
    "c3l2": (128, 256, 64, 8, 0.8, "Point-M2AE+GM3D level 2 (on level-1 centres): B=128 N=256 G=64 k=8 M=52"),
    "c3": (128, 2048, (512, 256, 64), (16, 8, 8), 0.8, "Point-M2AE+GM3D hierarchy B=128 N=2048 G=512/256/64 k=16/8/8 (level l+1 groups "
                                                        "level-l centres), M=410/205/52, multi-scale Chamfer-L2 fwd+bwd + hard-patch masks"),
    "c4": (32, 2048, 128, 32, 0.6, "ScanObjectNN finetune shape B=32 N=2048 G=128 k=32 (+loss for uniformity)"),
    "c5": (128, 8192, 512, 32, 0.6, "scaling sweep shard B=128/GPU N=8192 G=512 k=32 M=308"),
}
L2_BYTES = 126 * 1024 * 1024


def synthetic_batch(B, N, G, k, M, seed):
    """unit-ball clouds with per-cloud scale/translate; pred = gt-like patches + noise."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, N, 3))
    x -= x.mean(axis=1, keepdims=True)
    x /= np.linalg.norm(x, axis=-1).max(axis=1)[:, None, None]
    x = x * rng.uniform(2 / 3, 3 / 2, (B, 1, 3)) + rng.uniform(-0.2, 0.2, (B, 1, 3))
    loss_pred = rng.standard_normal((B, G))
    # placeholder predictions (patch-sized blobs); the timed inputs replace them by near_target_pred() / the CPU twin
    # below -- SURVEY 8(d): pred = gt + 0.02 * randn, which needs the grouping and the mask first
    pred = rng.standard_normal((B * M, k, 3)) * 0.08
    return x.astype(np.float32), loss_pred.astype(np.float32), pred.astype(np.float32)


def near_target_pred(s, seed):
    """the prediction is the masked target patch plus 0.02 * N(0, 1) noise (a decoder output near its
    target: every target point's nearest prediction is then usually its own, as in training, instead of the
    many-to-one matches random blobs produce).  Set-up only: runs the step's mask and group launches, then fills
    `s.pred` on the device; returns the prediction as a NumPy array."""
    import torch
    s.enqueue_mask()
    s.enqueue_group()
    torch.cuda.synchronize(s.dev)
    g = torch.Generator(device=s.dev)
    g.manual_seed(seed)
    gt = s.neighborhood.view(s.B * s.G, s.k, 3)[s.patch_index.long()]
    s.pred.copy_(gt + 0.02 * torch.randn(gt.shape, generator=g, device=s.dev, dtype=torch.float32))
    return s.pred.cpu().numpy()


def near_target_pred_cpu(co, x, lp, G, k, len_keep, len_loss, rk, seed):
    """The CPU arm's twin of near_target_pred (same law; the oracle's own grouping and mask)."""
    nb = co.group(x, G, k)["neighborhood"]
    mask = co.hard_mask(lp, len_keep, len_loss, rk).astype(bool)
    gt = nb[mask]
    return (gt + 0.02 * np.random.default_rng(seed).standard_normal(gt.shape)).astype(np.float32)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------ shared config
def buffer_set_bytes(B, N, G, k, M):
    """Bytes of one step's buffer set (inputs + every output), from the shapes alone (both arms print it)."""
    P = B * M
    return (B * N * 3 * 4 + P * k * 3 * 4 + B * G * 4                      # xyz, pred, loss_pred
            + B * G * 4 + B * G * 3 * 4 + B * G * k * 3 * 4                # fps_idx, centres, neighbourhood
            + B * G + P * 4                                                # mask, patch_index
            + 2 * P * k * 4 + 2 * P * k * 4 + P * 4 + P * k * 3 * 4)       # dist1/2, idx1/2, per_patch, grad_pred


def ring_size(B, N, G, k, M):
    return int(min(64, max(4, -(-2 * L2_BYTES // buffer_set_bytes(B, N, G, k, M)))))


def m2ae_levels(cfg):
    """[(N_l, G_l, k_l, M_l, len_keep, len_loss)] of a hierarchical config (G, k are tuples)."""
    from gm3d_b200.masking import mask_lengths
    B, N, Gs, ks, ratio, _ = cfg
    out, n = [], N
    for g, k in zip(Gs, ks):
        len_keep, len_loss = mask_lengths(g, ratio, 199, 400)
        out.append((n, g, k, g - len_keep, len_keep, len_loss))
        n = g
    return out


def config_dict(cfg):
    """The workload description both arms print (identical keys and values => the driver's same_config check)."""
    from gm3d_b200.masking import mask_lengths
    B, N, G, k, ratio, desc = cfg
    if isinstance(G, tuple):
        lv = m2ae_levels(cfg)
        per_set = sum(buffer_set_bytes(B, n, g, kk, m) for n, g, kk, m, _, _ in lv) - sum(B * n * 12 for n, *_ in lv[1:])
        ring = int(min(64, max(4, -(-2 * L2_BYTES // per_set))))
        return {"workload": desc, "B_per_gpu": B, "N": N, "G": list(G), "k": list(k), "M": [l[3] for l in lv],
                "l2_policy": f"inputs larger than L2: ring of {ring} buffer sets x {per_set / 1e6:.1f} MB"}
    len_keep, _ = mask_lengths(G, ratio, 199, 400)
    M = G - len_keep
    ring = ring_size(B, N, G, k, M)
    return {"workload": desc, "B_per_gpu": B, "N": N, "G": G, "k": k, "M": M,
            "l2_policy": f"inputs larger than L2: ring of {ring} buffer sets x {buffer_set_bytes(B, N, G, k, M) / 1e6:.1f} MB"}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_step(co, x, loss_pred, pred, G, k, len_keep, len_loss, rand_keys, want_center=False):
    """The same step with the CPU oracle (reference operator semantics)."""
    g = co.group(x, G, k)
    mask = co.hard_mask(loss_pred, len_keep, len_loss, rand_keys).astype(bool)
    gt = g["neighborhood"][mask]
    P = gt.shape[0]
    d1, d2, i1, i2 = co.chamfer_fwd(pred[:P], gt)
    pp = co.chamfer_per_patch(d1, d2, 2)
    gd = np.full((P, k), 1.0 / (P * k), dtype=np.float32)
    ga, _ = co.chamfer_bwd(pred[:P], gt, i1, i2, gd, gd)
    if want_center:
        return float(pp.mean()), ga, g["center"]
    return float(pp.mean()), ga


def m2ae_inputs(cfg, seed):
    """Synthetic inputs of the hierarchy: the cloud, and per level the predicted losses and the predicted patches."""
    B, N, Gs, ks, ratio, _ = cfg
    x, _, _ = synthetic_batch(B, N, Gs[0], ks[0], 1, seed)
    rng = np.random.default_rng(seed + 7)
    lps = [rng.standard_normal((B, g)).astype(np.float32) for _, g, _, _, _, _ in m2ae_levels(cfg)]
    preds = [(rng.standard_normal((B * m, k, 3)) * 0.08).astype(np.float32) for _, _, k, m, _, _ in m2ae_levels(cfg)]
    return x, lps, preds


def cpu_step_m2ae(co, cfg, x, lps, preds, rks):
    cloud, loss = x, 0.0
    for (n, g, k, m, len_keep, len_loss), lp, pred, rk in zip(m2ae_levels(cfg), lps, preds, rks):
        loss_l, _, cloud = cpu_step(co, cloud, lp, pred, g, k, len_keep, len_loss, rk, want_center=True)  # next level: the centres
        loss += loss_l
    return loss


def time_cpu_m2ae(cfg, budget_s: float, steps=None, warmup: int = 1):
    from oracle import c_oracle as co
    B = cfg[0]
    x, lps, _ = m2ae_inputs(cfg, 1234)
    lv = m2ae_levels(cfg)
    rks = [np.random.default_rng(5 + i).random((B, l[1])).astype(np.float32) for i, l in enumerate(lv)]
    preds, cloud = [], x
    for i, ((n, g, k, m, len_keep, len_loss), lp, rk) in enumerate(zip(lv, lps, rks)):
        preds.append(near_target_pred_cpu(co, cloud, lp, g, k, len_keep, len_loss, rk, 1235 + i))
        cloud = co.group(cloud, g, k)["center"]
    sample_B = max(co.num_threads(), min(B, 16))
    sub = lambda b: (x[:b], [a[:b] for a in lps], [p[: b * l[3]] for p, l in zip(preds, lv)], [r[:b] for r in rks])  # noqa: E731
    t0 = time.perf_counter()
    cpu_step_m2ae(co, cfg, *sub(sample_B))
    one = time.perf_counter() - t0
    if steps is None:
        steps = max(2, min(500, int(budget_s / max(one, 1e-4))))
    elif one * (steps + warmup) > budget_s:
        sample_B = max(co.num_threads(), int(sample_B * budget_s / (one * (steps + warmup))))
    args = sub(sample_B)
    for _ in range(warmup):
        cpu_step_m2ae(co, cfg, *args)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step_m2ae(co, cfg, *args)
    dt = time.perf_counter() - t0
    return {"value": sample_B * steps / dt, "unit": "clouds/s", "cores": co.num_threads(), "kind": "port",
            "sample": f"{steps} steps x {sample_B} clouds of the {B}-cloud batch, three chained levels, C oracle on "
                      f"{co.num_threads()} host threads, {dt:.1f} s"}, dt / steps


def time_cpu(cfg, budget_s: float, steps=None, warmup: int = 1):
    """Time the oracle on a bounded sample: whole batches of `sample_B` clouds; returns clouds/s."""
    from oracle import c_oracle as co
    from gm3d_b200.masking import mask_lengths
    if isinstance(cfg[2], tuple):
        return time_cpu_m2ae(cfg, budget_s, steps, warmup)
    B, N, G, k, ratio, _ = cfg
    len_keep, len_loss = mask_lengths(G, ratio, 199, 400)
    M = G - len_keep
    x, lp, _ = synthetic_batch(B, N, G, k, M, 1234)
    rk = np.random.default_rng(5).random((B, G)).astype(np.float32)
    pred = near_target_pred_cpu(co, x, lp, G, k, len_keep, len_loss, rk, 1235)
    t0 = time.perf_counter()
    cpu_step(co, x, lp, pred, G, k, len_keep, len_loss, rk)  # warm-up + calibration
    one = time.perf_counter() - t0
    sample_B = B
    if steps is None:
        steps = max(3, min(2000, int(budget_s / max(one, 1e-4))))  # ~budget_s seconds of CPU work
    elif one * (steps + warmup) > budget_s:  # shrink the per-step sample, keep whole clouds
        sample_B = max(co.num_threads(), int(B * budget_s / (one * (steps + warmup))))
        sample_B = min(B, sample_B)
    xs, lps, preds, rks = x[:sample_B], lp[:sample_B], pred[: sample_B * M], rk[:sample_B]
    for _ in range(warmup):
        cpu_step(co, xs, lps, preds, G, k, len_keep, len_loss, rks)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(co, xs, lps, preds, G, k, len_keep, len_loss, rks)
    dt = time.perf_counter() - t0
    return {"value": sample_B * steps / dt, "unit": "clouds/s", "cores": co.num_threads(), "kind": "port",
            "sample": f"{steps} steps x {sample_B} clouds of the {B}-cloud batch (N={N},G={G},k={k},M={M}), "
                      f"C oracle on {co.num_threads()} host threads, {dt:.1f} s"}, dt / steps


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms = time_cpu(cfg, budget_s=150.0, steps=args.steps, warmup=max(1, min(args.warmup, 3)))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "clouds/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(cfg),
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ native arm
_STAGE = ["start"]


def _stage(msg):
    _STAGE[0] = msg
    if os.environ.get("GM3D_BENCH_TRACE") or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        sys.stderr.write(f"[bench r{os.environ.get('RANK', '0')} {time.time() % 1000:8.3f}] {msg}\n")
        sys.stderr.flush()


def _watchdog(seconds: float):
    """A hung collective must end the run with a rank-tagged stage, not occupy the box until the driver kills it."""
    import faulthandler

    def fire():
        sys.stderr.write(f"[bench r{os.environ.get('RANK', '0')}] WATCHDOG after {seconds:.0f} s in stage '{_STAGE[0]}'\n")
        faulthandler.dump_traceback(file=sys.stderr)
        sys.stderr.flush()
        os._exit(3)

    t = threading.Timer(seconds, fire)
    t.daemon = True
    t.start()
    return t


class _near_gpu:
    """Allocate pinned host memory on the NUMA node the GPU hangs off: inside the block the thread runs on the CPUs
    listed in /sys/bus/pci/devices/<gpu>/local_cpulist (first touch pins the pages there), afterwards the original
    affinity is back.  A host-feed detail a deployment would set with numactl; a no-op wherever it cannot apply."""

    def __init__(self, index: int):
        self.index, self.saved, self.info = index, None, "unchanged"

    def __enter__(self):
        try:
            import torch
            pr = torch.cuda.get_device_properties(self.index)
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
                spec = f.read().strip()
            cpus = set()
            for part in spec.split(","):
                if "-" in part:
                    lo, hi = part.split("-")
                    cpus.update(range(int(lo), int(hi) + 1))
                elif part:
                    cpus.add(int(part))
            have = os.sched_getaffinity(0)
            want = cpus & have
            if want and want != have:
                self.saved = have
                os.sched_setaffinity(0, want)
                self.info = f"pinned buffers allocated from cpus {spec} (GPU {bdf})"
            else:
                self.info = f"GPU {bdf} local cpus {spec}: " + ("all of this process's cpus are local" if want else "none available to this process")
        except Exception as e:  # noqa: BLE001 -- purely advisory
            self.info = f"unavailable ({type(e).__name__})"
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            os.sched_setaffinity(0, self.saved)
        return False


def oracle_check_step(s, x, lp, pred):
    """Compare one ring slot (its buffers as the timed launches left them) with the CPU oracle: indices, mask and
    arg-mins bit-exact, losses / gradients within 1e-5 relative.  Returns 'ok' or raises AssertionError."""
    from oracle import c_oracle as co
    w = co.group(x, s.G, s.k)
    assert np.array_equal(s.fps_idx.cpu().numpy(), w["fps_idx"]), "fps_idx"
    assert np.array_equal(s.center.cpu().numpy(), w["center"]), "center"
    assert np.array_equal(s.neighborhood.cpu().numpy(), w["neighborhood"]), "neighborhood"
    mask = s.mask.cpu().numpy()
    assert (mask.sum(1) == s.M).all(), "mask cardinality"
    if s.len_loss > 0:
        order = np.argsort(lp, axis=1, kind="stable")[:, s.G - s.len_loss:]
        assert (np.take_along_axis(mask, order, axis=1) == 1).all(), "mask misses a top-loss patch"
    assert np.array_equal(s.patch_index.cpu().numpy(), np.flatnonzero(mask.reshape(-1))), "patch_index"
    gt = w["neighborhood"][mask.astype(bool)]
    d1, d2, i1, i2 = co.chamfer_fwd(pred, gt)
    assert np.array_equal(s.dist1.cpu().numpy(), d1) and np.array_equal(s.dist2.cpu().numpy(), d2), "chamfer dist"
    assert np.array_equal(s.idx1.cpu().numpy(), i1) and np.array_equal(s.idx2.cpu().numpy(), i2), "chamfer idx"
    pp = co.chamfer_per_patch(d1, d2, 2)
    assert np.allclose(s.per_patch.cpu().numpy(), pp, rtol=1e-5, atol=0), "per_patch"
    assert abs(s.total.item() - pp.mean()) <= 1e-5 * abs(pp.mean()), "total"
    g = np.full((s.P, s.k), 1.0 / (s.P * s.k), dtype=np.float32)
    ga, _ = co.chamfer_bwd(pred, gt, i1, i2, g, g)
    got = s.grad_pred.cpu().numpy()
    assert np.abs(got - ga).max() <= 1e-5 * np.abs(ga).max(), "grad_pred"
    return "ok"


def run_native(args, cfg):
    import datetime

    import torch
    import torch.distributed as dist

    from gm3d_b200 import _lib
    from gm3d_b200.pipeline import GroupLossStep, HostStagedGroup, HostStagedStep, StepRing

    _lib.load()  # fail loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the gm3d_b200 path has no CPU fallback "
                         "(use --impl reference for the CPU oracle)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wd = _watchdog(float(os.environ.get("GM3D_BENCH_WATCHDOG_S", "300")))
    if world > 1:
        # a mismatched / missing collective aborts within the timeout instead of spinning in NCCL kernels
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "1")
        os.environ.setdefault("TORCH_NCCL_ENABLE_MONITORING", "1")
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    B, N, G, k, ratio, desc = cfg
    K, W = args.steps, max(args.warmup, 3)
    cdict = config_dict(cfg)
    M = cdict["M"]
    ring = ring_size(B, N, G, k, M)

    # ---- collective mode: the per-step statistics all-reduce
    inbox = None
    mode = "none"
    if world > 1:
        mode = args.collective
        if mode in ("auto", "peer", "peer_sync", "peer_step"):
            from gm3d_b200.dist import PeerInbox
            try:
                inbox = PeerInbox(ring)
                mode = "peer" if mode == "auto" else mode
            except (RuntimeError, NotImplementedError) as e:  # raised on every rank together
                _stage(f"peer memory unavailable ({e}); falling back to NCCL")
                if args.collective != "auto":
                    raise
                mode = "nccl"
    collective = {"none": "none",
                  "peer": "per step over NVLink peer memory: every loss launch pushes {sum, sum_sq, count} to every rank's "
                          "inbox; one collect launch per graph of steps, beside the next graph's launches, sums them in "
                          "rank order (gm3d_step_reduce_t: defer + lagging collect)",
                  "peer_sync": "as peer, but the collect launch closes the same graph (one wait for the slowest rank per graph)",
                  "peer_step": "per step, inside the loss launch: {sum, sum_sq, count} pushed to every rank's inbox over "
                               "NVLink peer memory, waited for and summed in rank order by the same launch",
                  "nccl": "one NCCL all-reduce of the packed (steps, 4) statistics per graph of steps"}[mode]

    # ring of buffer sets larger than L2 so every step reads cold inputs and writes cold outputs
    steps, inputs = [], []
    for r in range(ring):
        s = GroupLossStep(B, N, G, k, ratio, device=dev, seed=1234, rand_offset=(rank * ring + r) * B * G)
        x, lp, _ = synthetic_batch(B, N, G, k, M, 1234 + 1000 * rank + r)
        s.xyz.copy_(torch.from_numpy(x)); s.loss_pred.copy_(torch.from_numpy(lp))
        pred = near_target_pred(s, 99 + 1000 * rank + r)
        steps.append(s)
        inputs.append((x, lp, pred) if r == 0 else None)

    # K timed steps = q replays of the whole ring captured as ONE graph + one graph of the first K % ring steps.
    # --no-overlap: one graph per step (kernel after kernel).
    overlap = not args.no_overlap
    chunks = {}

    def chunk(n):
        if n not in chunks:
            if overlap:
                chunks[n] = StepRing(steps[:n], reduce=mode, inbox=inbox).capture()
            else:
                rings = [StepRing([steps[i]], reduce=mode, inbox=_Slot(inbox, i) if inbox is not None else None).capture()
                         for i in range(n)]
                chunks[n] = rings
        return chunks[n]

    def run_chunk(n):
        c = chunk(n)
        if overlap:
            c.run()
        else:
            for s in c:
                s.run()

    def run_steps(n):
        for _ in range(n // ring):
            run_chunk(ring)
        if n % ring:
            run_chunk(n % ring)

    align = torch.zeros(1, device=dev)

    def device_align():
        # device-side start alignment: an all-reduce ENQUEUED on the stream completes on all ranks within a few us of
        # each other, so the ranks enter a timed region together without a host round trip
        if world > 1:
            dist.all_reduce(align)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    _stage("buffers ready; capturing")
    chunk(ring)
    _stage("ring captured")
    if K % ring:
        chunk(K % ring)
    _stage("warm-up")
    run_steps(max(W, ring))  # at least one pass over every buffer set
    run_steps(K)             # and one pass over exactly the graphs that are timed
    barrier()
    # ---- how many repetitions of the K-step region: identical on every rank (derived from an all-reduced estimate)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    device_align()
    e0.record()
    run_steps(K)
    e1.record()
    torch.cuda.synchronize()
    est = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    est_ms = max(est.item(), 1e-3)
    R = args.reps if args.reps > 0 else int(min(3000, max(25, -(-args.min_region_ms // est_ms))))
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.15)
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(R)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(R)]
    _stage(f"timed region: {R} repetitions of {K} steps")
    barrier()
    # Ranks that exchange nothing inside a repetition (collective none / peer: the sums of one replay are formed beside
    # the next) need no per-repetition alignment -- a repetition's duration on a rank does not depend on when the others
    # start, and the lagging collect keeps the ranks within one replay of each other.  Coupled modes re-align every time.
    coupled = mode in ("peer_sync", "peer_step", "nccl")
    device_align()
    for r in range(R):
        if coupled:
            device_align()
        ev0[r].record()
        run_steps(K)
        ev1[r].record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    times = torch.tensor([ev0[r].elapsed_time(ev1[r]) for r in range(R)], device=dev)
    rank_medians = [float(times.median().item()) / K]
    if world > 1:
        own = times.median().reshape(1) / K
        allm = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(allm, own)
        rank_medians = [float(t.item()) for t in allm]  # each rank's own median, before the max over ranks
        dist.all_reduce(times, op=dist.ReduceOp.MAX)  # per repetition: the slowest rank
    tl = np.sort(times.cpu().numpy())
    ms = float(np.median(tl))
    value = world * B * K / (ms * 1e-3)
    _stage("timed region done")

    # ---- checks on what the timed launches left behind
    parity = allreduce = None
    if rank == 0 and not args.no_checks:
        try:
            parity = oracle_check_step(steps[0], *inputs[0])
        except AssertionError as e:
            parity = f"FAILED: {e}"
    if world > 1 and not args.no_checks:
        # every rank's own [sum, sum_sq, count] of the ring's steps, gathered; head must be their sum over the ranks
        c = chunk(ring)
        rings = [c] if overlap else c
        for r_ in rings:  # one more pass over every slot, then drain the lagging sums
            r_.run()
        for r_ in rings:
            r_.flush()
        torch.cuda.synchronize()
        mine = torch.stack([s.stats[:3] for s in steps]).contiguous()
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        want = allv[0].clone()
        for q in range(1, world):
            want = want + allv[q]  # rank order, fp32: the order the peer reduction uses
        head = torch.cat([r_.head for r_ in rings])
        got = head[:, :3]
        if mode == "none":
            ok, ranks_ok = torch.equal(got, mine), True
        else:  # peer: summed in rank order => bit-identical; NCCL: its own (deterministic) order
            ok = torch.equal(got, want) if mode.startswith("peer") else torch.allclose(got, want, rtol=1e-6, atol=0)
            ranks_ok = bool((head[:, 3] == world).all())  # column 3 counts the ranks summed
        st_ok = inbox is None or int(inbox.status.item()) == 0
        flag = torch.tensor([1.0 if (ok and ranks_ok and st_ok) else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        allreduce = "ok" if flag.item() == 1.0 else "FAILED: head != sum over ranks of the per-rank statistics"

    # ---- per-kernel device times (CUDA events on the launching stream, same ring => cold L2)
    per_kernel = {}
    single = None
    if rank == 0:
        _stage("per-kernel times")
        L = steps[0].lib
        p = lambda t: t.data_ptr()  # noqa: E731
        g = 1.0 / (steps[0].P * k)
        launchers = {
            "fps": lambda s, st: L.gm3d_fps_f32(p(s.xyz), B, N, G, p(s.fps_idx), p(s.center), None, st),
            "knn_group": lambda s, st: _knn_group_only(L, s, st),
            "chamfer_fused": lambda s, st: L.gm3d_chamfer_fused_f32(
                p(s.pred), p(s.neighborhood), p(s.patch_index), s.P, k, k, g, g, p(s.dist1), p(s.dist2), p(s.idx1),
                p(s.idx2), p(s.per_patch), p(s.total), p(s.stats), 2, p(s.grad_pred), None, None, 0, p(s.cd_ws), st),
            "chamfer_fwd": lambda s, st: L.gm3d_chamfer_fwd_f32(p(s.pred), p(s.neighborhood), p(s.patch_index), s.P, k, k,
                                                            p(s.dist1), p(s.dist2), p(s.idx1), p(s.idx2), p(s.per_patch),
                                                            None, None, 2, None, st),
            "chamfer_bwd": lambda s, st: L.gm3d_chamfer_bwd_f32(p(s.pred), p(s.neighborhood), p(s.patch_index), p(s.idx1),
                                                            p(s.idx2), None, None, g, g, s.P, k, k, p(s.grad_pred), None, st),
            "hard_mask": lambda s, st: L.gm3d_hard_mask_f32(p(s.loss_pred), B, G, s.len_keep, s.len_loss, None, 1, 0,
                                                        p(s.mask), p(s.patch_index), 0, st),
        }
        if steps[0].group_per_cloud:
            launchers["group"] = lambda s, st: L.gm3d_cloud_step_f32(
                p(s.xyz), B, N, G, k, p(s.fps_idx), p(s.center), None, p(s.neighborhood), None, None, 0, 0, None, 0, 0,
                None, None, None, 0.0, 0.0, 2, None, None, None, None, None, None, None, None, 0, None, None, st)
            launchers["cloud_step"] = lambda s, st: L.gm3d_cloud_step_f32(
                p(s.xyz), B, N, G, k, p(s.fps_idx), p(s.center), None, p(s.neighborhood), None, p(s.loss_pred), s.len_keep,
                s.len_loss, None, s.seed, s.rand_offset, p(s.mask), p(s.patch_index), p(s.pred), g, g, 2, p(s.dist1),
                p(s.dist2), p(s.idx1), p(s.idx2), p(s.per_patch), p(s.total), p(s.stats), p(s.grad_pred), 0, None, p(s.cd_ws), st)
        for name, fn in launchers.items():
            # one graph holding `ring` launches of this kernel (one per buffer set => cold L2 every launch);
            # replayed so that host launch gaps do not pollute the per-launch time
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn_st = side.cuda_stream
                for s in steps:
                    fn(s, fn_st)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gk = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gk):
                cst = torch.cuda.current_stream().cuda_stream
                for s in steps:
                    fn(s, cst)
            gk.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            a.record()
            for _ in range(reps):
                gk.replay()
            b.record()
            torch.cuda.synchronize()
            per_kernel[name] = a.elapsed_time(b) * 1e3 / (reps * ring)  # us per launch, back-to-back in a graph

        # ---- the same work as ONE launch per step (gm3d_cloud_step_f32 with pred): reported, not the headline
        if steps[0].group_per_cloud:
            _stage("single-launch ring")
            one = []
            for r in range(ring):
                s = GroupLossStep(B, N, G, k, ratio, device=dev, seed=1234, rand_offset=r * B * G, path="single")
                s.in_arena.copy_(steps[r].in_arena)
                one.append(s)
            sr = StepRing(one).capture()
            for _ in range(3):
                sr.run()
            torch.cuda.synchronize()
            ts = []
            for _ in range(25):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                sr.run()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            t1 = float(np.median(ts)) / ring
            single = {"value": B / (t1 * 1e-3), "unit": "clouds/s", "ms_per_step": t1, "kernels_per_step": 1,
                      "note": "gm3d_cloud_step_f32 with pred: usable only when pred does not depend on this step's "
                              "grouping / mask; one GPU, this rank's share"}
            del sr, one

    # ---- the same step through the DROP-IN operator surface, eagerly, the way the reference's training loop calls it:
    # Group.forward -> generate_mask -> forward_loss -> backward (engine_pretrain_Classifier_SVM.py:108-118,176-184)
    dropin = None
    if rank == 0:
        _stage("drop-in sequence")
        from gm3d_b200 import loss as gl
        from gm3d_b200 import masking
        from gm3d_b200.group import Group
        grp = Group(G, k)
        s0 = steps[0]
        predv = s0.pred.reshape(B, M, k * 3)

        def dropin_step():
            nb, _ = grp(s0.xyz)
            mask = masking.generate_mask(s0.loss_pred, ratio, epoch=199, total_epoch=400, seed=1234, offset=0)
            pr = predv.detach().requires_grad_(True)
            out = gl.forward_loss_usual(pr, nb, mask)
            out["Chamfer_mean"].backward()
            return out["Chamfer_mean"], pr.grad
        for _ in range(5):
            dropin_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nd = 50
        for _ in range(nd):
            lossv, gradv = dropin_step()
        torch.cuda.synchronize()
        td = (time.perf_counter() - t0) / nd
        same = bool(torch.equal(gradv.reshape(s0.P, k, 3), s0.grad_pred)) and bool(torch.equal(lossv, s0.total.reshape(())))
        dropin = {"value": B / td, "unit": "clouds/s", "ms_per_step": td * 1e3,
                  "equals_step_outputs": same,
                  "note": "eager Python: Group -> generate_mask -> forward_loss_usual (one fused fwd+bwd launch behind "
                          "autograd) -> backward, allocations and host launch latency included; one GPU, this rank's share"}

    _stage("e2e")
    # ---- end-to-end: every step fed from pinned host memory, results read back
    # NGROUP graphs in flight on NGROUP streams, each = SUB x [H2D copy, the step, D2H copy]
    NGROUP, SUB = int(os.environ.get("GM3D_E2E_GROUPS", "6")), int(os.environ.get("GM3D_E2E_SUB", "4"))
    numa = _near_gpu(local)

    def build_groups(cloud_only):
        with numa:
            return _build_groups(cloud_only)

    def _build_groups(cloud_only):
        out = []
        for gi in range(NGROUP):
            sub = []
            for r in range(SUB):
                s = HostStagedStep(B, N, G, k, ratio, device=dev, seed=1234, rand_offset=(gi * SUB + r) * B * G, cloud_only=cloud_only)
                x, lp, _ = synthetic_batch(B, N, G, k, M, 4321 + 1000 * rank + gi * SUB + r)
                s.xyz.copy_(torch.from_numpy(x)); s.loss_pred.copy_(torch.from_numpy(lp))
                pred = near_target_pred(s, 4321 + 1000 * rank + gi * SUB + r)  # leaves s.pred / s.loss_pred on the device too
                s.h_xyz.copy_(torch.from_numpy(x)); s.h_pred.copy_(torch.from_numpy(pred)); s.h_loss_pred.copy_(torch.from_numpy(lp))
                sub.append(s)
            out.append(HostStagedGroup(sub).capture())
        torch.cuda.synchronize()
        return out

    Ke = -(-max(K, 240) // SUB) * SUB  # the host-fed pipeline needs a few hundred steps to reach steady state
    losses = []

    def e2e_steps(groups, n):
        # every step's inputs cross PCIe from pinned memory and its loss / per-patch matrix / mask come back;
        # the host reads the losses of a group before it re-launches that group
        busy = [False] * NGROUP
        for i in range(n // SUB):
            j = i % NGROUP
            if busy[j]:
                losses.extend(groups[j].losses())
            groups[j].launch()
            busy[j] = True
        for j in range(NGROUP):
            jj = (n // SUB + j) % NGROUP
            if busy[jj]:
                losses.extend(groups[jj].losses())

    def time_e2e(cloud_only):
        groups = build_groups(cloud_only)
        e2e_steps(groups, NGROUP * SUB * 2)
        barrier()
        losses.clear()
        t0 = time.perf_counter()
        e2e_steps(groups, Ke)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, groups[0].steps[0]

    dt, hs0 = time_e2e(False)
    n_read = len(losses)
    loss_last = losses[-1] if losses else None
    dt_cloud, hs1 = time_e2e(True)
    hs = [hs0]
    if world > 1:
        t = torch.tensor([dt, dt_cloud], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_cloud = t[0].item(), t[1].item()
    e2e = {"value": world * B * Ke / dt, "unit": "clouds/s", "h2d_bytes_per_step": hs[0].h2d_bytes,
           "d2h_bytes_per_step": hs[0].d2h_bytes, "ms_per_step": dt / Ke * 1e3, "steps": Ke,
           "pcie_gbs": (hs[0].h2d_bytes + hs[0].d2h_bytes) * Ke / dt / 1e9,
           "losses_read": n_read, "host_numa": numa.info,
           "cloud_only": {"value": world * B * Ke / dt_cloud, "unit": "clouds/s", "h2d_bytes_per_step": hs1.h2d_bytes,
                          "note": "only the point clouds cross PCIe; pred / loss_pred stay on the device, where the reference's "
                                  "decoder and loss predictor produce them"},
           "how": f"{NGROUP} graphs in flight on {NGROUP} streams, each {SUB} x [1 H2D copy from pinned memory, the step "
                  f"({steps[0].kernels_per_step} launches), 1 D2H copy]; every step's loss read on the host; wall clock, max over ranks"}

    _stage("e2e done")
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
        fp32_peak = 148 * 128 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12  # TFLOP/s, non-tensor FP32
        bpc = steps[0].bytes_per_cloud()
        evals = {"fps": (G - 1) * N, "knn_group": G * N, "group": (G - 1) * N + G * N, "chamfer_fused": M * k * k,
                 "chamfer_fwd": M * k * k, "cloud_step": (G - 1) * N + G * N + M * k * k}
        ncu = {}
        try:  # per-launch figures of the committed ncu --set full capture of the timed kernels (profiles/)
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                ncu = json.load(f).get(args.config, {})
        except (OSError, ValueError):
            pass
        us_step = ms / K * 1e3
        if steps[0].group_per_cloud:
            # the group launches chain back to back on their stream (the mask / loss launches run beside them on forked
            # streams), so over the timed region the average duration of a group launch on its stream IS the step time
            dom, dom_us = "group", us_step
        else:
            dom = max((n for n in per_kernel if n in bpc and n not in ("cloud_step", "group")), key=lambda n: per_kernel[n])
            dom_us = per_kernel[dom]
        hbm_gbs = bpc[dom] * B / (dom_us * 1e-6) / 1e9
        fp32_tf = 8 * evals.get(dom, 0) * B / (dom_us * 1e-6) / 1e12
        hbm_frac, fp32_frac = hbm_gbs / hbm_peak, fp32_tf / fp32_peak
        nk = ncu.get(dom) if isinstance(ncu.get(dom), dict) else {}
        by_fp32 = fp32_frac > hbm_frac  # the roof the kernel sits closest to (min(HBM, FP32))
        roofline = {"kernel": dom, "bound": "fp32" if by_fp32 else "hbm",
                    "achieved": fp32_tf if by_fp32 else hbm_gbs, "peak": fp32_peak if by_fp32 else hbm_peak,
                    "unit": "TFLOP/s" if by_fp32 else "GB/s", "frac": fp32_frac if by_fp32 else hbm_frac,
                    "traffic": nk.get("dram_bytes"), "peak_source": peak_src if not by_fp32 else "148 SM x 128 lanes x 2 x max SM clock (non-tensor FP32)",
                    "hbm_gbs": hbm_gbs, "hbm_peak": hbm_peak, "hbm_frac": hbm_frac,
                    "fp32_tflops": fp32_tf, "fp32_peak": fp32_peak, "fp32_frac": fp32_frac,
                    "issue_slot_frac": nk.get("issue_slot_frac"),  # of the kernel profiled alone (ncu serialises launches)
                    "us_per_launch": dom_us,
                    "algorithmic_bytes_per_launch": bpc[dom] * B, "algorithmic_flop_per_launch": 8 * evals.get(dom, 0) * B,
                    "limiter": "instruction issue + the dependent latency of the G-round sampling chain (one CTA per cloud); "
                               "neither roof binds -- see DESIGN.md 4.1"}
        detail = {}
        for n, us in per_kernel.items():
            d = {"us_per_launch": round(us, 3)}
            if n in bpc:
                d["hbm_gbs"] = round(bpc[n] * B / (us * 1e-6) / 1e9, 1)
                d["hbm_frac"] = round(d["hbm_gbs"] / hbm_peak, 4)
            if n in evals:
                d["fp32_tflops"] = round(8 * evals[n] * B / (us * 1e-6) / 1e12, 2)
                d["fp32_frac"] = round(d["fp32_tflops"] / fp32_peak, 4)
            if n == "fps":
                d["us_per_iteration"] = round(us / max(G - 1, 1), 4)
            detail[n] = d
        step_bytes = steps[0].step_bytes_per_cloud() * B
        step_flop = 8 * ((G - 1) * N + G * N + M * k * k) * B
        cpu_base, _ = time_cpu(cfg, budget_s=12.0) if not args.no_cpu_baseline else ({"value": None, "unit": "clouds/s", "cores": 0, "kind": "port", "sample": "skipped"}, 0)
        line = {"metric": METRIC, "value": value, "unit": "clouds/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": cdict,
                "rank_median_ms_per_step": rank_medians,
                "reps": R, "ms_per_step_min": float(tl[0]) / K, "ms_per_step_p90": float(tl[int(0.9 * (R - 1))]) / K,
                "timing": f"median over {R} repetitions of the {K}-step region (CUDA events per repetition, max over ranks per "
                          "repetition; barrier + device-side alignment before the first repetition" +
                          (", re-aligned before every repetition" if coupled else "") + ")",
                "run": {"path": steps[0].path, "kernels_per_step": steps[0].kernels_per_step, "cuda_graph": True,
                        "step_overlap": ((f"independent steps (mask -> group -> loss launches each) round-robin on forked streams "
                                          f"inside graphs of up to {ring} steps" if steps[0].group_per_cloud else
                                          f"independent steps round-robin on {os.environ.get('GM3D_RING_LANES', '4')} forked streams "
                                          f"inside graphs of up to {ring} steps") if overlap else "none"),
                        "collective": collective},
                "clocks": clk, "e2e": e2e, "gpu_launches": steps[0].kernels_per_step * K,
                "roofline": roofline, "roofline_detail": detail,
                "step_hbm": {"algorithmic_bytes_per_step": step_bytes, "gbs": step_bytes / (us_step * 1e-6) / 1e9,
                             "frac": step_bytes / (us_step * 1e-6) / 1e9 / hbm_peak},
                "step_fp32": {"flop_per_step": step_flop, "tflops": step_flop / (us_step * 1e-6) / 1e12,
                              "frac": step_flop / (us_step * 1e-6) / 1e12 / fp32_peak},
                "single_launch": single, "dropin": dropin,
                "cpu_baseline": cpu_base, "loss_check": loss_last, "parity_check": parity, "allreduce_check": allreduce}
        print(json.dumps(line), flush=True)
    failed = (parity is not None and parity != "ok") or (allreduce is not None and allreduce != "ok")
    wd.cancel()
    if world > 1:
        # Captured NCCL work keeps the communicator busy at teardown (destroy_process_group / interpreter exit
        # can hang on graphs that hold collectives): synchronise, flush and leave without running destructors.
        fl = torch.tensor([1.0 if failed else 0.0], device=dev)
        dist.all_reduce(fl, op=dist.ReduceOp.MAX)
        failed = fl.item() != 0.0
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1 if failed else 0)
    if failed:
        raise SystemExit(1)


def run_c3(args, cfg):
    """BASELINE config[2]: the three chained Point-M2AE grouping levels + masks + multi-scale Chamfer as ONE step."""
    import torch

    from gm3d_b200 import _lib
    from gm3d_b200.pipeline import HostStagedGroup, HostStagedM2AE, M2AEStep, StepRing

    _lib.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the gm3d_b200 path has no CPU fallback")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("bench.py --config c3 is a one-GPU configuration")
    wd = _watchdog(float(os.environ.get("GM3D_BENCH_WATCHDOG_S", "300")))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, N, Gs, ks, ratio, desc = cfg
    K, W = args.steps, max(args.warmup, 3)
    cdict = config_dict(cfg)
    ring = int(cdict["l2_policy"].split("ring of ")[1].split()[0])
    lv = m2ae_levels(cfg)

    def fill(s, seed, host=False):
        x, lps, _ = m2ae_inputs(cfg, seed)
        s.xyz.copy_(torch.from_numpy(x))
        preds = []
        for li, (l, lp) in enumerate(zip(s.levels, lps)):  # level by level: level l+1 groups level l's centres
            l.loss_pred.copy_(torch.from_numpy(lp))
            preds.append(near_target_pred(l, seed + 17 * li))
        if host:
            s.h_views[0][0].copy_(torch.from_numpy(x))
            for li in range(len(lv)):
                v = s.h_views[li]
                v[-2].copy_(torch.from_numpy(preds[li])); v[-1].copy_(torch.from_numpy(lps[li]))
        return x, lps, preds

    steps, first = [], None
    for r in range(ring):
        s = M2AEStep(B, N, Gs, ks, ratio, device=dev, seed=1234, rand_offset=r * B * Gs[0])
        inp = fill(s, 1234 + r)
        first = first or inp
        steps.append(s)
    chunks = {}

    def run_steps(n):
        for m in [ring] * (n // ring) + ([n % ring] if n % ring else []):
            if m not in chunks:
                chunks[m] = StepRing(steps[:m]).capture()
            chunks[m].run()

    run_steps(max(W, ring)); run_steps(K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run_steps(K); e1.record()
    torch.cuda.synchronize()
    R = args.reps if args.reps > 0 else int(min(3000, max(25, -(-args.min_region_ms // max(e0.elapsed_time(e1), 1e-3)))))
    clocks = ClockSampler(0)
    clocks.start()
    time.sleep(0.15)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(R)]
    torch.cuda.synchronize()
    for a, b in ev:
        a.record(); run_steps(K); b.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    tl = np.sort(np.array([a.elapsed_time(b) for a, b in ev]))
    ms = float(np.median(tl))
    # parity: every level of slot 0 against the oracle chain (level l+1 on the oracle's level-l centres)
    parity = None
    if not args.no_checks:
        from oracle import c_oracle as co
        try:
            cloud = first[0]
            for l, lp, pr in zip(steps[0].levels, first[1], first[2]):
                oracle_check_step(l, cloud, lp, pr)
                cloud = co.group(cloud, l.G, l.k)["center"]
            parity = "ok"
        except AssertionError as e:
            parity = f"FAILED: {e}"
    # e2e: host-fed steps, a few graphs in flight
    NG, SUB = 4, 2
    groups = []
    for gi in range(NG):
        sub = []
        for r in range(SUB):
            s = HostStagedM2AE(B, N, Gs, ks, ratio, device=dev, seed=1234, rand_offset=(gi * SUB + r) * B * Gs[0])
            fill(s, 4321 + gi * SUB + r, host=True)
            sub.append(s)
        groups.append(HostStagedGroup(sub).capture())
    Ke = -(-max(K, 64) // SUB) * SUB
    losses = []

    def e2e(n):
        busy = [False] * NG
        for i in range(n // SUB):
            j = i % NG
            if busy[j]:
                losses.extend(groups[j].losses())
            groups[j].launch(); busy[j] = True
        for j in range(NG):
            if busy[j]:
                losses.extend(groups[j].losses())
    e2e(NG * SUB * 2)
    torch.cuda.synchronize(); losses.clear()
    t0 = time.perf_counter(); e2e(Ke); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    hs = groups[0].steps[0]
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    fp32_peak = 148 * 128 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12
    us = ms / K * 1e3
    sb, sf = steps[0].step_bytes_per_cloud() * B, steps[0].flop_per_cloud() * B
    hbm_gbs, fp32_tf = sb / (us * 1e-6) / 1e9, sf / (us * 1e-6) / 1e12
    by_fp32 = fp32_tf / fp32_peak > hbm_gbs / hbm_peak
    cpu_base, _ = time_cpu(cfg, budget_s=12.0) if not args.no_cpu_baseline else ({"value": None, "unit": "clouds/s", "cores": 0, "kind": "port", "sample": "skipped"}, 0)
    line = {"metric": METRIC, "value": B * K / (ms * 1e-3), "unit": "clouds/s", "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cdict, "reps": R, "ms_per_step_min": float(tl[0]) / K,
            "ms_per_step_p90": float(tl[int(0.9 * (R - 1))]) / K,
            "run": {"path": "dataflow", "kernels_per_step": steps[0].kernels_per_step, "cuda_graph": True,
                    "step_overlap": f"independent steps round-robin on forked streams inside graphs of up to {ring} steps",
                    "collective": "none"},
            "clocks": clk,
            "e2e": {"value": B * Ke / dt, "unit": "clouds/s", "h2d_bytes_per_step": hs.h2d_bytes, "d2h_bytes_per_step": hs.d2h_bytes,
                    "ms_per_step": dt / Ke * 1e3, "steps": Ke, "losses_read": len(losses)},
            "gpu_launches": steps[0].kernels_per_step * K,
            "roofline": {"kernel": f"m2ae step ({steps[0].kernels_per_step} launches)", "bound": "fp32" if by_fp32 else "hbm",
                         "achieved": fp32_tf if by_fp32 else hbm_gbs, "peak": fp32_peak if by_fp32 else hbm_peak,
                         "unit": "TFLOP/s" if by_fp32 else "GB/s", "frac": (fp32_tf / fp32_peak) if by_fp32 else hbm_gbs / hbm_peak,
                         "traffic": None, "peak_source": peak_src, "hbm_gbs": hbm_gbs, "hbm_frac": hbm_gbs / hbm_peak,
                         "fp32_tflops": fp32_tf, "fp32_frac": fp32_tf / fp32_peak, "us_per_launch": us,
                         "algorithmic_bytes_per_launch": sb, "algorithmic_flop_per_launch": sf},
            "cpu_baseline": cpu_base, "loss_check": losses[-1] if losses else None, "parity_check": parity}
    print(json.dumps(line), flush=True)
    wd.cancel()
    if parity is not None and parity != "ok":
        raise SystemExit(1)


class _Slot:
    """View of a PeerInbox that maps a one-step ring's slot 0 to slot `i` of the real inbox (--no-overlap)."""

    def __init__(self, inbox, i):
        self.inbox, self.i = inbox, i

    def step_reduce(self, slot, head_ptr, **kw):
        return self.inbox.step_reduce(self.i + slot, head_ptr, **kw)


def _knn_group_only(L, s, st):
    """The kNN + gather + normalise kernel in its Group configuration (neighbourhood written, no int64 idx)."""
    return L.gm3d_knn_group_f32(s.xyz.data_ptr(), s.center.data_ptr(), s.B, s.N, s.G, s.k, None,
                                s.neighborhood.data_ptr(), None, st)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="one graph per step, no programmatic dependent launch")
    ap.add_argument("--reps", type=int, default=0, help="repetitions of the K-step timed region (0: enough for --min-region-ms, at least 25)")
    ap.add_argument("--min-region-ms", type=float, default=600.0, help="total timed time wanted (clock sampling needs a few hundred ms)")
    ap.add_argument("--collective", default="auto", choices=["auto", "peer", "peer_sync", "peer_step", "nccl", "none"],
                    help="N > 1: per-step peer-memory all-reduce inside the loss launch (auto: if peer memory maps), or one NCCL all-reduce per graph")
    ap.add_argument("--no-checks", action="store_true", help="skip the oracle / all-reduce checks after the timed region")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    elif isinstance(cfg[2], tuple):
        run_c3(args, cfg)
    else:
        run_native(args, cfg)


if __name__ == "__main__":
    main()
