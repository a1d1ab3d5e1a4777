"""Per-CTA timeline of the fused kernel (clock64 stamps written when GM3D_CS_TRACE is set).

Needs the tuning aids compiled in:  GM3D_NVCC_FLAGS=-DGM3D_CS_DEBUG python -m gm3d_b200.build  (rebuild without
the flag afterwards: the production kernel leaves them out).

    python tools/trace_cs.py [--config c2]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
a = ap.parse_args()
from bench import CONFIGS, synthetic_batch  # noqa: E402

B, N, G, k, ratio, _ = CONFIGS[a.config]
dev = torch.device("cuda", 0)
trace = torch.zeros((B, 64), dtype=torch.int64, device=dev)
os.environ["GM3D_CS_TRACE"] = hex(trace.data_ptr())
from gm3d_b200.pipeline import GroupLossStep, StepRing  # noqa: E402

steps = []
for r in range(8):
    s = GroupLossStep(B, N, G, k, ratio, device=dev, seed=1, path="single")  # the trace lives in the one-launch kernel (-DGM3D_CS_DEBUG build)
    x, lp, pred = synthetic_batch(B, N, G, k, s.M, 1234 + r)
    s.xyz.copy_(torch.from_numpy(x)); s.loss_pred.copy_(torch.from_numpy(lp)); s.pred.copy_(torch.from_numpy(pred))
    steps.append(s)
ring = StepRing(steps).capture()
for _ in range(3):
    ring.run()
torch.cuda.synchronize()
for _ in range(20):
    ring.run()  # keep clocks up
trace.zero_()
steps[0].run()  # then one launch alone on the GPU
torch.cuda.synchronize()
t = trace.cpu().numpy().astype(np.int64)
ns = t[:, 61] - t[:, 60]
cyc = t[:, 4:28].max(1) - t[:, 0]
mhz = float(np.median(cyc / np.maximum(ns, 1)) * 1e3)
print(f"CTA life: {np.median(ns) / 1e3:.2f} us by %globaltimer, {np.median(cyc):.0f} cycles => SM clock {mhz:.0f} MHz; "
      f"kernel span {(t[:, 61].max() - t[:, 60].min()) / 1e3:.2f} us")
t0 = t[:, 0:1]
pro, fps, end = (t[:, 1] - t[:, 0]) / mhz, (t[:, 2] - t[:, 0]) / mhz, (t[:, 4:28].max(1) - t[:, 0]) / mhz
print(f"isolated launch, us from CTA start (median over {B} CTAs): prologue {np.median(pro):.2f}  fps done {np.median(fps):.2f}  last warp done {np.median(end):.2f}")
wd = (t[:, 4:28] - t0) / mhz
print("warp finish (us, median over CTAs):", np.round(np.median(wd, 0), 1))
print("worker wait for centres (us, median):", np.round(np.median(t[:, 32:56], 0) / mhz, 1))
