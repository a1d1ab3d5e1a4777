"""Device-resident us/step of one configuration and path, for A/B runs of tuning builds (GM3D_TUNING_ENV).
    python tools/quick_step.py [--config c2] [--path dataflow|single] [--reps 30]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gm3d_b200.pipeline import GroupLossStep, StepRing  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--path", default="dataflow")
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--check", action="store_true")
ap.add_argument("--parts", default="gmc", help="dataflow path: which launches to keep (g group, m mask, c chamfer)")
a = ap.parse_args()
B, N, G, k, ratio, _ = bench.CONFIGS[a.config]
dev = torch.device("cuda", 0)
M = bench.config_dict(bench.CONFIGS[a.config])["M"]
ring = bench.ring_size(B, N, G, k, M)
steps, inp = [], None
for r in range(ring):
    s = GroupLossStep(B, N, G, k, ratio, device=dev, seed=1234, rand_offset=r * B * G, path=a.path)
    x, lp, _ = bench.synthetic_batch(B, N, G, k, M, 1234 + r)
    s.xyz.copy_(torch.from_numpy(x)); s.loss_pred.copy_(torch.from_numpy(lp))
    pred = bench.near_target_pred(s, 99 + r)  # target + 0.02 * noise
    steps.append(s)
    if r == 0:
        inp = (x, lp, pred)
if a.parts != "gmc":  # decomposition runs: drop launches (results are then stale / wrong -- timing only)
    for s in steps:
        if "g" not in a.parts:
            s.enqueue_group = lambda *a_, **k_: None
        if "m" not in a.parts:
            s.enqueue_mask = lambda *a_, **k_: None
        if "c" not in a.parts:
            s.enqueue_loss = lambda *a_, **k_: None
sr = StepRing(steps).capture()
for _ in range(3):
    sr.run()
torch.cuda.synchronize()
ts = []
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sr.run(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3 / ring)
chk = bench.oracle_check_step(steps[0], *inp) if a.check else "-"
env = {k_: v for k_, v in os.environ.items() if k_.startswith("GM3D_")}
print(f"{a.config} {a.path:8s} {a.parts:3s} {np.median(ts):7.2f} us/step (min {min(ts):.2f})  {B / np.median(ts):.3f} M clouds/s  check={chk}  {env}")
