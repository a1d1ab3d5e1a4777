"""Replay a ring of dataflow steps UN-captured (plain launches in stream order) so that ncu can attach to the kernels
bench.py times -- the 12-warp / 56-register per-cloud kernel, the warp-per-patch Chamfer kernel, the warp-per-row mask
kernel -- over a ring larger than L2 (write-back of one launch is then attributed to the following ones; summed over
>= ring launches it is complete).
    ncu --set full -k regex:cloud_step -s 24 -c 24 -o rep python tools/prof_ring.py --config c2 --rings 2"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from gm3d_b200.pipeline import GroupLossStep, StepRing  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--rings", type=int, default=2)
ap.add_argument("--path", default="dataflow")
a = ap.parse_args()
cfg = bench.CONFIGS[a.config]
B, N, G, k, ratio, _ = cfg
dev = torch.device("cuda", 0)
M = bench.config_dict(cfg)["M"]
ring = bench.ring_size(B, N, G, k, M)
steps = []
for r in range(ring):
    s = GroupLossStep(B, N, G, k, ratio, device=dev, seed=1234, rand_offset=r * B * G, path=a.path)
    x, lp, _ = bench.synthetic_batch(B, N, G, k, M, 1234 + r)
    s.xyz.copy_(torch.from_numpy(x)); s.loss_pred.copy_(torch.from_numpy(lp))
    pred = bench.near_target_pred(s, 99 + r)  # target + 0.02 * noise
    steps.append(s)
sr = StepRing(steps)
for _ in range(a.rings):
    sr.enqueue()  # the same launches, flags and streams the captured graph holds
    torch.cuda.synchronize()
print("ok", a.config, ring, "steps per ring,", steps[0].kernels_per_step, "launches per step, loss", steps[0].total.item())
