"""Launch each kernel of the path a few times on one configuration (for `ncu -k regex:...`).

    python tools/prof_kernels.py --config c2 [--reps 3] [--only knn_group]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, synthetic_batch  # noqa: E402
from gm3d_b200 import _lib  # noqa: E402
from gm3d_b200.pipeline import GroupLossStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--only", default="")
a = ap.parse_args()
B, N, G, k, ratio, _ = CONFIGS[a.config]
dev = torch.device("cuda", 0)
s = GroupLossStep(B, N, G, k, ratio, device=dev, seed=1, fused=False)
x, lp, pred = synthetic_batch(B, N, G, k, s.M, 1234)
s.xyz.copy_(torch.from_numpy(x)); s.loss_pred.copy_(torch.from_numpy(lp)); s.pred.copy_(torch.from_numpy(pred))
L = _lib.load()
st = torch.cuda.current_stream().cuda_stream
p = lambda t: t.data_ptr()  # noqa: E731
g = 1.0 / (s.P * k)
s.enqueue()
torch.cuda.synchronize()
kern = {
    "fps": lambda: L.gm3d_fps_f32(p(s.xyz), B, N, G, p(s.fps_idx), p(s.center), None, st),
    "knn_group": lambda: L.gm3d_knn_group_f32(p(s.xyz), p(s.center), B, N, G, k, None, p(s.neighborhood), None, st),
    "hard_mask": lambda: L.gm3d_hard_mask_f32(p(s.loss_pred), B, G, s.len_keep, s.len_loss, None, 1, 0, p(s.mask), p(s.patch_index), 0, st),
    "chamfer_fused": lambda: L.gm3d_chamfer_fused_f32(p(s.pred), p(s.neighborhood), p(s.patch_index), s.P, k, k, g, g, p(s.dist1),
                                                      p(s.dist2), p(s.idx1), p(s.idx2), p(s.per_patch), p(s.total), p(s.stats), 2,
                                                      p(s.grad_pred), None, None, 0, p(s.cd_ws), st),
}
if N <= 2048:
    kern["cloud_step"] = lambda: L.gm3d_cloud_step_f32(
        p(s.xyz), B, N, G, k, p(s.fps_idx), p(s.center), None, p(s.neighborhood), None, p(s.loss_pred), s.len_keep,
        s.len_loss, None, 1, 0, p(s.mask), p(s.patch_index), p(s.pred), g, g, 2, p(s.dist1), p(s.dist2), p(s.idx1),
        p(s.idx2), p(s.per_patch), p(s.total), p(s.stats), p(s.grad_pred), 0, None, p(s.cd_ws), st)
    kern["group"] = lambda: L.gm3d_cloud_step_f32(
        p(s.xyz), B, N, G, k, p(s.fps_idx), p(s.center), None, p(s.neighborhood), None, None, 0, 0, None, 0, 0,
        None, None, None, 0.0, 0.0, 2, None, None, None, None, None, None, None, None, 0, None, None, st)
for name, fn in kern.items():
    if a.only and a.only != name:
        continue
    for _ in range(a.reps):
        rc = fn()
        assert rc == 0, (name, rc)
    torch.cuda.synchronize()
print("ok", a.config, s.total.item())
