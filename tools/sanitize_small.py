"""Small invocations of every kernel family for `compute-sanitizer --tool memcheck` (one tool per gpurun call)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import synthetic_clouds  # noqa: E402
from gm3d_b200 import ops  # noqa: E402
from gm3d_b200.encoder import EncoderB200  # noqa: E402
from gm3d_b200.pipeline import GroupLossStep, StepRing  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
for (B, N, G, k, fused) in ((4, 1024, 64, 32, True), (3, 777, 50, 16, True), (2, 2048, 128, 32, True), (2, 4096, 64, 32, False)):
    steps = []
    for r in range(2):
        s = GroupLossStep(B, N, G, k, 0.6, device=dev, seed=1, fused=fused)
        s.xyz.copy_(torch.from_numpy(synthetic_clouds(B, N, 5 + r)))
        s.loss_pred.copy_(torch.from_numpy(rng.standard_normal((B, G)).astype(np.float32)))
        s.pred.copy_(torch.from_numpy((rng.standard_normal((s.P, k, 3)) * 0.08).astype(np.float32)))
        steps.append(s)
    StepRing(steps).enqueue()
    torch.cuda.synchronize()
    print("step ok", B, N, G, k, fused, float(steps[0].total))
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_next.npz"))
sd = {}
for kk in g.files:
    if kk.startswith("enc_w_"):
        rest = kk[len("enc_w_"):]
        for pre in ("first_conv", "second_conv"):
            if rest.startswith(pre + "_"):
                i, nm = rest[len(pre) + 1:].split("_", 1)
                sd[f"{pre}.{i}.{nm}"] = torch.from_numpy(g[kk])
enc = EncoderB200.from_state_dict(sd).to(dev)
out = enc(torch.randn(2, 7, 32, 3, device=dev) * 0.1)
torch.cuda.synchronize()
print("encoder ok", enc.last_status.item(), float(out.abs().mean()))
p = torch.randn(8, 39, device=dev)
t = torch.rand(8, 39, device=dev)
print("ll", float(ops.learning_loss(p, t, True)[0]), float(ops.learning_loss(p, t, False)[0]))
x = torch.randn(3, 100, 3, device=dev)
ops.scale_translate_(x, torch.rand(3, 6, device=dev))
idx = ops.furthest_point_sample(x, 40)
print("gather", ops.gather_points(x, idx, torch.arange(0, 40, 2, device=dev)).shape)
torch.cuda.synchronize()
