"""Device times of the SURVEY 8(f) rows next to the hot path, with the NumPy oracle timed beside them.

    python tools/bench_next.py   -> one JSON line per operator
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gm3d_b200 import ops  # noqa: E402
from gm3d_b200.pointnet2_utils import fps_subsample  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import np_oracle as no  # noqa: E402

dev = torch.device("cuda", 0)
HBM = 6549.1
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except OSError:
    pass


def gpu_us(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps // 10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def cpu_us(fn, budget=3.0):
    fn()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget:
        fn()
        n += 1
    return (time.perf_counter() - t0) / n * 1e6


rng = np.random.default_rng(0)
# forward_learning_loss at the pre-training shape: (128, 39) per-patch matrix
p = rng.standard_normal((128, 39)).astype(np.float32)
t = (rng.random((128, 39)) * 0.05).astype(np.float32)
dp, dt_ = torch.from_numpy(p).to(dev), torch.from_numpy(t).to(dev)
lib_loss = lambda: ops.learning_loss(dp, dt_, True)  # noqa: E731
us = gpu_us(lib_loss)
print(json.dumps({"op": "forward_learning_loss(relative) fwd+bwd", "shape": [128, 39], "gpu_us": round(us, 2),
                  "pair_terms_per_s": round(128 * 39 * 38 / us * 1e6), "cpu_oracle_us": round(cpu_us(lambda: no.learning_loss(p, t, True)), 1)}))

# PointcloudScaleAndTranslate on the pre-training batch (128, 1024, 3): read + write 12 B per point
x = rng.standard_normal((128, 1024, 3)).astype(np.float32)
ss = rng.uniform(0.7, 1.4, (128, 6)).astype(np.float32)
dx, dss = torch.from_numpy(x).to(dev), torch.from_numpy(ss).to(dev)
us = gpu_us(lambda: ops.scale_translate_(dx, dss))
print(json.dumps({"op": "PointcloudScaleAndTranslate", "shape": [128, 1024, 3], "gpu_us": round(us, 2),
                  "hbm_gbs": round(2 * x.nbytes / us / 1e3, 1), "hbm_frac": round(2 * x.nbytes / us / 1e3 / HBM, 4),
                  "note": "1.5 MB working set: L2-resident, launch-latency bound", "cpu_oracle_us": round(cpu_us(lambda: no.scale_translate(x, ss)), 1)}))
big = torch.randn(2048, 8192, 3, device=dev)
bss = torch.rand(2048, 6, device=dev)
us = gpu_us(lambda: ops.scale_translate_(big, bss), reps=50)
print(json.dumps({"op": "PointcloudScaleAndTranslate", "shape": [2048, 8192, 3], "gpu_us": round(us, 2),
                  "hbm_gbs": round(2 * big.numel() * 4 / us / 1e3, 1), "hbm_frac": round(2 * big.numel() * 4 / us / 1e3 / HBM, 4)}))

# fine-tune sub-sampling (engine_finetune.py:118-134): (32, 2048, 3) -> FPS 2048 -> 1024 random columns
pts = rng.standard_normal((32, 2048, 3)).astype(np.float32)
dpts = torch.from_numpy(pts).to(dev)
ch = rng.choice(2048, 1024, False)
dch = torch.from_numpy(ch.astype(np.int64)).to(dev)


def sub():
    idx = ops.furthest_point_sample(dpts, 2048)
    return ops.gather_points(dpts, idx, dch, validate=False)


us = gpu_us(sub, reps=20)
print(json.dumps({"op": "fine-tune sub-sampling (FPS 2048 of 2048 + 1024-column gather)", "shape": [32, 2048, 3], "gpu_us": round(us, 1),
                  "clouds_per_s": round(32 / us * 1e6), "cpu_oracle_us": round(cpu_us(lambda: no.gather_points(pts, co.fps(pts, 2048), ch), budget=5.0), 1)}))

# ---- Encoder (mini-PointNet) forward: tcgen05 kernel vs the same module run by PyTorch (library GEMMs)
from gm3d_b200.encoder import EncoderB200  # noqa: E402

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_next.npz"))
sd = {}
for kname in g.files:
    if kname.startswith("enc_w_"):
        rest = kname[len("enc_w_"):]
        for pre in ("first_conv", "second_conv"):
            if rest.startswith(pre + "_"):
                i, nm = rest[len(pre) + 1:].split("_", 1)
                sd[f"{pre}.{i}.{nm}"] = torch.from_numpy(g[kname])
enc = EncoderB200.from_state_dict(sd).to(dev)
import torch.nn as nn  # noqa: E402


class TorchEncoder(nn.Module):  # the reference module's structure (models/Point_MAE.py:16-47), library kernels
    def __init__(self):
        super().__init__()
        self.first_conv = nn.Sequential(nn.Conv1d(3, 128, 1), nn.BatchNorm1d(128), nn.ReLU(inplace=True), nn.Conv1d(128, 256, 1))
        self.second_conv = nn.Sequential(nn.Conv1d(512, 512, 1), nn.BatchNorm1d(512), nn.ReLU(inplace=True), nn.Conv1d(512, 384, 1))

    def forward(self, pg):
        bs, gg, n, _ = pg.shape
        f = self.first_conv(pg.reshape(bs * gg, n, 3).transpose(2, 1))
        fg = torch.max(f, dim=2, keepdim=True)[0]
        f = self.second_conv(torch.cat([fg.expand(-1, -1, n), f], dim=1))
        return torch.max(f, dim=2)[0].reshape(bs, gg, 384)


ref = TorchEncoder().to(dev).eval()
ref.load_state_dict(sd)
for B_, G_ in ((128, 64), (32, 128)):
    nb = (torch.randn(B_, G_, 32, 3, device=dev) * 0.08)
    P_ = B_ * G_
    flop = 2.0 * 32 * (3 * 128 + 128 * 256 + 512 * 512 + 512 * 384) * P_
    us = gpu_us(lambda: enc(nb), reps=20)
    with torch.no_grad():
        us32 = gpu_us(lambda: ref(nb), reps=20)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            us16 = gpu_us(lambda: ref(nb), reps=20)
    print(json.dumps({"op": "Encoder forward (eval)", "patches": P_, "gpu_us": round(us, 1), "tflops": round(flop / us / 1e6, 1),
                      "frac_of_measured_bf16_peak": round(flop / us / 1e6 / 1668.9, 4), "patches_per_s": round(P_ / us * 1e6),
                      "torch_fp32_us": round(us32, 1), "torch_bf16_autocast_us": round(us16, 1)}))
