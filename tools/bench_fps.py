"""Stand-alone FPS kernel: us per launch and per round for the BASELINE shapes."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gm3d_b200 import ops
dev = torch.device("cuda", 0)
for B, N, G in ((128, 1024, 64), (32, 2048, 128), (128, 2048, 512), (128, 8192, 512), (32, 2048, 2048)):
    x = torch.randn(B, N, 3, device=dev)
    for _ in range(3): ops.furthest_point_sample(x, G)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): ops.furthest_point_sample(x, G)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / 20
    print(json.dumps({"B": B, "N": N, "G": G, "us": round(us, 1), "us_per_round": round(us / (G - 1), 4)}))
