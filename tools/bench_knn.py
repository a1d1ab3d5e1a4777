"""Stand-alone kNN + grouping kernel: us per launch for the BASELINE shapes with N > 1024 (queries = FPS centres)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gm3d_b200 import ops
dev = torch.device("cuda", 0)
for B, N, G, k in ((128, 8192, 512, 32), (128, 2048, 512, 16), (32, 2048, 128, 32), (128, 4096, 256, 32)):
    x = torch.randn(B, N, 3, device=dev)
    idx = ops.furthest_point_sample(x, G)
    c = torch.gather(x, 1, idx.long()[..., None].expand(-1, -1, 3)).contiguous()
    for _ in range(3): ops.knn(x, c, k, want_dist=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): ops.knn(x, c, k, want_dist=False)
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / 10
    print(json.dumps({"B": B, "N": N, "G": G, "k": k, "us": round(us, 1), "Gpairs_per_s": round(B * G * N / us * 1e-3, 1)}))
