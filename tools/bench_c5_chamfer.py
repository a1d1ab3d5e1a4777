"""Stand-alone Chamfer fwd+bwd launch at the C5 shard and C2 shapes: us per launch over a ring larger than L2."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gm3d_b200 import _lib  # noqa: E402

L = _lib.load()
dev = torch.device("cuda", 0)
for name, P, k, pool in (("c5 shard", 128 * 308, 32, 128 * 512), ("c2", 128 * 39, 32, 128 * 64), ("c3 level 0", 128 * 410, 16, 128 * 512)):
    ring = max(4, int(2 * 126e6 // (P * k * 3 * 4 * 2 + pool * k * 12)) + 1)
    rng = np.random.default_rng(1)
    sets = []
    for r in range(ring):
        a = torch.from_numpy((rng.standard_normal((P, k, 3)) * 0.08).astype(np.float32)).to(dev)
        b = torch.from_numpy((rng.standard_normal((pool, k, 3)) * 0.08).astype(np.float32)).to(dev)
        idx = torch.from_numpy(np.sort(rng.permutation(pool)[:P]).astype(np.int32)).to(dev)
        out = [torch.empty((P, k), dtype=torch.float32, device=dev) for _ in range(2)] + \
              [torch.empty((P, k), dtype=torch.int32, device=dev) for _ in range(2)]
        pp, tot, st = torch.empty(P, device=dev), torch.empty(1, device=dev), torch.empty(8, device=dev)
        g = torch.empty((P, k, 3), device=dev)
        ws = torch.zeros(L.gm3d_workspace_bytes(_lib.OP_CHAMFER_FWD, P, k, k, 0), dtype=torch.uint8, device=dev)
        sets.append((a, b, idx, out, pp, tot, st, g, ws))
    p = lambda t: t.data_ptr()  # noqa: E731

    def launch(s, stream):
        a, b, idx, out, pp, tot, st, g, ws = s
        rc = L.gm3d_chamfer_fused_f32(p(a), p(b), p(idx), P, k, k, 1.0 / (P * k), 1.0 / (P * k), p(out[0]), p(out[1]), p(out[2]),
                                      p(out[3]), p(pp), p(tot), p(st), 2, p(g), None, None, 0, p(ws), stream)
        assert rc == 0, rc
    for s in sets:
        launch(s, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    gk = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gk):
        for s in sets:
            launch(s, torch.cuda.current_stream().cuda_stream)
    gk.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gk.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (10 * ring)
    byts = P * (24 * k + 16 * k + 4 + 12 * k)
    print(f"chamfer fwd+bwd {name}: P={P} k={k}: {us:.2f} us per launch, {byts / us / 1e3:.0f} GB/s algorithmic = {byts / us / 1e3 / 6549:.3f} of measured HBM")
