"""world_size-2 gloo tests (CPU) of the only collective the path needs and of the batch sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gm3d_b200 import dist as gd
        out = {}
        # misc.all_reduce_mean contract (util/misc.py:345-353): python float in, cross-rank mean out
        out["mean"] = gd.all_reduce_mean(float(rank + 1))
        # the step's statistics vector: [sum, sum_sq, count, min, max, ...]; per-rank shards of one loss vector
        full = np.random.default_rng(0).random(10).astype(np.float32)
        shard = full[rank::world]
        stats = torch.tensor([shard.sum(), (shard ** 2).sum(), len(shard), shard.min(), shard.max(), 0, 0, 0],
                             dtype=torch.float32)
        gd.all_reduce_stats(stats)
        out["stats"] = stats.tolist()
        v = gd.all_reduce_scalars([torch.tensor(float(rank)), torch.tensor(10.0 * rank)])
        out["scalars"] = v.tolist()
        # contiguous batch shard B/W per rank covers the batch exactly once
        B = 128
        lo, hi = rank * B // world, (rank + 1) * B // world
        cover = torch.zeros(B)
        cover[lo:hi] = 1
        dist.all_reduce(cover)
        out["cover"] = bool((cover == 1).all())
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = np.random.default_rng(0).random(10).astype(np.float32)
    for r in range(world):
        o = res[r]
        assert o["mean"] == pytest.approx(1.5)
        assert o["stats"][:5] == pytest.approx([full.sum(), (full ** 2).sum(), 10, full.min(), full.max()], rel=1e-6)
        assert o["scalars"] == pytest.approx([0.5, 5.0])
        assert o["cover"]


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (CPU oracle arm) emits one JSON line with the required keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "c1",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "clouds/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["higher_is_better"] is True and line["gpu_launches"] == 0
