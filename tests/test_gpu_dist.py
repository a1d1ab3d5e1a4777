"""-m gpu tests of the N > 1 path on real GPUs (skipped on a one-GPU box): the per-step statistics all-reduce over
peer memory inside the loss launch (dist.PeerInbox + gm3d_step_reduce_t), its NCCL alternative, and the
misc.all_reduce_mean drop-in (/root/reference/Point-MAE_SA3D/util/misc.py:345-353) over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import synthetic_clouds

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import datetime

    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=90))
    out = {}
    try:
        from gm3d_b200 import dist as gd
        from gm3d_b200.pipeline import GroupLossStep, StepRing
        out["mean"] = gd.all_reduce_mean(float(rank + 1))
        B, N, G, k, n = 16, 1024, 64, 32, 5
        rng = np.random.default_rng(100 + rank)
        steps = []
        for r in range(n):
            s = GroupLossStep(B, N, G, k, 0.6, device=dev, seed=9, rand_offset=(rank * n + r) * B * G)
            s.xyz.copy_(torch.from_numpy(synthetic_clouds(B, N, 1000 * rank + r)).to(dev))
            s.loss_pred.copy_(torch.from_numpy(rng.standard_normal((B, G)).astype(np.float32)).to(dev))
            s.pred.copy_(torch.from_numpy((rng.standard_normal((s.P, k, 3)) * 0.08).astype(np.float32)).to(dev))
            steps.append(s)
        inbox = gd.PeerInbox(n)
        for mode in ("peer", "peer_sync", "peer_step", "nccl"):
            ring = StepRing(steps, reduce=mode, inbox=inbox if mode != "nccl" else None).capture()
            for _ in range(4):  # several replays: launch-counter wrap of the inbox depth, slot reuse
                ring.run()
            if mode == "peer":  # lagging sums: head holds replay 3 until the drain
                torch.cuda.synchronize()
                out["peer_lag_ranks"] = ring.head[:, 3].tolist()
                ring.flush()
            torch.cuda.synchronize()
            mine = torch.stack([s.stats[:3] for s in steps]).contiguous()
            allv = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allv, mine)
            want = allv[0].clone()
            for r in range(1, world):
                want = want + allv[r]
            got = ring.head[:, :3]
            out[mode + "_exact"] = bool(torch.equal(got, want))
            out[mode + "_close"] = bool(torch.allclose(got, want, rtol=1e-6, atol=0))
            out[mode + "_ranks"] = ring.head[:, 3].tolist()
            # every rank holds the same reduced values
            same = [torch.empty_like(ring.head) for _ in range(world)]
            dist.all_gather(same, ring.head.contiguous())
            out[mode + "_same"] = all(torch.equal(same[0], t) for t in same)
        out["status"] = int(inbox.status.item())
        out["epoch"] = inbox.epoch.tolist()
        inbox.close()
        q.put((rank, out))
        q.close()
        q.join_thread()
    finally:
        # captured graphs hold NCCL work: destroy_process_group can hang on them (as in bench.py) -- leave directly
        torch.cuda.synchronize()
        os._exit(0)


def test_peer_and_nccl_step_reduce_two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in range(world))  # a hung worker fails the test here, not the box
    for p in procs:
        p.join(timeout=60)
        if p.exitcode is None:
            p.kill()
        assert p.exitcode == 0
    for r in range(world):
        o = res[r]
        assert o["mean"] == pytest.approx(1.5)
        for m in ("peer", "peer_sync", "peer_step"):  # summed in rank order: bit-identical on every rank
            assert o[m + "_exact"] and o[m + "_same"] and o[m + "_ranks"] == [2.0] * 5
        assert o["nccl_close"] and o["nccl_same"] and o["nccl_ranks"] == [2.0] * 5
        assert o["status"] == 0 and o["epoch"] == [15] * 5  # three peer modes x (1 warm-up enqueue + 4 replays) per slot
