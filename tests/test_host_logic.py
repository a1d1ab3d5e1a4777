"""CPU tests of the host-side mirror of the reference interface (no device work)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as no


def test_mask_lengths_follow_reference_arithmetic():
    from gm3d_b200.masking import mask_lengths
    # GM3D: L=64, ratio 0.6 -> len_keep = int(64*0.4) = 25 -> 39 masked (..._feature_besed.py:1064-1065)
    assert mask_lengths(64, 0.6, 0, 400)[0] == 25
    for epoch in (0, 9, 199, 399):
        for cap in (0.8, 0.5):
            assert mask_lengths(64, 0.6, epoch, 400, True, None, cap) == no.mask_lengths(64, 0.6, epoch, 400, True, None, cap)
    assert mask_lengths(64, 0.6, 199, 400) == (25, 15)        # keep_ratio 0.4 -> int(39*0.4)
    assert mask_lengths(64, 0.6, 399, 400) == (25, 31)
    assert mask_lengths(64, 0.6, 0, 400)[1] == 0               # first epochs: pure random branch
    assert mask_lengths(64, 0.6, 350, 400, True, True) == (25, 19)   # after_200_epoch variant, capped at 0.5
    assert mask_lengths(64, 0.6, 5, 400, False) == (25, 19)    # guide=False -> keep_ratio 0.5
    assert mask_lengths(512, 0.6, 199, 400)[0] == int(512 * (1 - 0.6))


def test_ops_reject_cpu_tensors_and_bad_dtypes():
    from gm3d_b200 import ops
    x = torch.zeros(2, 16, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.furthest_point_sample(x, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.knn(x, x[:, :2], 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.chamfer_forward(x, x)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.group(x, 4, 2)
    with pytest.raises(RuntimeError):
        ops.hard_mask(None, 2, 8, 4, 0, device="cpu")
    with pytest.raises(ValueError):
        ops.hard_mask(None, 2, 8, 4, 0)  # no device known
    with pytest.raises(TypeError):
        ops.furthest_point_sample(np.zeros((2, 16, 3), dtype=np.float32), 4)


def test_dropin_modules_have_reference_surface():
    from gm3d_b200.chamfer import ChamferDistanceL1, ChamferDistanceL2, ChamferDistanceL2_split, ChamferFunction
    from gm3d_b200.group import Group, GroupGM3D
    from gm3d_b200.knn import KNN
    g = Group(64, 32)
    assert (g.num_group, g.group_size) == (64, 32) and isinstance(g.knn, KNN) and g.knn.k == 32 and g.knn._t is True
    assert len(g.state_dict()) == 0 and len(list(g.parameters())) == 0
    assert GroupGM3D(64, 32).return_org is True
    assert len(ChamferDistanceL2().state_dict()) == 0
    assert ChamferDistanceL1(ignore_zeros=True).ignore_zeros is True
    assert issubclass(ChamferFunction, torch.autograd.Function)
    ChamferDistanceL2_split()
    with pytest.raises(ValueError):
        ChamferDistanceL2(reduction="bogus")
    with pytest.raises(AssertionError):  # KNN's own batch-size assert, as upstream
        KNN(4, True)(torch.zeros(2, 8, 3), torch.zeros(3, 2, 3))


def test_install_shims_makes_reference_imports_resolve():
    import gm3d_b200
    gm3d_b200.install_shims()
    from extensions.chamfer_dist import ChamferDistanceL1, ChamferDistanceL2  # noqa: F401
    from knn_cuda import KNN
    from pointnet2_ops import pointnet2_utils
    import gm3d_b200.knn
    assert KNN is gm3d_b200.knn.KNN
    assert callable(pointnet2_utils.furthest_point_sample) and callable(pointnet2_utils.gather_operation)


def test_all_reduce_mean_single_process_is_identity():
    from gm3d_b200 import dist as gd
    assert gd.all_reduce_mean(3.5) == 3.5
    v = gd.all_reduce_scalars([torch.tensor(1.0), torch.tensor(2.0)])
    assert v.tolist() == [1.0, 2.0]
    assert gd.all_reduce_stats(torch.zeros(8)) is None


def test_bench_configs_match_baseline_json():
    import json
    import os
    import bench
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base = json.load(open(os.path.join(root, "BASELINE.json")))
    B, N, G, k, ratio, desc = bench.CONFIGS["c2"]
    assert (B, N, G, k) == (128, 1024, 64, 32) and "B=128, N=1024, G=64, k=32" in base["configs"][1]
    assert bench.CONFIGS["c1"][:4] == (8, 1024, 64, 32)
    assert bench.CONFIGS["c4"][:4] == (32, 2048, 128, 32)
    assert bench.CONFIGS["c5"][1:4] == (8192, 512, 32)
    x, lp, pred = bench.synthetic_batch(4, 256, 16, 8, 9, 0)
    assert x.shape == (4, 256, 3) and x.dtype == np.float32 and lp.shape == (4, 16) and pred.shape == (36, 8, 3)
    assert np.abs(x).max() < 2.0


def test_bench_config_dict_is_shared_by_both_arms_and_c3_is_the_m2ae_hierarchy():
    """Both arms print bench.config_dict (same keys and values => the driver's same_config), the ring is larger than
    L2, and --config c3 is BASELINE config[2]: groups 512/256/64, sizes 16/8/8, level l+1 on level l's centres."""
    import json
    import os
    import bench
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base = json.load(open(os.path.join(root, "BASELINE.json")))
    c2 = bench.config_dict(bench.CONFIGS["c2"])
    assert c2["M"] == 39 and set(c2) == {"workload", "B_per_gpu", "N", "G", "k", "M", "l2_policy"}
    B, N, G, k, M = 128, 1024, 64, 32, 39
    assert bench.ring_size(B, N, G, k, M) * bench.buffer_set_bytes(B, N, G, k, M) >= 2 * bench.L2_BYTES
    assert 11e6 < bench.buffer_set_bytes(B, N, G, k, M) < 12e6  # "x 11.3 MB" in l2_policy
    c3 = bench.config_dict(bench.CONFIGS["c3"])
    assert c3["G"] == [512, 256, 64] and c3["k"] == [16, 8, 8] and c3["N"] == 2048 and "512/256/64" in base["configs"][2]
    lv = bench.m2ae_levels(bench.CONFIGS["c3"])
    assert [l[0] for l in lv] == [2048, 512, 256]            # level l+1 groups the G_l centres of level l
    assert [l[3] for l in lv] == c3["M"] == [410, 205, 52]    # G - int(G * (1 - 0.8))
    x, lps, preds = bench.m2ae_inputs((4,) + bench.CONFIGS["c3"][1:], 3)
    assert x.shape == (4, 2048, 3) and [p.shape for p in preds] == [(4 * 410, 16, 3), (4 * 205, 8, 3), (4 * 52, 8, 3)]


def test_reference_arm_prints_the_native_config(capsys):
    """`bench.py --impl reference` (the CPU oracle arm) prints the same `config` dictionary as the native arm."""
    import json
    import types
    import bench
    cfg = (8,) + bench.CONFIGS["c2"][1:]
    bench.run_reference(types.SimpleNamespace(steps=2, warmup=1, gpus=1), cfg)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["config"] == bench.config_dict(cfg)
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
