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
