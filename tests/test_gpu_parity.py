"""Parity tests proper: the CUDA path, called through the C ABI (gm3d_b200.ops / drop-in modules), against
the CPU oracle on identical seeded inputs and against the committed golden fixtures.

Bars (BASELINE.json north_star): FPS and kNN indices bit-exact; Chamfer dist / idx exact (same FP32
expression) and reductions / gradients within 1e-5 relative; masks exact given the same keys.
"""
import numpy as np
import pytest
import torch

from conftest import synthetic_clouds
from oracle import c_oracle as co
from oracle import np_oracle as no

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star: "Chamfer values and gradients within 1e-5 relative"


def dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def host(t):
    return t.detach().cpu().numpy()


def test_native_library_is_loaded(cuda):
    from gm3d_b200 import _lib
    lib = _lib.load()
    assert lib.gm3d_abi_version() == _lib.GM3D_ABI_VERSION
    with open("/proc/self/maps") as f:
        assert "libgm3d_sm100.so" in f.read()


# ------------------------------------------------------------------------------------------ FPS
@pytest.mark.parametrize("kind", ["ball", "sphere"])
@pytest.mark.parametrize("B,N,G", [
    (8, 1024, 64),     # C1
    (4, 2048, 128),    # C4 model
    (2, 2048, 2048),   # C4 pre-sample, G == N
    (3, 333, 17),      # ragged: N % 4 != 0 -> non-bulk loader
    (2, 100, 100),     # tiny, G == N
    (2, 512, 256), (2, 256, 64),   # M2AE levels 1, 2
    (2, 8192, 512),    # C5
    (1, 4096, 40), (2, 127, 9), (1, 1, 1), (1, 33, 40),
])
def test_fps_bit_exact(cuda, kind, B, N, G):
    from gm3d_b200 import ops
    xyz = synthetic_clouds(B, N, 100 + N + G, kind)
    idx, ctr = ops.fps_centers(dev(xyz, cuda), G)
    want = co.fps(xyz, G)
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (B, G)
    assert np.array_equal(host(idx), want)
    assert np.array_equal(host(ctr), np.take_along_axis(xyz, want[..., None].astype(np.int64), axis=1))


def test_fps_two_stage_ties_and_ragged(cuda):
    """4096 < N <= 8192 runs the two-stage arg-max (value first, index only in the warp(s) holding it): ragged N
    (padding slots, non-bulk loader), clouds of few distinct points (every round ties across many warps, and G
    exceeds the number of distinct points) and an all-identical cloud."""
    from gm3d_b200 import ops
    rng = np.random.default_rng(3)
    base = synthetic_clouds(1, 50, 9, "ball")[0]
    clouds = [
        (synthetic_clouds(2, 5001, 21, "sphere"), 64),
        (np.stack([base[rng.integers(0, 50, size=6000)] for _ in range(2)]), 300),
        (np.full((1, 8192, 3), 0.25, dtype=np.float32), 40),
        (synthetic_clouds(1, 4097, 5, "ball"), 4097),
    ]
    for xyz, G in clouds:
        xyz = np.ascontiguousarray(xyz, dtype=np.float32)
        idx, _ = ops.fps_centers(dev(xyz, cuda), G)
        assert np.array_equal(host(idx), co.fps(xyz, G))


def test_fps_large_n_global_kernel(cuda):
    from gm3d_b200 import ops
    xyz = synthetic_clouds(2, 10000, 7, "ball")
    idx, ctr = ops.fps_centers(dev(xyz, cuda), 50)
    want = co.fps(xyz, 50)
    assert np.array_equal(host(idx), want)
    assert np.array_equal(host(ctr), np.take_along_axis(xyz, want[..., None].astype(np.int64), axis=1))


def test_fps_skip_rule_edge(cuda):
    from gm3d_b200 import ops
    xyz = np.zeros((2, 16, 3), dtype=np.float32)
    xyz[:, :, 0] = np.linspace(0.0, 0.02, 16)
    xyz[1, 7] = [1.0, 0, 0]
    xyz[1, 9] = [-1.0, 0, 0]
    got = host(ops.furthest_point_sample(dev(xyz, cuda), 5))
    assert np.array_equal(got, co.fps(xyz, 5))
    assert (got[0] == 0).all() and got[1].tolist()[:4] == [0, 7, 9, 7]


def test_pointnet2_dropin(cuda):
    """furthest_point_sample + gather_operation exactly as miscc.fps calls them (utils/miscc.py:13-20)."""
    from gm3d_b200 import pointnet2_utils
    xyz = synthetic_clouds(4, 1024, 5, "ball")
    data = dev(xyz, cuda)
    fps_idx = pointnet2_utils.furthest_point_sample(data, 64)
    assert not fps_idx.requires_grad
    feats = data.transpose(1, 2).contiguous().requires_grad_(True)
    out = pointnet2_utils.gather_operation(feats, fps_idx)
    fps_data = out.transpose(1, 2).contiguous()
    want_idx = co.fps(xyz, 64)
    assert np.array_equal(host(fps_idx), want_idx)
    assert np.array_equal(host(fps_data), no.fps_centers(xyz, 64))
    assert torch.equal(fps_data, pointnet2_utils.fps(data, 64))
    # backward = deterministic scatter-add, duplicates accumulate
    idx = fps_idx.clone()
    idx[:, :5] = 3
    go = torch.randn(4, 3, 64, device=cuda)
    pointnet2_utils.gather_operation(feats, idx).backward(go)
    assert np.array_equal(host(feats.grad), co.gather_grad(host(go), host(idx), 1024))
    with pytest.raises(RuntimeError):
        pointnet2_utils.furthest_point_sample(data.cpu(), 4)
    with pytest.raises(RuntimeError):
        pointnet2_utils.furthest_point_sample(data.transpose(1, 2), 4)  # non-contiguous


# ------------------------------------------------------------------------------------------ kNN
@pytest.mark.parametrize("kind", ["ball", "sphere"])
@pytest.mark.parametrize("B,N,G,k", [
    (8, 1024, 64, 32), (4, 2048, 128, 32), (2, 2048, 512, 16), (2, 512, 256, 8), (2, 256, 64, 8),
    (3, 333, 17, 5), (2, 8192, 512, 32), (2, 40, 40, 32), (1, 32, 3, 32), (2, 50, 7, 1), (1, 3000, 100, 31),
])
def test_knn_bit_exact(cuda, kind, B, N, G, k):
    from gm3d_b200.knn import KNN
    xyz = synthetic_clouds(B, N, 200 + N + k, kind)
    rng = np.random.default_rng(N + G)
    q = np.stack([xyz[b, rng.choice(N, G, replace=G > N)] for b in range(B)])
    D, I = KNN(k, transpose_mode=True)(dev(xyz, cuda), dev(q, cuda))
    Dw, Iw = co.knn(xyz, q, k)
    assert I.dtype == torch.int64 and D.dtype == torch.float32 and tuple(I.shape) == (B, G, k)
    assert np.array_equal(host(I), Iw)
    assert np.array_equal(host(D).view(np.uint32), Dw.view(np.uint32))


def test_knn_all_ties_and_transpose_mode(cuda):
    from gm3d_b200.knn import KNN
    ref = np.zeros((2, 200, 3), dtype=np.float32)  # every distance equal: order must be by index
    q = np.ones((2, 5, 3), dtype=np.float32)
    _, I = KNN(32, True)(dev(ref, cuda), dev(q, cuda))
    assert (host(I) == np.arange(32)[None, None]).all()
    xyz = synthetic_clouds(2, 300, 1, "ball")
    qq = xyz[:, :11].copy()
    D0, I0 = KNN(9, transpose_mode=False)(dev(xyz.transpose(0, 2, 1), cuda), dev(qq.transpose(0, 2, 1), cuda))
    Dw, Iw = co.knn(xyz, qq, 9)
    assert tuple(I0.shape) == (2, 9, 11)
    assert np.array_equal(host(I0), Iw.transpose(0, 2, 1)) and np.array_equal(host(D0), Dw.transpose(0, 2, 1))
    with pytest.raises(ValueError):
        KNN(301, True)(dev(xyz, cuda), dev(qq, cuda))


@pytest.mark.parametrize("B,N,G,dim,k", [(2, 300, 11, 3, 33), (2, 1024, 20, 3, 64), (1, 5000, 7, 3, 100), (2, 200, 9, 2, 5),
                                         (2, 257, 13, 5, 40), (1, 64, 64, 16, 64), (1, 3, 2, 1, 3)])
def test_knn_general_any_k_any_dim(cuda, B, N, G, dim, k):
    """k > 32 or dim != 3 (upstream KNN_CUDA has no such limits): the general selection kernel, bit-exact against the
    C oracle for dim == 3 and against a NumPy restatement of `ssd += t*t` per dimension (exact FMA) otherwise;
    duplicate points included (ties -> lower index)."""
    from gm3d_b200.knn import KNN
    rng = np.random.default_rng(N + dim + k)
    ref = rng.standard_normal((B, N, dim)).astype(np.float32)
    ref[:, N // 2:N // 2 + N // 8] = ref[:, :N // 8]  # duplicates
    q = np.stack([ref[b, rng.choice(N, G, replace=G > N)] for b in range(B)])
    q[:, ::2] += (rng.standard_normal((B, q[:, ::2].shape[1], dim)) * 0.1).astype(np.float32)
    D, I = KNN(k, transpose_mode=True)(dev(ref, cuda), dev(q, cuda))
    if dim == 3:
        Dw, Iw = co.knn(ref, q, k)
    else:
        d = np.zeros((B, G, N), dtype=np.float32)
        for c in range(dim):
            t = (ref[:, None, :, c] - q[:, :, None, c]).astype(np.float32)
            d = co.fmaf(t, t, d).reshape(B, G, N)
        Iw = np.argsort(d, axis=2, kind="stable")[:, :, :k]
        Dw = np.sqrt(np.take_along_axis(d, Iw, axis=2))
    assert I.dtype == torch.int64 and tuple(I.shape) == (B, G, k)
    assert np.array_equal(host(I), Iw)
    assert np.array_equal(host(D).view(np.uint32), Dw.view(np.uint32))
    Dt, It = KNN(k, transpose_mode=False)(dev(ref.transpose(0, 2, 1), cuda), dev(q.transpose(0, 2, 1), cuda))
    assert np.array_equal(host(It), Iw.transpose(0, 2, 1)) and tuple(Dt.shape) == (B, k, G)


@pytest.mark.parametrize("B,N,G,k", [
    (2, 1025, 33, 32), (1, 2047, 9, 7), (3, 5000, 70, 32), (1, 6145, 64, 16), (2, 8192, 77, 1), (1, 12000, 40, 32),
    (1, 16384, 50, 32), (1, 20000, 24, 32), (40, 2048, 128, 32),
])
def test_knn_large_two_phase_edges(cuda, B, N, G, k):
    """1024 < N <= 16384 runs the two-phase kernel (knn_large.cuh): ragged last chunk, 2..16 chunks (warp roles
    differ for <= 8 and > 8 chunks), query blocks that do not divide G, and N just past its range (streaming kernel)."""
    from gm3d_b200.knn import KNN
    xyz = synthetic_clouds(B, N, 900 + N + k, "sphere")
    rng = np.random.default_rng(N + G)
    q = np.stack([xyz[b, rng.choice(N, G, replace=False)] for b in range(B)])
    q[:, ::3] += rng.standard_normal((B, q[:, ::3].shape[1], 3)).astype(np.float32) * 0.05  # off-cloud queries too
    D, I = KNN(k, transpose_mode=True)(dev(xyz, cuda), dev(q, cuda))
    Dw, Iw = co.knn(xyz, q, k)
    assert np.array_equal(host(I), Iw)
    assert np.array_equal(host(D).view(np.uint32), Dw.view(np.uint32))


@pytest.mark.parametrize("shift,scale", [(0.0, 1.0), (3.0, 1.0), (30.0, 1.0), (1000.0, 1.0), (0.0, 1e-12), (0.0, 1e9), (5e4, 50.0)])
def test_knn_large_screen_margin(cuda, shift, scale):
    """The two-phase kernel screens points with |p|^2 - 2 p.q (three FMAs) and keeps an error margin E ~ (|p|+|q|)^2:
    clouds far from the origin (cancellation: E comparable to or larger than the neighbour distances), tiny and huge
    scales must still give the oracle's indices bit for bit (through the margin or the exact fallback)."""
    from gm3d_b200.knn import KNN
    B, N, G, k = 2, 4096, 96, 32
    xyz = (synthetic_clouds(B, N, 77, "sphere") * np.float32(scale) + np.float32(shift)).astype(np.float32)
    rng = np.random.default_rng(11)
    q = np.stack([xyz[b, rng.choice(N, G, replace=False)] for b in range(B)])
    q[:, ::2] += (rng.standard_normal((B, q[:, ::2].shape[1], 3)) * 0.03 * scale).astype(np.float32)
    D, I = KNN(k, transpose_mode=True)(dev(xyz, cuda), dev(q, cuda))
    Dw, Iw = co.knn(xyz, q, k)
    assert np.array_equal(host(I), Iw)
    assert np.array_equal(host(D).view(np.uint32), Dw.view(np.uint32))


def test_knn_large_heavy_ties(cuda):
    """More than 64 points within the bound (duplicates, lattice distances): the exact streaming fallback inside the
    two-phase kernel, and bit-equal distances among the best k+1 (64-bit key ordering)."""
    from gm3d_b200.knn import KNN
    rng = np.random.default_rng(5)
    N = 4096
    ref = np.zeros((3, N, 3), dtype=np.float32)
    ref[1] = rng.integers(0, 6, size=(N, 3)).astype(np.float32)            # lattice: many equal distances
    base = synthetic_clouds(1, 64, 3, "ball")[0]
    ref[2] = base[rng.integers(0, 64, size=N)]                              # 64 distinct points, 64 copies each
    q = np.stack([np.ones((5, 3), np.float32), ref[1, :5] + 0.5, ref[2, :5]])
    for k in (32, 9):
        D, I = KNN(k, True)(dev(ref, cuda), dev(q, cuda))
        Dw, Iw = co.knn(ref, q, k)
        assert np.array_equal(host(I), Iw)
        assert np.array_equal(host(D).view(np.uint32), Dw.view(np.uint32))
    assert (host(I)[0] == np.arange(9)[None]).all()


@pytest.mark.parametrize("N", [800, 1500, 9000])
def test_knn_nonfinite_inputs_stay_in_bounds(cuda, N):
    """NaN / inf coordinates make the result meaningless (as upstream), but every kernel must stay in bounds: indices in
    [0, N), no CUDA error, and the clouds of the batch without bad points are still exact."""
    from gm3d_b200.knn import KNN
    xyz = synthetic_clouds(3, N, 31, "ball")
    q = xyz[:, :40].copy()
    bad = xyz.copy()
    bad[1, 5::97] = np.nan
    bad[1, 7::131, 1] = np.inf
    D, I = KNN(32, True)(dev(bad, cuda), dev(q, cuda))
    torch.cuda.synchronize()
    Ih = host(I)
    assert Ih.min() >= 0 and Ih.max() < N
    _, Iw = co.knn(xyz, q, 32)
    assert np.array_equal(Ih[0], Iw[0]) and np.array_equal(Ih[2], Iw[2])


# ------------------------------------------------------------------------------------------ Group
@pytest.mark.parametrize("tag", ["c1", "m2ae_l2", "ragged", "g_eq_n"])
def test_group_matches_reference_golden(cuda, golden, tag):
    """The CUDA Group against outputs of the reference's own Group.forward (tests/golden/make_golden.py)."""
    from gm3d_b200.group import Group, GroupGM3D
    xyz = golden[f"group_{tag}_xyz"]
    G, k = golden[f"group_{tag}_G_k"].tolist()
    nb, ctr = Group(G, k)(dev(xyz, cuda))
    nb2, ctr2, org = GroupGM3D(G, k)(dev(xyz, cuda))
    assert np.array_equal(host(ctr), golden[f"group_{tag}_center"])
    assert np.array_equal(host(nb), golden[f"group_{tag}_neighborhood"])
    assert np.array_equal(host(org), golden[f"group_{tag}_neighborhood_org"])
    assert torch.equal(nb, nb2) and torch.equal(ctr, ctr2)
    assert nb.is_contiguous() and tuple(nb.shape) == (xyz.shape[0], G, k, 3)
    assert len(Group(G, k).state_dict()) == 0


@pytest.mark.parametrize("B,N,G,k", [(128, 1024, 64, 32), (32, 2048, 128, 32), (16, 2048, 512, 16), (8, 8192, 512, 32)])
def test_group_full_size_vs_oracle(cuda, B, N, G, k):
    from gm3d_b200 import ops
    xyz = synthetic_clouds(B, N, 1234 + N, "ball")
    r = ops.group(dev(xyz, cuda), G, k, want_org=True, want_idx=True)
    w = co.group(xyz, G, k)
    assert np.array_equal(host(r["fps_idx"]), w["fps_idx"])
    assert np.array_equal(host(r["knn_idx"]), w["knn_idx"])
    assert np.array_equal(host(r["center"]), w["center"])
    assert np.array_equal(host(r["neighborhood"]), w["neighborhood"])
    assert np.array_equal(host(r["neighborhood_org"]), w["neighborhood_org"])


def test_group_properties_at_baseline_size(cuda):
    """Size-independent properties at the C5 shape (oracle-free): centres are cloud points, the first
    neighbour of a centre is the centre itself (distance 0), kNN distances ascend, FPS is a prefix code
    (the first G' picks of a G-sample equal the G'-sample)."""
    from gm3d_b200 import ops
    B, N, G, k = 16, 8192, 512, 32
    xyz = dev(synthetic_clouds(B, N, 99, "ball"), cuda)
    r = ops.group(xyz, G, k, want_org=True, want_idx=True)
    ctr = torch.gather(xyz, 1, r["fps_idx"].long()[..., None].expand(-1, -1, 3))
    assert torch.equal(ctr, r["center"])
    assert torch.equal(r["neighborhood_org"][:, :, 0], r["center"])
    assert (r["neighborhood"][:, :, 0] == 0).all()
    gathered = torch.gather(xyz, 1, r["knn_idx"].reshape(B, -1, 1).expand(-1, -1, 3)).reshape(B, G, k, 3)
    assert torch.equal(gathered, r["neighborhood_org"])
    d = (r["neighborhood"].double() ** 2).sum(-1)
    assert (d[..., 1:] >= d[..., :-1] - 1e-9).all()
    assert torch.equal(ops.furthest_point_sample(xyz, 64), r["fps_idx"][:, :64])
    D, I = ops.knn(xyz, r["center"], k)
    assert torch.equal(I, r["knn_idx"])


# ------------------------------------------------------------------------------------------ Chamfer
@pytest.mark.parametrize("P,n,m", [(304, 32, 32), (4992, 32, 32), (1000, 16, 16), (777, 8, 8), (5, 20, 31), (3, 1, 1),
                                   (6, 100, 70), (2, 700, 300), (3, 33, 8)])
def test_chamfer_forward_backward(cuda, P, n, m):
    from gm3d_b200 import ops
    rng = np.random.default_rng(P + n)
    a = rng.standard_normal((P, n, 3)).astype(np.float32) * 0.1
    b = (a + 0.02 * rng.standard_normal((P, n, 3))).astype(np.float32) if n == m else \
        rng.standard_normal((P, m, 3)).astype(np.float32) * 0.1
    for norm in (2, 1):
        d1, d2, i1, i2, pp, tot = ops.chamfer_forward(dev(a, cuda), dev(b, cuda), norm=norm, want_per_patch=True,
                                                      want_total=True)
        w1, w2, wi1, wi2 = co.chamfer_fwd(a, b)
        assert np.array_equal(host(d1).view(np.uint32), w1.view(np.uint32))
        assert np.array_equal(host(d2).view(np.uint32), w2.view(np.uint32))
        assert np.array_equal(host(i1), wi1) and np.array_equal(host(i2), wi2)
        wpp = co.chamfer_per_patch(w1, w2, norm)
        assert np.allclose(host(pp), wpp, rtol=RTOL, atol=0)
        assert np.isclose(host(tot)[0], wpp.mean(), rtol=RTOL, atol=0)
    g1 = rng.standard_normal((P, n)).astype(np.float32)
    g2 = rng.standard_normal((P, m)).astype(np.float32)
    ga, gb = ops.chamfer_backward(dev(a, cuda), dev(b, cuda), i1, i2, dev(g1, cuda), dev(g2, cuda))
    wa, wb = co.chamfer_bwd(a, b, wi1, wi2, g1, g2)
    # same summation order as the C oracle -> bit-exact; and within RTOL of the float64 gradient
    assert np.array_equal(host(ga), wa) and np.array_equal(host(gb), wb)
    ta, tb = no.chamfer_bwd(a, b, wi1, wi2, g1, g2)
    scale = max(np.abs(ta).max(), np.abs(tb).max())
    assert np.abs(host(ga) - ta).max() <= RTOL * scale and np.abs(host(gb) - tb).max() <= RTOL * scale
    ga_only, none = ops.chamfer_backward(dev(a, cuda), dev(b, cuda), i1, i2, dev(g1, cuda), dev(g2, cuda), want_grad2=False)
    assert none is None and torch.equal(ga_only, ga)


@pytest.mark.parametrize("P,n,m", [(4992, 32, 32), (1000, 16, 16), (777, 8, 8), (5, 20, 31), (3, 1, 1), (301, 32, 17)])
def test_chamfer_fused_forward_backward(cuda, P, n, m):
    """gm3d_chamfer_fused_f32 (fwd + bwd of the mean in one launch, loss reduced by the last CTA) against the
    oracle: dist / idx exact, gradients bit-exact for L2 (same summation order), loss + stats within 1e-5."""
    from gm3d_b200 import ops
    rng = np.random.default_rng(P + n + m)
    a = rng.standard_normal((P, n, 3)).astype(np.float32) * 0.1
    b = (a + 0.02 * rng.standard_normal((P, n, 3))).astype(np.float32) if n == m else \
        rng.standard_normal((P, m, 3)).astype(np.float32) * 0.1
    w1, w2, wi1, wi2 = co.chamfer_fwd(a, b)
    for norm in (2, 1):
        gs1, gs2 = (1.0 / (P * n), 1.0 / (P * m)) if norm == 2 else (0.5 / (P * n), 0.5 / (P * m))
        for rep in range(2):  # twice: the self-resetting ticket must leave the workspace reusable
            r = ops.chamfer_fused(dev(a, cuda), dev(b, cuda), gs1, gs2, norm=norm, want_grad2=True, want_dist=True)
        assert np.array_equal(host(r["dist1"]), w1) and np.array_equal(host(r["dist2"]), w2)
        assert np.array_equal(host(r["idx1"]), wi1) and np.array_equal(host(r["idx2"]), wi2)
        wpp = co.chamfer_per_patch(w1, w2, norm)
        assert np.allclose(host(r["per_patch"]), wpp, rtol=RTOL, atol=0)
        assert np.isclose(host(r["total"])[0], wpp.mean(), rtol=RTOL, atol=0)
        st = host(r["stats"])
        assert np.allclose(st[:6], [wpp.sum(), (wpp * wpp).sum(), P, wpp.min(), wpp.max(), wpp.mean()], rtol=RTOL)
        if norm == 2:
            g1 = np.full((P, n), np.float32(gs1), dtype=np.float32)
            g2 = np.full((P, m), np.float32(gs2), dtype=np.float32)
            wa, wb = co.chamfer_bwd(a, b, wi1, wi2, g1, g2)
            assert np.array_equal(host(r["grad1"]), wa) and np.array_equal(host(r["grad2"]), wb)
        else:
            with np.errstate(divide="ignore"):
                g1 = np.float32(gs1) * (np.float32(0.5) / np.sqrt(w1))
                g2 = np.float32(gs2) * (np.float32(0.5) / np.sqrt(w2))
            ok = np.isfinite(g1).all(axis=1) & np.isfinite(g2).all(axis=1)  # d == 0 -> inf upstream, as in torch
            ta, tb = no.chamfer_bwd(a[ok], b[ok], wi1[ok], wi2[ok], g1[ok], g2[ok])
            scale = max(np.abs(ta).max(), np.abs(tb).max()) if ok.any() else 1.0
            assert np.abs(host(r["grad1"])[ok] - ta).max(initial=0) <= 1e-4 * scale
            assert np.abs(host(r["grad2"])[ok] - tb).max(initial=0) <= 1e-4 * scale
    # indexed target pool (masked-patch select folded into the load), grad w.r.t. the prediction only
    if n == m:
        pool = rng.standard_normal((2 * P, m, 3)).astype(np.float32) * 0.1
        index = rng.permutation(2 * P)[:P].astype(np.int32)
        r = ops.chamfer_fused(dev(a, cuda), dev(pool, cuda), 1.0 / (P * n), 1.0 / (P * m), xyz2_index=dev(index, cuda))
        v1, v2, vi1, vi2 = co.chamfer_fwd(a, pool[index])
        g = np.full((P, n), np.float32(1.0 / (P * n)), dtype=np.float32)
        va, _ = co.chamfer_bwd(a, pool[index], vi1, vi2, g, g)
        assert np.array_equal(host(r["grad1"]), va)
        assert np.isclose(host(r["total"])[0], co.chamfer_per_patch(v1, v2, 2).mean(), rtol=RTOL)


def test_hard_mask_patch_index(cuda):
    from gm3d_b200 import ops
    rng = np.random.default_rng(8)
    for B, L, len_keep, len_loss in [(128, 64, 25, 15), (16, 512, 205, 100), (3, 33, 5, 7), (4, 64, 25, 0)]:
        lp = rng.standard_normal((B, L)).astype(np.float32)
        rk = rng.random((B, L)).astype(np.float32)
        mask, index = ops.hard_mask(dev(lp, cuda), B, L, len_keep, len_loss, rand_keys=dev(rk, cuda), want_index=True)
        want = co.hard_mask(lp, len_keep, len_loss, rk)
        assert np.array_equal(host(mask), want)
        assert np.array_equal(host(index), np.flatnonzero(want.reshape(-1)))


def test_chamfer_modules_match_stock_autograd(cuda):
    """ChamferDistanceL2 / L1 / L2_split and their backward against a plain-PyTorch fp64 restatement."""
    from gm3d_b200.chamfer import ChamferDistanceL1, ChamferDistanceL2, ChamferDistanceL2_split, ChamferFunction
    torch.manual_seed(0)
    P, n = 304, 32
    gt = torch.randn(P, n, 3, device=cuda) * 0.1
    pred = (gt + 0.02 * torch.randn(P, n, 3, device=cuda)).requires_grad_(True)

    def ref(pred64, gt64, l1):
        d = ((pred64[:, :, None] - gt64[:, None]) ** 2).sum(-1)
        d1, d2 = d.min(2).values, d.min(1).values
        return (d1.sqrt().mean() + d2.sqrt().mean()) / 2 if l1 else d1.mean() + d2.mean()

    for mod, l1 in ((ChamferDistanceL2(), False), (ChamferDistanceL1(), True)):
        pred.grad = None
        loss = mod(pred, gt)
        loss.backward()
        p64 = pred.detach().double().requires_grad_(True)
        want = ref(p64, gt.double(), l1)
        want.backward()
        assert loss.dim() == 0
        assert abs(loss.item() - want.item()) <= RTOL * abs(want.item())
        assert (pred.grad.double() - p64.grad).abs().max() <= RTOL * p64.grad.abs().max()
    s1, s2 = ChamferDistanceL2_split()(pred, gt)
    assert abs((s1 + s2).item() - ChamferDistanceL2()(pred, gt).item()) <= RTOL * (s1 + s2).item()
    # reductions used by GM3D
    d1, d2 = ChamferFunction.apply(pred, gt)
    assert torch.allclose(ChamferDistanceL2(reduction="patch")(pred, gt), d1.mean(1) + d2.mean(1), rtol=RTOL, atol=0)
    assert torch.equal(ChamferDistanceL2(reduction="dist1")(pred, gt), d1)
    assert torch.equal(ChamferDistanceL2(reduction="sum")(pred, gt), d1 + d2)
    # gradient flows to both inputs through ChamferFunction
    gt2 = gt.clone().requires_grad_(True)
    a, b = ChamferFunction.apply(pred, gt2)
    (a.sum() + b.sum()).backward()
    assert gt2.grad is not None and gt2.grad.abs().sum() > 0
    # ignore_zeros (batch of one)
    x1 = torch.randn(1, 40, 3, device=cuda)
    x1[0, 5:9] = 0
    x2 = torch.randn(1, 30, 3, device=cuda)
    v = ChamferDistanceL2(ignore_zeros=True)(x1, x2)
    keep = x1[0].sum(1) != 0
    assert abs(v.item() - ChamferDistanceL2()(x1[:, keep], x2).item()) <= RTOL * abs(v.item())


def test_chamfer_roundtrip_properties(cuda):
    """Oracle-free properties at the C5 patch count: identical clouds -> zero distance and identity
    arg-min; a permutation of the target only permutes idx; the scalar is symmetric in its arguments."""
    from gm3d_b200 import ops
    P, n = 39296, 32
    torch.manual_seed(1)
    a = torch.randn(P, n, 3, device=cuda)
    d1, d2, i1, i2, pp, tot = ops.chamfer_forward(a, a.clone(), want_per_patch=True, want_total=True)
    ar = torch.arange(n, device=cuda, dtype=torch.int32)
    assert (d1 == 0).all() and (d2 == 0).all() and (i1 == ar).all() and (i2 == ar).all() and tot.item() == 0
    perm = torch.randperm(n, device=cuda)
    b = a + 0.05 * torch.randn_like(a)
    e1, e2, j1, j2, _, t1 = ops.chamfer_forward(a, b, want_total=True)
    f1, f2, k1, k2, _, _ = ops.chamfer_forward(a, b[:, perm].contiguous())
    assert torch.equal(e1, f1) and torch.equal(perm[k1.long()], j1.long())
    _, _, _, _, _, t2 = ops.chamfer_forward(b, a, want_total=True)
    assert abs(t1.item() - t2.item()) <= RTOL * t1.item()


# ------------------------------------------------------------------------------------------ masks / select / loss glue
def test_hard_mask_matches_reference_golden(cuda, golden):
    from gm3d_b200 import ops
    lp = golden["mask_loss_pred"]
    B, L = lp.shape
    for key in golden["mask_cases"].tolist():
        _, cls, e, t, a = key.split("_")
        ref = golden[key]
        len_keep, len_loss = no.mask_lengths(L, 0.6, int(e[1:]), int(t[1:]), True, bool(int(a[1:])) or None,
                                             0.8 if cls == "fb" else 0.5)
        order = np.argsort(lp, axis=1, kind="stable")
        top = np.zeros((B, L), dtype=bool)
        if len_loss:
            np.put_along_axis(top, order[:, L - len_loss:], True, axis=1)
        keys = ((ref == 1) & ~top).astype(np.float32)  # replay the reference's random picks
        got = ops.hard_mask(dev(lp, cuda), B, L, len_keep, len_loss, rand_keys=dev(keys, cuda))
        assert np.array_equal(host(got), ref.astype(np.uint8)), key


def test_generate_mask_dropin(cuda):
    from gm3d_b200 import masking
    torch.manual_seed(3)
    lp = torch.randn(128, 64, device=cuda)
    for epoch in (0, 50, 199, 399):
        m = masking.generate_mask(lp, mask_ratio=0.6, epoch=epoch, total_epoch=400)
        assert m.dtype == torch.float32 and tuple(m.shape) == (128, 64)
        assert (m.sum(1) == 39).all()
        len_keep, len_loss = masking.mask_lengths(64, 0.6, epoch, 400)
        if len_loss:
            top = lp.argsort(dim=1)[:, -len_loss:]
            assert (torch.gather(m, 1, top) == 1).all()
    # philox path: reproducible from (seed, offset), different for different offsets, and uniform
    a = masking.generate_mask(lp, 0.6, epoch=0, total_epoch=400, seed=7, offset=0)
    b = masking.generate_mask(lp, 0.6, epoch=0, total_epoch=400, seed=7, offset=0)
    c = masking.generate_mask(lp, 0.6, epoch=0, total_epoch=400, seed=7, offset=128 * 64)
    assert torch.equal(a, b) and not torch.equal(a, c)
    many = torch.stack([masking.generate_mask(lp, 0.6, epoch=0, total_epoch=400, seed=11, offset=i * 8192) for i in range(40)])
    freq = many.mean(dim=(0, 1))
    assert (freq - 39 / 64).abs().max() < 0.03
    # explicit keys == oracle
    keys = torch.rand(128, 64, device=cuda)
    got = masking.generate_mask(lp, 0.6, epoch=199, total_epoch=400, rand_keys=keys)
    lk, ll = masking.mask_lengths(64, 0.6, 199, 400)
    assert np.array_equal(host(got).astype(np.uint8), co.hard_mask(host(lp), lk, ll, host(keys)))
    rm = masking.mask_center_rand(torch.zeros(8, 64, 3, device=cuda), 0.6)
    assert rm.dtype == torch.bool and (rm.sum(1) == 38).all()
    assert not masking.mask_center_rand(torch.zeros(8, 64, 3, device=cuda), 0.6, noaug=True).any()


def test_mask_center_block_dropin(cuda):
    """_mask_center_block (models/Point_MAE.py:268-295): the int(ratio * G) centres nearest to one random centre per
    cloud, with the reference's own `random.randint` stream, against a NumPy restatement of the reference lines."""
    import random
    from gm3d_b200 import masking
    B, G, ratio = 9, 64, 0.6
    c = synthetic_clouds(B, G, 123)
    random.seed(7)
    got = host(masking.mask_center_block(dev(c, cuda), ratio))
    random.seed(7)
    for b in range(B):
        index = random.randint(0, G - 1)
        d = np.linalg.norm(c[b, index].reshape(1, 3) - c[b], axis=-1)
        want = np.zeros(G, dtype=bool)
        want[np.argsort(d, kind="stable")[: int(ratio * G)]] = True
        assert np.array_equal(got[b], want) and got[b, index]
    assert not masking.mask_center_block(dev(c, cuda), ratio, noaug=True).any()
    # against the reference method's own output (tests/golden/reference_block_mask.npz), same `random` stream
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_block_mask.npz"))
    for tag in g["cases"]:
        random.seed(int(g[f"{tag}_seed"][0]))
        got = host(masking.mask_center_block(dev(g[f"{tag}_centers"], cuda), float(g[f"{tag}_ratio"][0])))
        assert np.array_equal(got, g[f"{tag}_mask"]), tag
    idx = torch.arange(B) % G
    assert host(masking.mask_center_block(dev(c, cuda), 0.25, index=idx)).sum(1).tolist() == [16] * B


def test_hard_mask_ties_and_sizes(cuda):
    from gm3d_b200 import ops
    rng = np.random.default_rng(5)
    for L, len_keep, len_loss in [(64, 25, 15), (64, 25, 0), (64, 25, 39), (512, 205, 100), (33, 5, 7), (1, 0, 1)]:
        lp = rng.integers(0, 5, (9, L)).astype(np.float32)
        rk = rng.integers(0, 3, (9, L)).astype(np.float32)
        got = ops.hard_mask(dev(lp, cuda), 9, L, len_keep, len_loss, rand_keys=dev(rk, cuda))
        assert np.array_equal(host(got), co.hard_mask(lp, len_keep, len_loss, rk))
    with pytest.raises(ValueError):
        ops.hard_mask(dev(lp, cuda), 9, 1, 0, 2)


def test_select_patches(cuda):
    from gm3d_b200 import ops
    rng = np.random.default_rng(6)
    B, G, k, M = 16, 64, 32, 39
    nb = rng.standard_normal((B, G, k, 3)).astype(np.float32)
    mask = co.hard_mask(rng.standard_normal((B, G)).astype(np.float32), G - M, 10, rng.random((B, G)).astype(np.float32))
    out, idx = ops.select_patches(dev(nb, cuda), dev(mask, cuda), M, want_index=True)
    assert np.array_equal(host(out), nb[mask.astype(bool)].reshape(B * M, k, 3))
    assert np.array_equal(host(idx), np.flatnonzero(mask.reshape(-1)))
    vis, _ = ops.select_patches(dev(nb, cuda), dev(mask.astype(bool), cuda), G - M, invert=True)
    assert np.array_equal(host(vis), nb[~mask.astype(bool)].reshape(B * (G - M), k, 3))
    status = torch.zeros(1, dtype=torch.int32, device=cuda)
    bad = mask.copy()
    bad[3, np.flatnonzero(bad[3])[0]] = 0
    ops.select_patches(dev(nb, cuda), dev(bad, cuda), M, status=status)
    assert status.item() == 4  # row index + 1


def test_forward_loss_matches_reference_golden(cuda, golden):
    """forward_loss (usual + feature mode) against the reference's own forward_loss outputs."""
    from gm3d_b200 import loss
    nb, mask, pred = dev(golden["loss_neighborhood"], cuda), dev(golden["loss_mask"], cuda), dev(golden["loss_pred_points"], cuda)
    for mode in ("dist1", "sum"):
        p = pred.clone().requires_grad_(True)
        r = loss.forward_loss_usual(p, nb, mask, per_point=mode)
        assert np.allclose(host(r["matrix"]), golden[f"loss_usual_{mode}_matrix"], rtol=RTOL, atol=1e-9)
        assert np.isclose(r["Chamfer_mean"].item(), golden[f"loss_usual_{mode}_chamfer_mean"], rtol=RTOL)
        r["Chamfer_mean"].backward()
        # gradient against autograd through a dense float64 restatement
        p64 = pred.double().reshape(-1, 32, 3).requires_grad_(True)
        gt64 = nb[mask].double().reshape(-1, 32, 3)
        d = ((p64[:, :, None] - gt64[:, None]) ** 2).sum(-1)
        want = d.min(2).values if mode == "dist1" else d.min(2).values + d.min(1).values
        want.mean().backward()
        assert (p.grad.reshape(-1, 32, 3).double() - p64.grad).abs().max() <= RTOL * p64.grad.abs().max()
        fp = dev(golden["loss_feature_pred"], cuda).clone().requires_grad_(True)
        r2 = loss.forward_loss_feature(fp, dev(golden["loss_feature_target"], cuda), mask, nb, pred, per_point=mode)
        assert np.allclose(host(r2["matrix"]), golden[f"loss_feature_{mode}_matrix"], rtol=1e-4, atol=1e-6)
        assert np.isclose(r2["Chamfer_mean"].item(), golden[f"loss_feature_{mode}_chamfer_mean"], rtol=RTOL)
        assert np.isclose(r2["MSE_mean"].item(), golden[f"loss_feature_{mode}_mse_mean"], rtol=1e-5)
        # gradient of the normalised-feature MSE (gm3d_feature_mse_f32) against float64 autograd of the reference lines
        wgt = torch.rand(r2["matrix"].shape, device=cuda)
        (r2["MSE_mean"] + (r2["matrix"].detach() * 0 + wgt * (r2["matrix"] - r2["matrix"].detach())).sum()).backward()
        f64 = dev(golden["loss_feature_pred"], cuda).double().requires_grad_(True)
        t64 = dev(golden["loss_feature_target"], cuda).double()[mask].reshape(f64.shape)
        m64 = ((torch.nn.functional.normalize(f64, dim=-1) - torch.nn.functional.normalize(t64, dim=-1)) ** 2).sum(-1)
        (m64.mean() + (wgt.double() * m64).sum()).backward()
        assert (fp.grad.double() - f64.grad).abs().max() <= 1e-5 * f64.grad.abs().max()
    # stock scalar losses (models/Point_MAE.py:426)
    from gm3d_b200.chamfer import ChamferDistanceL1, ChamferDistanceL2
    gt = nb[mask].reshape(-1, 32, 3)
    assert np.isclose(ChamferDistanceL2()(pred.reshape(-1, 32, 3), gt).item(), golden["cdl2_scalar"], rtol=RTOL)
    assert np.isclose(ChamferDistanceL1()(pred.reshape(-1, 32, 3), gt).item(), golden["cdl1_scalar"], rtol=RTOL)


def test_forward_loss_fused_default_agrees_with_the_step(cuda):
    """One per-patch convention everywhere: forward_loss_usual(...)['matrix'] is bit-identical to
    GroupLossStep.per_patch.view(B, M) on the same inputs, 'Chamfer_mean' to step.total, and backward() through the
    fused autograd Function returns the step's grad_pred; a gradient sent through 'matrix' matches float64 autograd."""
    from gm3d_b200 import loss
    from gm3d_b200.group import Group
    from gm3d_b200.pipeline import GroupLossStep
    B, N, G, k = 6, 1024, 64, 32
    rng = np.random.default_rng(41)
    s = GroupLossStep(B, N, G, k, 0.6, device=cuda, seed=7, rand_offset=5)
    x = synthetic_clouds(B, N, 55)
    s.xyz.copy_(dev(x, cuda))
    s.loss_pred.copy_(dev(rng.standard_normal((B, G)).astype(np.float32), cuda))
    s.pred.copy_(dev((rng.standard_normal((s.P, k, 3)) * 0.08).astype(np.float32), cuda))
    s.run()
    torch.cuda.synchronize()
    nb, _ = Group(G, k)(s.xyz)
    assert torch.equal(nb, s.neighborhood)
    pred = s.pred.clone().reshape(B, s.M, k * 3).requires_grad_(True)
    r = loss.forward_loss_usual(pred, nb, s.mask.view(torch.bool))
    assert torch.equal(r["matrix"], s.per_patch.view(B, s.M))
    assert torch.equal(r["Chamfer_mean"], s.total.reshape(()))
    r["Chamfer_mean"].backward()
    assert torch.equal(pred.grad.reshape(s.P, k, 3), s.grad_pred)
    # a gradient through the per-patch matrix (the reference detaches it; autograd must still be right)
    pred2 = s.pred.clone().requires_grad_(True)
    r2 = loss.forward_loss_usual(pred2, nb, s.mask.view(torch.bool))
    wgt = torch.rand(B, s.M, device=cuda)
    (r2["matrix"] * wgt).sum().backward()
    p64 = s.pred.double().requires_grad_(True)
    gt64 = nb.reshape(B * G, k, 3)[s.patch_index.long()].double()
    d = ((p64[:, :, None] - gt64[:, None]) ** 2).sum(-1)
    ((d.min(2).values.mean(1) + d.min(1).values.mean(1)) * wgt.reshape(-1).double()).sum().backward()
    assert (pred2.grad.double() - p64.grad).abs().max() <= RTOL * p64.grad.abs().max()


def test_three_launches_with_networks_in_between(cuda):
    """the step as a training loop issues it -- group launch, [teacher], mask launch,
    [student], loss launch -- on the caller's stream, inputs arriving between the launches, equals the captured step."""
    from gm3d_b200.pipeline import GroupLossStep
    B, N, G, k = 5, 1024, 64, 32
    rng = np.random.default_rng(61)
    x = synthetic_clouds(B, N, 91)
    lp = rng.standard_normal((B, G)).astype(np.float32)
    a = GroupLossStep(B, N, G, k, 0.6, device=cuda, seed=2, rand_offset=9)
    b = GroupLossStep(B, N, G, k, 0.6, device=cuda, seed=2, rand_offset=9)
    pred = (rng.standard_normal((a.P, k, 3)) * 0.08).astype(np.float32)
    for s in (a, b):
        s.loss_pred.fill_(0); s.pred.fill_(0)
    b.xyz.copy_(dev(x, cuda)); b.loss_pred.copy_(dev(lp, cuda)); b.pred.copy_(dev(pred, cuda))
    b.capture().run()
    st = torch.cuda.Stream(cuda)
    with torch.cuda.stream(st):
        a.xyz.copy_(dev(x, cuda), non_blocking=True)
        a.enqueue_group()
        teacher_out = a.neighborhood.sum() * 0 + dev(lp, cuda)      # "teacher": consumes the grouping, emits loss_pred
        a.loss_pred.copy_(teacher_out)
        a.enqueue_mask()
        student_out = dev(pred, cuda) + a.mask.sum() * 0              # "student": consumes grouping + mask, emits pred
        a.pred.copy_(student_out)
        a.enqueue_loss()
    torch.cuda.synchronize()
    for n in ("fps_idx", "center", "neighborhood", "mask", "patch_index", "dist1", "idx2", "per_patch", "total", "stats", "grad_pred"):
        assert torch.equal(getattr(a, n), getattr(b, n)), n


def test_loss_stats(cuda):
    from gm3d_b200 import ops
    v = torch.rand(4992, device=cuda)
    s = host(ops.loss_stats(v))
    h = host(v).astype(np.float64)
    assert np.allclose(s[:6], [h.sum(), (h * h).sum(), 4992, h.min(), h.max(), h.mean()], rtol=1e-6)
    assert s[6] == 0 and s[7] == 0


def test_reference_import_lines_resolve(cuda):
    """The reference's own import statements work unchanged after install_shims()."""
    import gm3d_b200
    gm3d_b200.install_shims()
    from extensions.chamfer_dist import ChamferDistanceL1, ChamferDistanceL2  # noqa: F401  (models/Point_MAE.py:13)
    from knn_cuda import KNN  # noqa: F401                                                  (models/Point_MAE.py:12)
    from pointnet2_ops import pointnet2_utils  # noqa: F401                                 (utils/miscc.py:10)
    xyz = dev(synthetic_clouds(2, 128, 1), cuda)
    idx = pointnet2_utils.furthest_point_sample(xyz, 8)
    ctr = pointnet2_utils.gather_operation(xyz.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    _, I = KNN(4, transpose_mode=True)(xyz, ctr)
    assert tuple(I.shape) == (2, 8, 4)
    assert ChamferDistanceL2()(xyz, xyz).item() == 0


# ------------------------------------------------------------------------------------------ fused per-cloud step
@pytest.mark.parametrize("kind", ["ball", "sphere"])
@pytest.mark.parametrize("B,N,G,k,ratio,norm", [
    (8, 1024, 64, 32, 0.6, 2),      # BASELINE config[0] / config[1] shape
    (5, 2048, 128, 32, 0.6, 2),     # two tiles: bootstrap + streamed tile
    (3, 777, 50, 16, 0.6, 1),       # ragged N (no bulk copy), k < 32, L1
    (4, 512, 256, 8, 0.8, 2),       # M2AE level: mask over 256 patches (shared-memory sort path)
    (2, 2048, 512, 16, 0.8, 2),     # M2AE level 0
    (6, 96, 40, 32, 0.5, 2),        # tiny cloud: bootstrap overflows, streaming fallback
])
def test_cloud_step_fused_equals_kernel_sequence_and_oracle(cuda, kind, B, N, G, k, ratio, norm):
    """gm3d_cloud_step_f32 (one launch, one CTA per cloud) against the four-kernel sequence (bit-identical
    on every output) and against the CPU oracle (indices exact, loss / gradients within 1e-5)."""
    from gm3d_b200.pipeline import GroupLossStep
    x = synthetic_clouds(B, N, 77 + N + G, kind)
    rng = np.random.default_rng(5)
    steps = [GroupLossStep(B, N, G, k, ratio, device=cuda, seed=11, rand_offset=3, norm=norm, fused=f) for f in (True, False, None)]
    M = steps[0].M
    lp = rng.standard_normal((B, G)).astype(np.float32)
    pred = (rng.standard_normal((B * M, k, 3)) * 0.08).astype(np.float32)
    for s in steps:
        s.xyz.copy_(dev(x, cuda)); s.loss_pred.copy_(dev(lp, cuda)); s.pred.copy_(dev(pred, cuda))
        for t in (s.fps_idx, s.center, s.neighborhood, s.mask, s.patch_index, s.dist1, s.dist2, s.idx1, s.idx2,
                  s.per_patch, s.total, s.grad_pred, s.stats):
            t.fill_(0)  # outputs must be fully written by the step itself
        s.run()
    torch.cuda.synchronize()
    f, u, d = steps  # single launch; fps + knn + mask + chamfer; the default dataflow path (group -> mask -> chamfer)
    assert f.kernels_per_step == 1 and u.kernels_per_step == 4 and d.path == "dataflow" and not d.fused
    assert d.kernels_per_step == (3 if G <= 128 else 4)
    for name in ("fps_idx", "center", "neighborhood", "mask", "patch_index", "dist1", "dist2", "idx1", "idx2",
                 "per_patch", "total", "stats", "grad_pred"):
        assert np.array_equal(host(getattr(f, name)), host(getattr(u, name))), name
        assert np.array_equal(host(getattr(f, name)), host(getattr(d, name))), name
    # oracle
    w = co.group(x, G, k)
    assert np.array_equal(host(f.fps_idx), w["fps_idx"])
    assert np.array_equal(host(f.center), w["center"])
    assert np.array_equal(host(f.neighborhood), w["neighborhood"])
    mask = host(f.mask).astype(bool)
    assert (mask.sum(1) == M).all()
    gt = w["neighborhood"][mask]
    d1, d2, i1, i2 = co.chamfer_fwd(pred, gt)
    assert np.array_equal(host(f.dist1), d1) and np.array_equal(host(f.dist2), d2)
    assert np.array_equal(host(f.idx1), i1) and np.array_equal(host(f.idx2), i2)
    pp = co.chamfer_per_patch(d1, d2, norm)
    assert np.allclose(host(f.per_patch), pp, rtol=RTOL, atol=0)
    assert abs(f.total.item() - pp.mean()) <= RTOL * abs(pp.mean())
    if norm == 2:
        g = np.full((B * M, k), 1.0 / (B * M * k), dtype=np.float32)
        ga, _ = co.chamfer_bwd(pred, gt, i1, i2, g, g)
        assert np.abs(host(f.grad_pred) - ga).max() <= RTOL * np.abs(ga).max()
    else:  # L1: d mean(sqrt d)/2 / d dist = 0.5 / (P k) * 0.5 / sqrt(d); float64 oracle, patches without a zero distance
        with np.errstate(divide="ignore"):
            g1 = np.float32(0.5 / (B * M * k)) * (np.float32(0.5) / np.sqrt(d1))
            g2 = np.float32(0.5 / (B * M * k)) * (np.float32(0.5) / np.sqrt(d2))
        ok = np.isfinite(g1).all(axis=1) & np.isfinite(g2).all(axis=1)
        ta, _ = no.chamfer_bwd(pred[ok], gt[ok], i1[ok], i2[ok], g1[ok], g2[ok])
        assert ok.any() and np.abs(host(f.grad_pred)[ok] - ta).max() <= 1e-4 * np.abs(ta).max()
    # second run through a captured graph gives the same bits (ticket self-reset, no stale state)
    before = {n: host(getattr(f, n)).copy() for n in ("neighborhood", "grad_pred", "total", "mask")}
    f.capture()
    f.run(); f.run()
    torch.cuda.synchronize()
    for n, v in before.items():
        assert np.array_equal(host(getattr(f, n)), v), n


def test_cloud_step_rejects_unsupported(cuda):
    from gm3d_b200 import _lib
    from gm3d_b200.pipeline import GroupLossStep
    with pytest.raises(NotImplementedError):
        GroupLossStep(2, 4096, 64, 32, device=cuda, fused=True)
    with pytest.raises(NotImplementedError):
        GroupLossStep(2, 4096, 64, 32, device=cuda, path="single")
    s = GroupLossStep(2, 4096, 64, 32, device=cuda)  # the dataflow path with separate fps / kNN kernels
    assert not s.fused and not s.group_per_cloud and s.kernels_per_step == 4
    c2 = GroupLossStep(128, 1024, 64, 32, device=cuda)                   # BASELINE config[1]: group -> mask -> chamfer
    assert c2.path == "dataflow" and c2.group_per_cloud and c2.kernels_per_step == 3
    assert GroupLossStep(128, 1024, 64, 32, device=cuda, path="single").fused
    assert not GroupLossStep(128, 2048, 512, 16, device=cuda).group_per_cloud  # M2AE level 0: too much selection work per CTA
    lib = _lib.load()
    p = s.xyz.data_ptr()
    assert lib.gm3d_cloud_step_f32(p, 2, 4096, 64, 32, p, p, None, p, None, None, 0, 0, None, 0, 0, None, None, None,
                                   0.0, 0.0, 2, None, None, None, None, None, None, None, None, 0, None, None, None) == _lib.GM3D_ENOSUP


def test_step_ring_overlapped_steps_equal_serial_steps(cuda):
    """A ring of buffer sets replayed as one graph with programmatic dependent launch between the fused kernels
    (steps overlap tail-to-head) gives, for every step, the bits of the same step run alone."""
    from gm3d_b200.pipeline import GroupLossStep, StepRing
    B, N, G, k = 16, 1024, 64, 32
    rng = np.random.default_rng(9)
    steps, want = [], []
    for r in range(9):  # >= 8 steps of a small batch: the ring runs them as four chains on forked streams
        s = GroupLossStep(B, N, G, k, 0.6, device=cuda, seed=4, rand_offset=r * B * G, fused=True)
        s.xyz.copy_(dev(synthetic_clouds(B, N, 300 + r), cuda))
        s.loss_pred.copy_(dev(rng.standard_normal((B, G)).astype(np.float32), cuda))
        s.pred.copy_(dev((rng.standard_normal((s.P, k, 3)) * 0.08).astype(np.float32), cuda))
        s.run()
        torch.cuda.synchronize()
        want.append({n: host(getattr(s, n)).copy() for n in ("fps_idx", "neighborhood", "mask", "per_patch", "total", "stats", "grad_pred")})
        for n in want[-1]:
            getattr(s, n).fill_(0)
        steps.append(s)
    ring = StepRing(steps).capture()
    for _ in range(3):
        ring.run()
    torch.cuda.synchronize()
    for s, w in zip(steps, want):
        for n, v in w.items():
            assert np.array_equal(host(getattr(s, n)), v), n


def test_step_ring_kernel_sequence_lanes_equal_serial_steps(cuda):
    """Kernel-sequence steps (N > 2048 or many groups) replayed as one graph with the steps spread round-robin over
    forked streams: every step still gives the bits of the same step run alone, and the oracle's grouping."""
    from gm3d_b200.pipeline import GroupLossStep, StepRing
    B, N, G, k = 6, 4096, 96, 32
    rng = np.random.default_rng(19)
    steps, want, clouds = [], [], []
    for r in range(7):
        s = GroupLossStep(B, N, G, k, 0.6, device=cuda, seed=4, rand_offset=r * B * G)
        assert not s.fused
        clouds.append(synthetic_clouds(B, N, 700 + r))
        s.xyz.copy_(dev(clouds[-1], cuda))
        s.loss_pred.copy_(dev(rng.standard_normal((B, G)).astype(np.float32), cuda))
        s.pred.copy_(dev((rng.standard_normal((s.P, k, 3)) * 0.08).astype(np.float32), cuda))
        s.run()
        torch.cuda.synchronize()
        want.append({n: host(getattr(s, n)).copy() for n in ("fps_idx", "neighborhood", "mask", "per_patch", "total", "stats", "grad_pred")})
        for n in want[-1]:
            getattr(s, n).fill_(0)
        steps.append(s)
    ring = StepRing(steps).capture()
    for _ in range(3):
        ring.run()
    torch.cuda.synchronize()
    for s, w in zip(steps, want):
        for n, v in w.items():
            assert np.array_equal(host(getattr(s, n)), v), n
    w0 = co.group(clouds[3], G, k)
    assert np.array_equal(host(steps[3].fps_idx), w0["fps_idx"]) and np.array_equal(host(steps[3].neighborhood), w0["neighborhood"])


def test_step_ring_dataflow_equals_serial_steps(cuda):
    """The dataflow ring (group launches chained by programmatic dependent launch, masks and Chamfer launches on
    forked streams) gives, for every step, the bits of the same step run alone -- also with a deliberately LONG first
    step in front of short ones (4x the clouds would be a different shape; here: the first step's buffers are
    re-used by nobody, and completion must still be transitive along the chain)."""
    from gm3d_b200.pipeline import GroupLossStep, StepRing
    B, N, G, k = 16, 1024, 64, 32
    rng = np.random.default_rng(29)
    steps, want = [], []
    for r in range(7):
        s = GroupLossStep(B, N, G, k, 0.6, device=cuda, seed=4, rand_offset=r * B * G)
        assert s.path == "dataflow" and s.group_per_cloud
        s.xyz.copy_(dev(synthetic_clouds(B, N, 900 + r), cuda))
        s.loss_pred.copy_(dev(rng.standard_normal((B, G)).astype(np.float32), cuda))
        s.pred.copy_(dev((rng.standard_normal((s.P, k, 3)) * 0.08).astype(np.float32), cuda))
        s.run()
        torch.cuda.synchronize()
        want.append({n: host(getattr(s, n)).copy() for n in ("fps_idx", "center", "neighborhood", "mask", "patch_index",
                                                             "dist1", "idx2", "per_patch", "total", "stats", "grad_pred")})
        for n in want[-1]:
            getattr(s, n).fill_(0)
        steps.append(s)
    for schedule in ("lanes", "chain"):  # forked streams (default) / one programmatic-dependent-launch chain G0 M0 G1 C0 M1 ...
        for s, w in zip(steps, want):
            for n in w:
                getattr(s, n).fill_(0)
        ring = StepRing(steps, schedule=schedule).capture()
        for _ in range(3):
            ring.run()
        torch.cuda.synchronize()
        for i, (s, w) in enumerate(zip(steps, want)):
            for n, v in w.items():
                assert np.array_equal(host(getattr(s, n)), v), (schedule, i, n)
            assert np.array_equal(host(ring.head[i, :3]), host(s.stats[:3])) and ring.head[i, 3].item() == 1.0


@pytest.mark.parametrize("path", ["dataflow", "single"])
def test_step_ring_bench_variant_every_slot_vs_oracle(cuda, path):
    """The exact variant bench.py times -- a ring of 24 x (B=128, N=1024, G=64, k=32) steps replayed as ONE graph
    (12-warp per-cloud CTAs chained by programmatic dependent launch, Chamfer launches on a forked stream) -- with
    EVERY slot's fps_idx / neighbourhood / mask / dist / idx / grad_pred against the CPU oracle."""
    import bench
    from gm3d_b200.pipeline import GroupLossStep, StepRing
    B, N, G, k, ratio = bench.CONFIGS["c2"][:5]
    ring_n = 24
    steps, inputs = [], []
    for r in range(ring_n):
        s = GroupLossStep(B, N, G, k, ratio, device=cuda, seed=1234, rand_offset=r * B * G, path=path)
        x, lp, _ = bench.synthetic_batch(B, N, G, k, s.M, 1234 + r)
        s.xyz.copy_(dev(x, cuda)); s.loss_pred.copy_(dev(lp, cuda))
        pred = bench.near_target_pred(s, 99 + r)  # the bench's predictions: masked target + 0.02 * noise
        steps.append(s)
        inputs.append((x, lp, pred))
    ring = StepRing(steps).capture()
    for s in steps:
        for t in (s.fps_idx, s.center, s.neighborhood, s.mask, s.patch_index, s.dist1, s.dist2, s.idx1, s.idx2,
                  s.per_patch, s.total, s.grad_pred, s.stats):
            t.fill_(0)
    ring.run()
    ring.run()
    torch.cuda.synchronize()
    for i, s in enumerate(steps):
        assert bench.oracle_check_step(s, *inputs[i]) == "ok", i
        assert np.array_equal(host(ring.head[i, :3]), host(s.stats[:3]))


def test_m2ae_chain_vs_oracle_chain(cuda):
    """BASELINE config[2] as ONE step: three chained Group levels (level l+1 groups the centres of level l), masks
    and multi-scale Chamfer -- every level against the oracle chain, the GroupM2AE module against the same chain,
    and a ring of steps against the steps run alone."""
    import bench
    from gm3d_b200.group import GroupM2AE
    from gm3d_b200.pipeline import M2AEStep, StepRing
    cfg = (6,) + bench.CONFIGS["c3"][1:]
    B, N, Gs, ks, ratio, _ = cfg
    steps, inputs = [], []
    for r in range(3):
        s = M2AEStep(B, N, Gs, ks, ratio, device=cuda, seed=3, rand_offset=r * B * Gs[0])
        x, lps, preds = bench.m2ae_inputs(cfg, 50 + r)
        s.xyz.copy_(dev(x, cuda))
        for l, lp, pr in zip(s.levels, lps, preds):
            l.loss_pred.copy_(dev(lp, cuda)); l.pred.copy_(dev(pr, cuda))
        steps.append(s)
        inputs.append((x, lps, preds))
    assert steps[0].levels[1].xyz.data_ptr() == steps[0].levels[0].center.data_ptr()  # chained without a copy
    ring = StepRing(steps).capture()
    ring.run(); ring.run()
    torch.cuda.synchronize()
    for s, (x, lps, preds) in zip(steps, inputs):
        cloud = x
        for l, lp, pr in zip(s.levels, lps, preds):
            assert bench.oracle_check_step(l, cloud, lp, pr) == "ok"
            cloud = co.group(cloud, l.G, l.k)["center"]
    nbs, cs, idxs = GroupM2AE(Gs, ks)(steps[0].xyz)
    cloud = inputs[0][0]
    for l, nb, c, idx, g, k in zip(steps[0].levels, nbs, cs, idxs, Gs, ks):
        w = co.group(cloud, g, k)
        assert torch.equal(nb, l.neighborhood) and torch.equal(c, l.center)
        assert np.array_equal(host(idx), w["knn_idx"]) and idx.dtype == torch.int64
        cloud = w["center"]


def test_peer_reduce_two_ranks_on_one_gpu(cuda):
    """The per-step statistics all-reduce over peer memory (gm3d_step_reduce_t), with both 'ranks' played by two
    streams of ONE GPU sharing two local inboxes: each rank's head must hold rank0 + rank1 (summed in rank order, so
    bit-identical on both), over several launches (launch-counter parity), and a missing peer must time out into
    `status` instead of hanging."""
    from gm3d_b200 import _lib
    from gm3d_b200.pipeline import GroupLossStep
    B, N, G, k = 4, 256, 64, 8
    rng = np.random.default_rng(3)
    inbox = torch.zeros((2, _lib.INBOX_BYTES), dtype=torch.uint8, device=cuda)
    epoch = torch.zeros((2,), dtype=torch.int32, device=cuda)
    status = torch.zeros((2,), dtype=torch.int32, device=cuda)
    head = torch.zeros((2, 4), dtype=torch.float32, device=cuda)
    ranks, reds, streams = [], [], [torch.cuda.Stream(cuda), torch.cuda.Stream(cuda)]
    for r in range(2):
        s = GroupLossStep(B, N, G, k, 0.6, device=cuda, seed=r, rand_offset=0)
        s.xyz.copy_(dev(synthetic_clouds(B, N, 40 + r), cuda))
        s.loss_pred.copy_(dev(rng.standard_normal((B, G)).astype(np.float32), cuda))
        s.pred.copy_(dev((rng.standard_normal((s.P, k, 3)) * 0.08).astype(np.float32), cuda))
        red = _lib.StepReduce()
        red.head, red.world, red.rank = head[r].data_ptr(), 2, r
        red.inbox[0], red.inbox[1] = inbox[0].data_ptr(), inbox[1].data_ptr()
        red.epoch, red.timeout_us, red.status = epoch[r:].data_ptr(), 500_000, status[r:].data_ptr()
        ranks.append(s)
        reds.append(red)
    torch.cuda.synchronize()
    for it in range(3):
        for r in (it % 2, 1 - it % 2):  # alternate which rank is enqueued first
            with torch.cuda.stream(streams[r]):
                ranks[r].enqueue(0, reds[r])
        torch.cuda.synchronize()
        a, b = ranks[0].stats[:3], ranks[1].stats[:3]
        want = host(a + b)
        assert np.array_equal(host(head[0, :3]), want) and np.array_equal(host(head[1, :3]), want), it
        assert host(head[:, 3]).tolist() == [2.0, 2.0] and host(status).tolist() == [0, 0]
        assert host(epoch).tolist() == [it + 1, it + 1]
    # a peer that never launches: rank 0 alone must come back (NaN head, status = 1 + missing rank) within the timeout
    with torch.cuda.stream(streams[0]):
        ranks[0].enqueue(0, reds[0])
    torch.cuda.synchronize()
    assert status[0].item() == 2 and np.isnan(host(head[0, :3])).all()


def test_peer_reduce_deferred_collect_on_one_gpu(cuda):
    """Deferred form (defer = 1 + gm3d_step_reduce_collect): the loss launches of three step slots only push; one
    collect launch per 'rank' sums all slots.  Both ranks on one GPU, two streams; two rounds (launch-counter parity)."""
    import ctypes
    from gm3d_b200 import _lib
    from gm3d_b200.pipeline import GroupLossStep
    lib = _lib.load()
    B, N, G, k, n = 4, 256, 64, 8, 3
    rng = np.random.default_rng(13)
    inbox = torch.zeros((2, n, _lib.INBOX_BYTES), dtype=torch.uint8, device=cuda)
    epoch = torch.zeros((2, n), dtype=torch.int32, device=cuda)
    status = torch.zeros((2,), dtype=torch.int32, device=cuda)
    head = torch.zeros((2, n, 4), dtype=torch.float32, device=cuda)
    streams = [torch.cuda.Stream(cuda), torch.cuda.Stream(cuda)]
    steps, reds = [[], []], [[], []]
    for r in range(2):
        for i in range(n):
            s = GroupLossStep(B, N, G, k, 0.6, device=cuda, seed=r, rand_offset=i * B * G)
            s.xyz.copy_(dev(synthetic_clouds(B, N, 70 + 10 * r + i), cuda))
            s.loss_pred.copy_(dev(rng.standard_normal((B, G)).astype(np.float32), cuda))
            s.pred.copy_(dev((rng.standard_normal((s.P, k, 3)) * 0.08).astype(np.float32), cuda))
            red = _lib.StepReduce()
            red.head, red.world, red.rank, red.defer = head[r, i].data_ptr(), 2, r, 1
            red.inbox[0], red.inbox[1] = inbox[0, i].data_ptr(), inbox[1, i].data_ptr()
            red.epoch, red.timeout_us, red.status = epoch[r, i:].data_ptr(), 500_000, status[r:].data_ptr()
            steps[r].append(s)
            reds[r].append(red)
    torch.cuda.synchronize()
    for it in range(2):
        head.zero_()
        for r in (0, 1):
            with torch.cuda.stream(streams[r]):
                for i in range(n):
                    steps[r][i].enqueue(0, reds[r][i])
                assert lib.gm3d_step_reduce_collect(ctypes.byref(reds[r][0]), n, streams[r].cuda_stream) == 0
        torch.cuda.synchronize()
        for i in range(n):
            want = host(steps[0][i].stats[:3] + steps[1][i].stats[:3])
            assert np.array_equal(host(head[0, i, :3]), want) and np.array_equal(host(head[1, i, :3]), want), (it, i)
        assert (head[:, :, 3] == 2.0).all() and host(status).tolist() == [0, 0] and (epoch == it + 1).all()
    # lagging form: `collected` starts at 0 while two launches have been pushed -- every call sums the next launch count
    # (older pushes are still in the 4-deep inbox), a call with nothing outstanding leaves head alone
    collected = torch.zeros((2, n), dtype=torch.int32, device=cuda)
    for r in range(2):
        reds[r][0].collected = collected[r].data_ptr()
    for call in range(3):
        head.zero_()
        for r in (0, 1):
            assert lib.gm3d_step_reduce_collect(ctypes.byref(reds[r][0]), n, streams[r].cuda_stream) == 0
        torch.cuda.synchronize()
        assert (collected == min(call + 1, 2)).all()
        if call < 2:
            for i in range(n):
                want = host(steps[0][i].stats[:3] + steps[1][i].stats[:3])
                assert np.array_equal(host(head[0, i, :3]), want) and np.array_equal(host(head[1, i, :3]), want)
        else:
            assert (head == 0).all()


# ------------------------------------------------------------------------------------------ SURVEY 8(f) rows
@pytest.fixture(scope="module")
def golden_next():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_next.npz"))


@pytest.mark.parametrize("tag", ["c2", "small", "m2ae"])
@pytest.mark.parametrize("rel", [True, False])
def test_learning_loss_matches_reference_golden_and_oracle(cuda, golden_next, tag, rel):
    """forward_learning_loss value and autograd gradient against the reference's own function (golden) and the
    float64 oracle; tolerance 1e-5 relative on the loss, 1e-4 of the largest gradient entry (fp32 exp/log)."""
    from gm3d_b200.loss import forward_learning_loss
    p = dev(golden_next[f"ll_{tag}_pred"], cuda).requires_grad_(True)
    t = dev(golden_next[f"ll_{tag}_target"], cuda)
    loss = forward_learning_loss(p.unsqueeze(-1) if tag == "small" else p, None, t, relative=rel)
    (3.0 * loss).backward()
    want = float(golden_next[f"ll_{tag}_rel{int(rel)}_loss"])
    assert abs(loss.item() - want) <= 1e-5 * abs(want)
    wg = 3.0 * golden_next[f"ll_{tag}_rel{int(rel)}_grad"]
    assert np.abs(host(p.grad) - wg).max() <= 1e-4 * np.abs(wg).max()
    ol, og = no.learning_loss(golden_next[f"ll_{tag}_pred"], golden_next[f"ll_{tag}_target"], rel)
    assert abs(loss.item() - ol) <= 1e-5 * abs(ol)
    assert np.abs(host(p.grad) - 3.0 * og).max() <= 1e-4 * np.abs(og).max() * 3.0


def test_learning_loss_full_size_and_determinism(cuda):
    from gm3d_b200 import ops
    rng = np.random.default_rng(3)
    p = rng.standard_normal((128, 39)).astype(np.float32)
    t = (rng.random((128, 39)) * 0.05).astype(np.float32)
    a = ops.learning_loss(dev(p, cuda), dev(t, cuda), True)
    b = ops.learning_loss(dev(p, cuda), dev(t, cuda), True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])  # fixed summation order
    ol, og = no.learning_loss(p, t, True)
    assert abs(a[0].item() - ol) <= 1e-5 * abs(ol)
    assert np.abs(host(a[1]) - og).max() <= 1e-4 * np.abs(og).max()
    # antisymmetry of the pairwise logits: a constant shift of the predictions changes nothing
    c = ops.learning_loss(dev(p + 0.5, cuda), dev(t, cuda), True)
    assert abs(c[0].item() - a[0].item()) <= 1e-5 * abs(a[0].item())
    assert abs(host(a[1]).sum()) <= 1e-4 * np.abs(host(a[1])).sum()


def test_scale_and_translate_dropin(cuda, golden_next):
    """Seeded like the reference run that produced the golden: same NumPy stream, bit-identical clouds."""
    from gm3d_b200.transforms import PointcloudScaleAndTranslate
    pc = dev(golden_next["sat_in"], cuda)
    np.random.seed(2003)
    out = PointcloudScaleAndTranslate()(pc)
    assert out.data_ptr() == pc.data_ptr()  # in place, like the reference
    assert np.array_equal(host(out), golden_next["sat_out"])
    # extra channels (normals) stay untouched
    from gm3d_b200 import ops
    x6 = np.random.default_rng(1).standard_normal((3, 50, 6)).astype(np.float32)
    ss = np.random.default_rng(2).uniform(0.5, 1.5, (3, 6)).astype(np.float32)
    got = host(ops.scale_translate_(dev(x6, cuda), dev(ss, cuda)))
    assert np.array_equal(got[:, :, :3], no.scale_translate(x6[:, :, :3], ss)) and np.array_equal(got[:, :, 3:], x6[:, :, 3:])


def test_fps_subsample_matches_reference_golden(cuda, golden_next):
    from gm3d_b200.pointnet2_utils import fps_subsample
    pts = dev(golden_next["ft_points"], cuda)
    got = fps_subsample(pts, 1024, choice=golden_next["ft_choice"])
    assert np.array_equal(host(got), golden_next["ft_out"])
    np.random.seed(2007)  # the same draw the reference made
    assert np.array_equal(host(fps_subsample(pts, 1024)), golden_next["ft_out"])
    # G == N edge (engine_finetune.py:129-130): point_all capped at N
    small = dev(synthetic_clouds(2, 600, 9), cuda)
    np.random.seed(1)
    ch = np.random.choice(600, 512, False)
    assert np.array_equal(host(fps_subsample(small, 512, point_all=1200, choice=ch)),
                          no.gather_points(host(small), co.fps(host(small), 600), ch))


def _encoder_sd(golden_next):
    from test_oracle import _encoder_state_dict
    return {k: torch.from_numpy(v) for k, v in _encoder_state_dict(golden_next).items()}


def test_encoder_tcgen05_matches_reference_golden(cuda, golden_next):
    """tcgen05 Encoder (BF16 operands, FP32 accumulate) against the reference module's own FP32 output: the
    tolerance is BF16's -- 1e-2 of the output range at worst, 2e-3 on average."""
    from gm3d_b200.encoder import EncoderB200
    enc = EncoderB200.from_state_dict(_encoder_sd(golden_next)).to(cuda)
    got = host(enc(dev(golden_next["enc_neighborhood"], cuda)))
    assert enc.last_status.item() == 0
    want = golden_next["enc_out"]
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-2 * scale
    assert np.abs(got - want).mean() <= 2e-3 * scale


def test_encoder_tcgen05_full_size_vs_oracle_and_properties(cuda, golden_next):
    """Pre-training size (B=128, G=64: 8192 patches, 64 tiles per SM-slot), ragged patch count, determinism, and
    permutation invariance over the points of a patch (both max-pools)."""
    from gm3d_b200.encoder import EncoderB200
    sd = _encoder_sd(golden_next)
    enc = EncoderB200.from_state_dict(sd).to(cuda)
    rng = np.random.default_rng(12)
    nb = (rng.standard_normal((128, 64, 32, 3)) * 0.08).astype(np.float32)
    a = enc(dev(nb, cuda))
    b = enc(dev(nb, cuda))
    assert enc.last_status.item() == 0 and torch.equal(a, b)
    sub = no.encoder_eval(nb[:2], {k: v.numpy() for k, v in sd.items()})
    assert np.abs(host(a)[:2] - sub).max() <= 1e-2 * np.abs(sub).max()
    perm = rng.permutation(32)
    c = enc(dev(nb[:, :, perm], cuda))
    assert np.abs(host(c) - host(a)).max() <= 1e-2 * np.abs(host(a)).max()   # BF16 rounding differs per row order only
    odd = enc(dev(nb[:1, :7], cuda))                                          # 7 patches: a partial tile
    assert np.array_equal(host(odd), host(a)[:1, :7])
    with pytest.raises(NotImplementedError):
        enc(dev(nb[:1, :4, :16], cuda))                                       # n != 32
