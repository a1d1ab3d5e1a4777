"""The two-phase kNN kernel (gm3d_b200/csrc/knn_large.cuh) screens points with t = |p|^2 - 2 p.q (three chained FP32
FMAs) and relies on |t_fp + |q|^2_fp - d_fp| <= E = 24 * 2^-24 * (max|p| + |q|)^2, where d_fp is the reference FP32
distance (KNN_CUDA's `ssd += t*t` chain).  This CPU test restates both expressions with the oracle's exact FP32 FMA
emulation and checks the bound -- and how much headroom it has -- over scales, offsets and the cancellation-heavy case
of queries that coincide with far-away points."""
import numpy as np
import pytest

from oracle import np_oracle as no

U = np.float64(2.0) ** -24


def screen_and_reference(p, q):
    """p (n,3), q (m,3) float32 -> t_fp + cq_fp and d_fp as float64 (m, n), computed with FP32 roundings."""
    x, y, z = (p[:, i][None, :] for i in range(3))
    qx, qy, qz = (q[:, i][:, None] for i in range(3))
    w = no.sumsq_acc(p[:, 0], p[:, 1], p[:, 2])[None, :]                      # |p|^2, as the kernel's cloud load
    two = np.float32(-2.0)
    t = no.fma32(x, two * qx, no.fma32(y, two * qy, no.fma32(z, two * qz, np.broadcast_to(w, (q.shape[0], p.shape[0])))))
    cq = no.sumsq_acc(q[:, 0], q[:, 1], q[:, 2])[:, None]
    d = no.sumsq_acc((x - qx).astype(np.float32), (y - qy).astype(np.float32), (z - qz).astype(np.float32))
    return t.astype(np.float64) + cq.astype(np.float64), d.astype(np.float64), cq


@pytest.mark.parametrize("shift,scale", [(0.0, 1.0), (3.0, 1.0), (30.0, 1.0), (1000.0, 1.0), (0.0, 1e-12), (0.0, 1e9), (5e4, 50.0)])
def test_screen_error_within_margin(shift, scale):
    rng = np.random.default_rng(17)
    p = (rng.standard_normal((4096, 3)) * scale + shift).astype(np.float32)
    q = np.concatenate([p[:64], (p[64:128] + rng.standard_normal((64, 3)).astype(np.float32) * np.float32(0.03 * scale))]).astype(np.float32)
    s, d, cq = screen_and_reference(p, q)
    r = np.sqrt(no.sumsq_acc(p[:, 0], p[:, 1], p[:, 2]).max().astype(np.float64)) + np.sqrt(cq.astype(np.float64))
    E = np.maximum(24.0 * U * r * r, 1e-30)                                   # per query, as in the kernel
    err = np.abs(s - d)
    assert (err <= E).all(), float((err / E).max())
    assert (err / E).max() < 0.6                                               # derivation: 14.2 u of the 24 u margin


def test_screen_selects_a_superset_of_the_k_nearest():
    """The decisions the kernel takes on t: with Tt = any value such that k points have t <= Tt, every one of the k
    nearest by the reference expression has t <= Tt + 2E and reference distance <= Tt + |q|^2 + 1.5E."""
    rng = np.random.default_rng(23)
    k = 32
    for shift in (0.0, 2.0, 40.0):
        p = (rng.standard_normal((2048, 3)) * 0.5 + shift).astype(np.float32)
        q = p[rng.choice(2048, 32, replace=False)]
        s, d, cq = screen_and_reference(p, q)
        t = s - cq.astype(np.float64)
        r = np.sqrt(no.sumsq_acc(p[:, 0], p[:, 1], p[:, 2]).max().astype(np.float64)) + np.sqrt(cq.astype(np.float64))
        E = 24.0 * U * r * r
        Tt = np.sort(t, axis=1)[:, k - 1:k]                                  # the tightest admissible bound
        kth = np.sort(d, axis=1)[:, k - 1:k]
        near = d <= kth                                                      # the k nearest (with ties)
        assert (t[near.nonzero()] <= (Tt + 2 * E + np.zeros_like(t))[near.nonzero()]).all()
        assert (kth <= Tt + cq + 1.5 * E).all()
