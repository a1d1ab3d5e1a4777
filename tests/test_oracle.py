"""CPU tests of the oracle itself: the C restatement against the independent NumPy restatement, and both
against the golden vectors produced by the reference's own glue code (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import synthetic_clouds
from oracle import c_oracle as co
from oracle import np_oracle as no


def test_fma_emulation_matches_libm():
    rng = np.random.default_rng(0)
    n = 400_000
    a = rng.standard_normal(n).astype(np.float32)
    b = rng.standard_normal(n).astype(np.float32)
    # c close to -a*b: heavy cancellation, the case where double rounding would show
    c = (-(a.astype(np.float64) * b)).astype(np.float32) + (rng.standard_normal(n) * 1e-7).astype(np.float32)
    assert np.array_equal(co.fmaf(a, b, c).view(np.uint32), no.fma32(a, b, c).view(np.uint32))
    # halfway cases: a*b + c exactly between two floats
    a = np.full(1000, 1 + 2.0**-12, dtype=np.float32)
    b = np.full(1000, 1 + 2.0**-12, dtype=np.float32)
    c = (rng.integers(-4, 4, 1000) * 2.0**-24).astype(np.float32)
    assert np.array_equal(co.fmaf(a, b, c).view(np.uint32), no.fma32(a, b, c).view(np.uint32))


@pytest.mark.parametrize("kind", ["ball", "sphere"])
@pytest.mark.parametrize("B,N,G", [(3, 256, 32), (2, 333, 17), (2, 64, 64)])
def test_fps_c_vs_numpy(kind, B, N, G):
    xyz = synthetic_clouds(B, N, 10 + N, kind)
    assert np.array_equal(co.fps(xyz, G), no.fps(xyz, G))


def test_fps_properties():
    xyz = synthetic_clouds(2, 200, 3, "ball")
    idx = co.fps(xyz, 200)  # G == N (engine_finetune.py:129-132)
    assert (idx[:, 0] == 0).all()
    # every eligible point is selected exactly once before duplicates start
    mag = (xyz.astype(np.float64) ** 2).sum(-1)
    for b in range(2):
        elig = np.flatnonzero(mag[b] > 1.1e-3)
        first = idx[b, : len(elig)]
        assert len(set(first.tolist())) == len(first)


def test_fps_skip_rule_and_all_skipped():
    xyz = np.zeros((1, 16, 3), dtype=np.float32)
    xyz[0, :, 0] = np.linspace(0.0, 0.02, 16)  # all |p|^2 <= 4e-4 < 1e-3  => nothing eligible
    assert (co.fps(xyz, 5) == 0).all()
    xyz[0, 7] = [1.0, 0, 0]
    xyz[0, 9] = [-1.0, 0, 0]
    idx = co.fps(xyz, 4)[0]
    assert idx.tolist() == [0, 7, 9, 7]  # only 7 and 9 are eligible; once both are at distance 0 -> lowest index
    assert np.array_equal(no.fps(xyz, 4)[0], idx)
    # without the rule the near-origin points take part
    assert co.fps(xyz, 4, skip_near_origin=False)[0].tolist() != idx.tolist()


def test_fps_skip_threshold_is_a_double_compare():
    # float(1e-3) > 1e-3 (double): a point with |p|^2 == float32(1e-3) is NOT skipped upstream
    t = np.float32(1e-3)
    assert float(t) > 1e-3
    x = np.sqrt(np.float64(t))
    xyz = np.zeros((1, 3, 3), dtype=np.float32)
    xyz[0, 1] = [2, 0, 0]
    for cand in np.nextafter(np.float32(x), np.float32([0, 1])).tolist() + [np.float32(x)]:
        xyz[0, 2] = [cand, 0, 0]
        mag = no.sumsq_nvcc(xyz[0, 2, 0], xyz[0, 2, 1], xyz[0, 2, 2])
        skipped = float(mag) <= 1e-3
        idx = co.fps(xyz, 3)[0]
        # point 0 is the origin (skipped, but it is the fixed start); 1 is picked second; third is 2 iff eligible
        assert idx[1] == 1
        assert (idx[2] == 2) == (not skipped)


def test_fps_thread_order_tie_mode():
    # 4 points at the same distance from point 0: contract picks the lowest index, upstream's block of 2
    # threads picks the lowest (k mod 2) first
    xyz = np.array([[[0, 0, 0.5], [1, 0, 0.5], [0, 1, 0.5], [-1, 0, 0.5], [0, -1, 0.5]]], dtype=np.float32)
    assert co.fps(xyz, 2, tie="lowest_index")[0, 1] == 1
    assert co.fps(xyz, 2, tie="pointnet2_thread_order", block=2)[0, 1] == 2


def test_fps_tie_rules_agree_on_the_baseline_inputs():
    """The contract resolves equal running-min distances to the lowest point index; upstream pointnet2 resolves them
    by its block-reduction thread order.  On the BASELINE inputs (bench.synthetic_batch, the seeds
    bench.py uses, continuous coordinates) no round ever sees a tie, so the two rules select identical indices and
    the contract choice is immaterial there."""
    import bench
    for name in ("c1", "c2", "c3l0", "c3l1", "c3l2", "c4", "c5"):
        B, N, G, k, ratio, _ = bench.CONFIGS[name]
        Bs = min(B, 16)  # the generator draws cloud by cloud in order: the first clouds of the bench batch
        x, _, _ = bench.synthetic_batch(B, N, G, 1, 1, 1234)
        x = x[:Bs]
        a = co.fps(x, G, tie="lowest_index")
        b = co.fps(x, G, tie="pointnet2_thread_order")
        assert int((a != b).sum()) == 0, name


def test_fps_tie_rules_diverge_only_at_exact_ties():
    """With exact duplicate points the two rules may pick different indices -- but always points of bit-equal
    coordinates or bit-equal running-min distance in that round; the divergence is counted, not hidden."""
    x = synthetic_clouds(8, 1024, 5, "sphere")  # ~1 % exact duplicates
    G = 256
    a = co.fps(x, G, tie="lowest_index")
    b = co.fps(x, G, tie="pointnet2_thread_order")
    diff = np.argwhere(a != b)
    for bi in np.unique(diff[:, 0]):
        j = diff[diff[:, 0] == bi][:, 1].min()          # first round where the rules part
        pa, pb = x[bi, a[bi, j]], x[bi, b[bi, j]]
        sel = x[bi, a[bi, :j]]                           # centres chosen so far (identical under both rules)
        da = ((sel - pa) ** 2).sum(-1).min()
        db = ((sel - pb) ** 2).sum(-1).min()
        assert np.isclose(da, db, rtol=1e-6), (bi, j)    # same running-min distance (to rounding): a genuine tie
    # lowest_index never picks a higher index than thread order could justify: indices are valid either way
    assert a.min() >= 0 and a.max() < 1024 and b.min() >= 0 and b.max() < 1024


def test_numpy_fps_pin(golden):
    """The in-tree CPU FPS (datasets/ModelNetDataset.py:25-46) selects the points our restatement selects."""
    pts = golden["npfps_points"]
    sel = no.fps_numpy_reference(pts, 48, int(golden["npfps_start"][0]))
    assert np.array_equal(pts[sel], golden["npfps_selected"])


@pytest.mark.parametrize("kind", ["ball", "sphere"])
def test_knn_c_vs_numpy(kind):
    xyz = synthetic_clouds(2, 300, 5, kind)
    q = xyz[:, ::7].copy()
    D1, I1 = co.knn(xyz, q, 16)
    D2, I2 = no.knn(xyz, q, 16)
    assert np.array_equal(I1, I2) and np.array_equal(D1.view(np.uint32), D2.view(np.uint32))
    assert (I1[:, :, 0] == np.arange(0, 300, 7)[None]).all() or kind == "sphere"  # the query itself is nearest
    assert (np.diff(D1, axis=-1) >= 0).all()


def test_knn_ties_keep_lower_index():
    ref = np.zeros((1, 6, 3), dtype=np.float32)
    ref[0, :, 0] = [1, 1, 2, 1, 0.5, 2]
    q = np.zeros((1, 1, 3), dtype=np.float32)
    _, I = co.knn(ref, q, 5)
    assert I[0, 0].tolist() == [4, 0, 1, 3, 2]
    with pytest.raises(ValueError):
        co.knn(ref, q, 7)


@pytest.mark.parametrize("tag", ["c1", "m2ae_l2", "ragged", "g_eq_n"])
def test_group_matches_reference_glue(golden, tag):
    """oracle.group == the reference's Group.forward (both variants) run on the oracle's operators."""
    xyz = golden[f"group_{tag}_xyz"]
    G, k = golden[f"group_{tag}_G_k"].tolist()
    for impl in (co, no):
        r = impl.group(xyz, G, k)
        assert np.array_equal(r["center"], golden[f"group_{tag}_center"])
        assert np.array_equal(r["neighborhood"], golden[f"group_{tag}_neighborhood"])
        assert np.array_equal(r["neighborhood_org"], golden[f"group_{tag}_neighborhood_org"])


def test_chamfer_c_vs_numpy_and_bwd():
    rng = np.random.default_rng(3)
    for (P, n, m) in [(40, 32, 32), (7, 16, 16), (5, 8, 8), (3, 20, 45), (2, 100, 70)]:
        a = rng.standard_normal((P, n, 3)).astype(np.float32)
        b = rng.standard_normal((P, m, 3)).astype(np.float32)
        if n == m:
            b = (a + 0.02 * rng.standard_normal((P, n, 3))).astype(np.float32)
        r1, r2 = co.chamfer_fwd(a, b), no.chamfer_fwd(a, b)
        for x, y in zip(r1, r2):
            assert np.array_equal(x, y)
        g1 = rng.standard_normal((P, n)).astype(np.float32)
        g2 = rng.standard_normal((P, m)).astype(np.float32)
        ga, gb = co.chamfer_bwd(a, b, r1[2], r1[3], g1, g2)
        ga64, gb64 = no.chamfer_bwd(a, b, r1[2], r1[3], g1, g2)
        assert np.allclose(ga, ga64, rtol=1e-5, atol=1e-6) and np.allclose(gb, gb64, rtol=1e-5, atol=1e-6)


def test_chamfer_bwd_is_the_gradient():
    """finite differences of sum(w1*dist1)+sum(w2*dist2) in float64 agree with the restated backward."""
    rng = np.random.default_rng(4)
    a = rng.standard_normal((2, 6, 3)).astype(np.float32)
    b = rng.standard_normal((2, 5, 3)).astype(np.float32)
    w1 = rng.standard_normal((2, 6)).astype(np.float32)
    w2 = rng.standard_normal((2, 5)).astype(np.float32)

    def f(a_, b_):
        d = ((a_[:, :, None, :].astype(np.float64) - b_[:, None, :, :]) ** 2).sum(-1)
        return (w1 * d.min(2)).sum() + (w2 * d.min(1)).sum()

    _, _, i1, i2 = co.chamfer_fwd(a, b)
    ga, gb = no.chamfer_bwd(a, b, i1, i2, w1, w2)
    eps = 1e-4
    for arr, g in ((a, ga), (b, gb)):
        for ix in np.ndindex(arr.shape):
            hi, lo = arr.astype(np.float64).copy(), arr.astype(np.float64).copy()
            hi[ix] += eps
            lo[ix] -= eps
            num = (f(hi, b) - f(lo, b)) / (2 * eps) if arr is a else (f(a, hi) - f(a, lo)) / (2 * eps)
            assert abs(num - g[ix]) < 1e-4 * max(1, abs(num))


def test_forward_loss_glue_pin(golden):
    """reference forward_loss (usual mode) == oracle restatement, for both candidate per-point definitions."""
    nb, mask, pred = golden["loss_neighborhood"], golden["loss_mask"], golden["loss_pred_points"]
    for mode in ("dist1", "sum"):
        r = no.forward_loss_usual(pred, nb, mask, per_point=mode)
        assert np.allclose(r["matrix"], golden[f"loss_usual_{mode}_matrix"], rtol=1e-5, atol=1e-8)
        assert np.isclose(r["Chamfer_mean"], golden[f"loss_usual_{mode}_chamfer_mean"], rtol=1e-5)
    # stock scalars
    gt = nb[mask].reshape(-1, 32, 3)
    pr = pred.reshape(-1, 32, 3)
    assert np.isclose(no.chamfer_l2(pr, gt), golden["cdl2_scalar"], rtol=1e-5)
    assert np.isclose(no.chamfer_l1(pr, gt), golden["cdl1_scalar"], rtol=1e-5)
    # per-patch reduction of the C oracle: 'patch' == matrix under the 'sum' definition (n == m)
    d1, d2, _, _ = co.chamfer_fwd(pr, gt)
    pp = co.chamfer_per_patch(d1, d2, 2).reshape(4, 39)
    assert np.allclose(pp, golden["loss_usual_sum_matrix"], rtol=1e-5, atol=1e-8)


def _case_params(key):
    # mask_{cls}_e{epoch}_t{total}_a{after}
    _, cls, e, t, a = key.split("_")
    return cls, int(e[1:]), int(t[1:]), bool(int(a[1:]))


def test_generate_mask_pin(golden):
    """Against the reference's generate_mask outputs: exact cardinality, exact top-len_loss membership, and
    exact equality once the reference's random choice is replayed through rand_keys."""
    lp = golden["mask_loss_pred"]
    B, L = lp.shape
    for key in golden["mask_cases"].tolist():
        cls, epoch, total, after = _case_params(key)
        ref = golden[key]
        len_keep, len_loss = no.mask_lengths(L, 0.6, epoch, total, True, after or None, 0.8 if cls == "fb" else 0.5)
        assert (ref.sum(1) == L - len_keep).all(), key
        order = np.argsort(lp, axis=1, kind="stable")
        top = np.zeros((B, L), dtype=bool)
        if len_loss:
            np.put_along_axis(top, order[:, L - len_loss:], True, axis=1)
        assert (ref[top] == 1).all(), key
        # replay: keys = 1 where the reference masked a non-top patch
        keys = ((ref == 1) & ~top).astype(np.float32)
        for impl in (co, no):
            assert np.array_equal(impl.hard_mask(lp, len_keep, len_loss, keys), ref.astype(np.uint8)), key


def test_rand_mask_pin(golden):
    ref = golden["rand_mask_c1"]
    num_mask = int(golden["rand_mask_c1_num_mask"][0])
    assert num_mask == int(0.6 * 64) and (ref.sum(1) == num_mask).all()
    assert np.array_equal(no.rand_mask(ref.astype(np.float32), num_mask), ref.astype(np.uint8))
    assert np.array_equal(co.hard_mask(np.zeros_like(ref, dtype=np.float32), 64 - num_mask, 0, ref.astype(np.float32)),
                          ref.astype(np.uint8))


def test_block_mask_pin():
    """`_mask_center_block` restated in NumPy against the reference method's own output (reference_block_mask.npz,
    tests/golden/make_golden_block_mask.py), with the reference's `random.randint` picks."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_block_mask.npz"))
    for tag in g["cases"]:
        m = no.mask_center_block(g[f"{tag}_centers"], float(g[f"{tag}_ratio"][0]), g[f"{tag}_picks"])
        assert np.array_equal(m, g[f"{tag}_mask"]), tag
        assert (m.sum(1) == int(float(g[f"{tag}_ratio"][0]) * m.shape[1])).all()


def test_hard_mask_c_vs_numpy_with_ties():
    rng = np.random.default_rng(9)
    lp = rng.integers(0, 6, (12, 64)).astype(np.float32)  # many ties
    rk = rng.integers(0, 4, (12, 64)).astype(np.float32)
    for len_keep, len_loss in [(25, 15), (25, 0), (25, 39), (0, 10), (64, 0), (13, 1)]:
        m1, m2 = co.hard_mask(lp, len_keep, len_loss, rk), no.hard_mask(lp, len_keep, len_loss, rk)
        assert np.array_equal(m1, m2)
        assert (m1.sum(1) == 64 - len_keep).all()
    with pytest.raises(ValueError):
        co.hard_mask(lp, 25, 40, rk)


def test_gather_and_grad():
    rng = np.random.default_rng(2)
    f = rng.standard_normal((2, 3, 50)).astype(np.float32)
    idx = rng.integers(0, 50, (2, 20)).astype(np.int32)
    idx[0, :4] = 7  # duplicates accumulate in the backward
    assert np.array_equal(co.gather(f, idx), no.gather(f, idx))
    go = rng.standard_normal((2, 3, 20)).astype(np.float32)
    assert np.array_equal(co.gather_grad(go, idx, 50), no.gather_grad(go, idx, 50))


# ------------------------------------------------------------------ SURVEY 8(f) rows vs the reference's own functions
@pytest.fixture(scope="module")
def golden_next():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_next.npz"))


@pytest.mark.parametrize("tag", ["c2", "small", "m2ae"])
@pytest.mark.parametrize("rel", [1, 0])
def test_learning_loss_oracle_matches_reference(golden_next, tag, rel):
    from oracle import np_oracle as no
    loss, grad = no.learning_loss(golden_next[f"ll_{tag}_pred"], golden_next[f"ll_{tag}_target"], bool(rel))
    assert abs(loss - float(golden_next[f"ll_{tag}_rel{rel}_loss"])) <= 2e-6 * abs(loss)
    want = golden_next[f"ll_{tag}_rel{rel}_grad"]
    assert np.abs(grad - want).max() <= 2e-5 * np.abs(want).max()


def test_scale_translate_and_subsample_oracle_match_reference(golden_next):
    from oracle import c_oracle as co
    from oracle import np_oracle as no
    out = no.scale_translate(golden_next["sat_in"], golden_next["sat_draws"].astype(np.float32))
    assert np.array_equal(out, golden_next["sat_out"])
    pts, choice = golden_next["ft_points"], golden_next["ft_choice"]
    idx = co.fps(pts, int(golden_next["ft_point_all"][0]))
    assert np.array_equal(no.gather_points(pts, idx, choice), golden_next["ft_out"])


def _encoder_state_dict(golden_next):
    sd = {}
    for k in golden_next.files:
        if k.startswith("enc_w_"):
            a, i, name = k[len("enc_w_"):].split("_", 2)[0:3] if not k[len("enc_w_"):].startswith(("first_conv", "second_conv")) else (None, None, None)
            rest = k[len("enc_w_"):]
            for pre in ("first_conv", "second_conv"):
                if rest.startswith(pre + "_"):
                    idx, name = rest[len(pre) + 1:].split("_", 1)
                    sd[f"{pre}.{idx}.{name}"] = golden_next[k]
    return sd


def test_encoder_oracle_matches_reference(golden_next):
    from oracle import np_oracle as no
    out = no.encoder_eval(golden_next["enc_neighborhood"], _encoder_state_dict(golden_next))
    want = golden_next["enc_out"]
    assert np.abs(out - want).max() <= 2e-5 * np.abs(want).max()
