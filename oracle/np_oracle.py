"""NumPy restatement of the GM3D grouping + reconstruction-loss path (second, independent oracle).

TEST INFRASTRUCTURE ONLY (see oracle/gm3d_oracle.c header).  PARITY UNPINNED for the operator
arithmetic -- pointnet2_ops / KNN_CUDA / chamfer_dist are absent from /root/reference; the glue
functions below (group, forward_loss_*, generate_mask, mask_center_rand, fps_numpy_reference) follow
in-tree reference code and are pinned by tests/golden/.

It is written differently from the C oracle on purpose (vectorised, FMA emulated in float64 with
round-to-odd, stable sorts instead of insertion loops) so that a shared bug is unlikely; the two are
cross-checked bit-for-bit in tests/test_oracle.py.  Paths cited are relative to
/root/reference/Point-MAE_SA3D.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# ---------------------------------------------------------------------------------------------
# exact fp32 FMA without libm: p = a*b is exact in float64 (24+24 <= 53 bits); s = RN64(p + c) with
# its rounding error e from TwoSum; if inexact, force the float64 result to an odd last bit
# (round-to-odd) so that the final RN to 24 bits rounds once.
# ---------------------------------------------------------------------------------------------
def fma32(a, b, c) -> np.ndarray:
    a64 = np.asarray(a, dtype=F32).astype(np.float64)
    b64 = np.asarray(b, dtype=F32).astype(np.float64)
    c64 = np.asarray(c, dtype=F32).astype(np.float64)
    p = a64 * b64
    s = p + c64
    bb = s - p
    e = (p - (s - bb)) + (c64 - bb)  # TwoSum: s + e == p + c exactly
    bits = s.view(np.int64) if s.ndim else np.array(s).view(np.int64)
    even = (bits & 1) == 0
    inexact = e != 0
    toward = np.where(e > 0, np.inf, -np.inf)
    s_odd = np.where(inexact & even, np.nextafter(s, toward), s)
    return s_odd.astype(F32)


def sumsq_nvcc(a, b, c) -> np.ndarray:
    """`a*a + b*b + c*c` as nvcc contracts it: fma(c,c, fma(a,a, b*b))  (pointnet2_ops, chamfer)."""
    a, b, c = (np.asarray(v, dtype=F32) for v in (a, b, c))
    return fma32(c, c, fma32(a, a, (b * b).astype(F32)))


def sumsq_acc(a, b, c) -> np.ndarray:
    """KNN_CUDA: ssd = 0; ssd += t*t per dim -> fma(c,c, fma(b,b, a*a))."""
    a, b, c = (np.asarray(v, dtype=F32) for v in (a, b, c))
    return fma32(c, c, fma32(b, b, (a * a).astype(F32)))


# ---------------------------------------------------------------------------------------------
# a1 furthest_point_sample (utils/miscc.py:18)
# ---------------------------------------------------------------------------------------------
def fps(xyz, G: int, skip_near_origin: bool = True) -> np.ndarray:
    xyz = np.asarray(xyz, dtype=F32)
    B, N, _ = xyz.shape
    out = np.zeros((B, G), dtype=np.int32)
    for b in range(B):
        p = xyz[b]
        mag = sumsq_nvcc(p[:, 0], p[:, 1], p[:, 2])
        eligible = ~(mag.astype(np.float64) <= 1e-3) if skip_near_origin else np.ones(N, dtype=bool)
        temp = np.full(N, 1e10, dtype=F32)
        old = 0
        for j in range(1, G):
            d = sumsq_nvcc(p[:, 0] - p[old, 0], p[:, 1] - p[old, 1], p[:, 2] - p[old, 2])
            temp = np.where(eligible, np.minimum(d, temp), temp)
            if eligible.any():
                cand = np.where(eligible, temp, F32(-1.0))
                old = int(np.argmax(cand))  # first maximum = lowest index
            else:
                old = 0
            out[b, j] = old
    return out


def fps_numpy_reference(point: np.ndarray, npoint: int, start: int) -> np.ndarray:
    """In-tree CPU FPS, datasets/ModelNetDataset.py:25-46, with the random start made explicit.
    float64, no skip rule; returns the selected INDICES (the reference returns point[indices])."""
    xyz = point[:, :3]
    N = xyz.shape[0]
    centroids = np.zeros((npoint,), dtype=np.int64)
    distance = np.ones((N,)) * 1e10
    farthest = start
    for i in range(npoint):
        centroids[i] = farthest
        dist = np.sum((xyz - xyz[farthest, :]) ** 2, -1)
        distance = np.minimum(distance, dist)
        farthest = int(np.argmax(distance, -1))
    return centroids


# ---------------------------------------------------------------------------------------------
# a2 gather_operation (+ backward)
# ---------------------------------------------------------------------------------------------
def gather(features, idx) -> np.ndarray:
    features = np.asarray(features, dtype=F32)
    idx = np.asarray(idx)
    return np.take_along_axis(features, idx[:, None, :].astype(np.int64).repeat(features.shape[1], 1), axis=2)


def gather_grad(gout, idx, N: int) -> np.ndarray:
    gout = np.asarray(gout, dtype=F32)
    B, C, G = gout.shape
    g = np.zeros((B, C, N), dtype=F32)
    for b in range(B):
        for j in range(G):  # j ascending, duplicates accumulate
            g[b, :, idx[b, j]] += gout[b, :, j]
    return g


def fps_centers(xyz, G: int) -> np.ndarray:
    """miscc.fps (utils/miscc.py:13-20): fps idx -> gather on (B,3,N) -> back to (B,G,3)."""
    xyz = np.asarray(xyz, dtype=F32)
    idx = fps(xyz, G)
    return np.ascontiguousarray(gather(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx).transpose(0, 2, 1))


# ---------------------------------------------------------------------------------------------
# a4 KNN(k, transpose_mode=True) (models/Point_MAE.py:55,68)
# ---------------------------------------------------------------------------------------------
def knn(ref, query, k: int):
    ref = np.asarray(ref, dtype=F32)
    query = np.asarray(query, dtype=F32)
    B, N, _ = ref.shape
    G = query.shape[1]
    if k > N:
        raise ValueError("k > N")
    D = np.empty((B, G, k), dtype=F32)
    I = np.empty((B, G, k), dtype=np.int64)
    for b in range(B):
        d = sumsq_acc(ref[b, None, :, 0] - query[b, :, None, 0], ref[b, None, :, 1] - query[b, :, None, 1],
                      ref[b, None, :, 2] - query[b, :, None, 2])  # (G,N)
        order = np.argsort(d, axis=1, kind="stable")[:, :k]  # ascending by (distance, index)
        I[b] = order
        D[b] = np.sqrt(np.take_along_axis(d, order, axis=1))
    return D, I


# ---------------------------------------------------------------------------------------------
# a5 Group.forward (models/Point_MAE.py:57-78; ..._feature_besed.py:1238-1260)
# ---------------------------------------------------------------------------------------------
def group(xyz, G: int, k: int):
    xyz = np.asarray(xyz, dtype=F32)
    B, N, _ = xyz.shape
    center = fps_centers(xyz, G)
    _, idx = knn(xyz, center, k)
    idx_base = np.arange(B).reshape(-1, 1, 1) * N
    flat = (idx + idx_base).reshape(-1)
    nb_org = xyz.reshape(B * N, 3)[flat].reshape(B, G, k, 3)
    nb = nb_org - center[:, :, None, :]
    return {"center": center, "knn_idx": idx, "neighborhood": nb.astype(F32), "neighborhood_org": nb_org}


# ---------------------------------------------------------------------------------------------
# a6 chamfer
# ---------------------------------------------------------------------------------------------
def chamfer_fwd(xyz1, xyz2):
    a = np.asarray(xyz1, dtype=F32)
    b = np.asarray(xyz2, dtype=F32)
    d = sumsq_nvcc(b[:, None, :, 0] - a[:, :, None, 0], b[:, None, :, 1] - a[:, :, None, 1],
                   b[:, None, :, 2] - a[:, :, None, 2])  # (P,n,m)
    idx1 = np.argmin(d, axis=2).astype(np.int32)  # first minimum = lowest index
    idx2 = np.argmin(d, axis=1).astype(np.int32)
    return d.min(axis=2), d.min(axis=1), idx1, idx2


def chamfer_bwd(xyz1, xyz2, idx1, idx2, gdist1, gdist2):
    """float64 accumulation: the 'true' gradient both fp32 summation orders are compared against."""
    a = np.asarray(xyz1, dtype=np.float64)
    b = np.asarray(xyz2, dtype=np.float64)
    P, n, _ = a.shape
    m = b.shape[1]
    ga = np.zeros_like(a)
    gb = np.zeros_like(b)
    pi = np.arange(P)[:, None]
    t1 = 2.0 * np.asarray(gdist1, dtype=np.float64)[..., None] * (a - b[pi, idx1])  # (P,n,3)
    ga += t1
    np.add.at(gb, (pi.repeat(n, 1), idx1), -t1)
    t2 = 2.0 * np.asarray(gdist2, dtype=np.float64)[..., None] * (b - a[pi, idx2])  # (P,m,3)
    gb += t2
    np.add.at(ga, (pi.repeat(m, 1), idx2), -t2)
    return ga, gb


def chamfer_l2(xyz1, xyz2) -> float:
    """ChamferDistanceL2.forward = mean(dist1) + mean(dist2)."""
    d1, d2, _, _ = chamfer_fwd(xyz1, xyz2)
    return float(d1.astype(np.float64).mean() + d2.astype(np.float64).mean())


def chamfer_l1(xyz1, xyz2) -> float:
    """ChamferDistanceL1.forward = (mean(sqrt dist1) + mean(sqrt dist2)) / 2."""
    d1, d2, _, _ = chamfer_fwd(xyz1, xyz2)
    return float((np.sqrt(d1.astype(np.float64)).mean() + np.sqrt(d2.astype(np.float64)).mean()) / 2)


# ---------------------------------------------------------------------------------------------
# a7 forward_loss -- usual mode (models_mae_learn_loss_Classifier_SVM.py:968-982).  The per-point
# tensor GM3D's modified extension returned is unknown; `per_point` picks the candidate.
# ---------------------------------------------------------------------------------------------
def forward_loss_usual(pred, target, mask, per_point: str = "dist1"):
    target = np.asarray(target, dtype=F32)
    N, t, n, D = target.shape
    mask = np.asarray(mask).astype(bool)
    tgt = target[mask].reshape(-1, n, D)
    prd = np.asarray(pred, dtype=F32).reshape(-1, n, D)
    d1, d2, _, _ = chamfer_fwd(prd, tgt)
    loss = {"dist1": d1, "dist2": d2, "sum": d1 + d2}[per_point].astype(np.float64)
    loss = loss.reshape(N, -1, n)
    return {"Chamfer_mean": loss.mean(), "matrix": loss.mean(axis=-1)}


# ---------------------------------------------------------------------------------------------
# a8 generate_mask (..._feature_besed.py:1062-1109) -- the deterministic part
# ---------------------------------------------------------------------------------------------
def mask_lengths(L: int, mask_ratio: float, epoch: int, total_epoch: int, guide: bool = True,
                 after_200_epoch=None, ratio_cap: float = 0.8):
    """len_keep / len_loss exactly as the reference computes them (python float arithmetic + int()).
    ratio_cap = 0.8 for ..._feature_besed.py:1083, 0.5 for ..._Classifier_SVM.py:1052."""
    len_keep = int(L * (1 - mask_ratio))
    keep_ratio = 0.5
    if guide:
        if after_200_epoch:
            keep_ratio = min(float((epoch + 1) / (total_epoch / 2)) * 0.5, 0.5)
        else:
            keep_ratio = float((epoch + 1) / total_epoch) * ratio_cap
    len_loss = int((L - len_keep) * keep_ratio)
    return len_keep, max(len_loss, 0)


def hard_mask(loss_pred, len_keep: int, len_loss: int, rand_keys) -> np.ndarray:
    loss_pred = np.asarray(loss_pred, dtype=F32)
    rand_keys = np.asarray(rand_keys, dtype=F32)
    B, L = loss_pred.shape
    mask = np.zeros((B, L), dtype=np.uint8)
    n_rand = L - len_keep - len_loss
    for b in range(B):
        order = np.argsort(loss_pred[b], kind="stable")  # ascending, ties: lower index first
        top = order[L - len_loss:] if len_loss > 0 else order[:0]
        mask[b, top] = 1
        rest = np.setdiff1d(np.arange(L), top)
        r_order = rest[np.argsort(rand_keys[b, rest], kind="stable")]
        if n_rand > 0:
            mask[b, r_order[len(r_order) - n_rand:]] = 1
    return mask


# ---------------------------------------------------------------------------------------------
# a9 _mask_center_rand (models/Point_MAE.py:297-320): exactly int(ratio*G) ones per row.
# ---------------------------------------------------------------------------------------------
def rand_mask(rand_keys, num_mask: int) -> np.ndarray:
    """Rows of (B,G) keys -> the num_mask largest keys (ties: higher index) are masked."""
    rand_keys = np.asarray(rand_keys, dtype=F32)
    B, G = rand_keys.shape
    return hard_mask(np.zeros((B, G), dtype=F32), G - num_mask, 0, rand_keys)


# ------------------------------------------------------------------ SURVEY 8(f) rows
def learning_loss(loss_pred, loss_target, relative):
    """forward_learning_loss and its gradient w.r.t. loss_pred, float64
    (/root/reference/Point-MAE_SA3D/models_mae_learn_loss_Classifier_SVM_feature_besed.py:1111-1135)."""
    p = np.asarray(loss_pred, dtype=np.float64)
    t = np.asarray(loss_target, dtype=np.float64)
    if relative:
        pos = t[:, None, :] > t[:, :, None]          # [n, i, j]: t_j > t_i
        neg = t[:, None, :] < t[:, :, None]
        m = p[:, None, :] - p[:, :, None]            # p_j - p_i
        s = 1.0 / (1.0 + np.exp(-m))
        valid = float(pos.sum() + neg.sum())
        loss = (-(pos * np.log(s + 1e-6)) - (neg * np.log(1 - s + 1e-6))).sum() / valid
        dm = (-(pos * s * (1 - s) / (s + 1e-6)) + (neg * s * (1 - s) / (1 - s + 1e-6))) / valid  # d loss / d m[n,i,j]
        grad = dm.sum(axis=1) - dm.sum(axis=2)       # + as second index j, - as first index i
        return loss, grad
    mean = t.mean(axis=1, keepdims=True)
    var = t.var(axis=1, keepdims=True, ddof=1)
    tn = (t - mean) / np.sqrt(var + 1e-6)
    d = p - tn
    return (d ** 2).mean(), 2.0 * d / d.size


def scale_translate(pc, scale_shift):
    """PointcloudScaleAndTranslate arithmetic (datasets/data_transforms.py:33): fp32 multiply, then fp32 add."""
    pc = np.array(pc, dtype=np.float32, copy=True)
    ss = np.asarray(scale_shift, dtype=np.float32)
    pc[:, :, 0:3] = (pc[:, :, 0:3] * ss[:, None, 0:3]).astype(np.float32) + ss[:, None, 3:6]
    return pc


def gather_points(xyz, idx, choice=None):
    """xyz (B,N,3), idx (B,G) -> xyz[b, idx[b, choice]] (engine_finetune.py:132-134 without the transposes)."""
    xyz, idx = np.asarray(xyz), np.asarray(idx)
    if choice is not None:
        idx = idx[:, np.asarray(choice)]
    return np.take_along_axis(xyz, idx[:, :, None].astype(np.int64), axis=1)


def encoder_eval(point_groups, sd, eps=1e-5):
    """Encoder.forward in eval mode, float64 (/root/reference/Point-MAE_SA3D/models/Point_MAE.py:16-47).
    point_groups (B,G,n,3); sd: dict of the reference state_dict arrays keyed like `first_conv.0.weight`."""
    x = np.asarray(point_groups, dtype=np.float64)
    bs, g, n, _ = x.shape
    x = x.reshape(bs * g, n, 3)

    def conv(v, w, b):  # v (P, n, cin), w (cout, cin, 1)
        return v @ np.asarray(w, dtype=np.float64)[:, :, 0].T + np.asarray(b, dtype=np.float64)

    def bn(v, pre):
        m, var = np.asarray(sd[pre + ".running_mean"], np.float64), np.asarray(sd[pre + ".running_var"], np.float64)
        return (v - m) / np.sqrt(var + eps) * np.asarray(sd[pre + ".weight"], np.float64) + np.asarray(sd[pre + ".bias"], np.float64)

    h = np.maximum(bn(conv(x, sd["first_conv.0.weight"], sd["first_conv.0.bias"]), "first_conv.1"), 0.0)
    f = conv(h, sd["first_conv.3.weight"], sd["first_conv.3.bias"])            # (P, n, 256)
    gmax = f.max(axis=1, keepdims=True)
    cat = np.concatenate([np.broadcast_to(gmax, f.shape), f], axis=2)           # [global ; per-point]
    h2 = np.maximum(bn(conv(cat, sd["second_conv.0.weight"], sd["second_conv.0.bias"]), "second_conv.1"), 0.0)
    out = conv(h2, sd["second_conv.3.weight"], sd["second_conv.3.bias"]).max(axis=1)
    return out.reshape(bs, g, -1)


# ---------------------------------------------------------------------------------------------
# a9' _mask_center_block (models/Point_MAE.py:268-295): per cloud the int(ratio * G) centres nearest to one picked
# centre (torch.norm of the difference, argsort ascending) are masked
# ---------------------------------------------------------------------------------------------
def mask_center_block(center, mask_ratio: float, picks):
    center = np.asarray(center, dtype=F32)
    B, G, _ = center.shape
    num = int(mask_ratio * G)
    out = np.zeros((B, G), dtype=bool)
    for b in range(B):
        d = np.sqrt(((center[b, picks[b]][None] - center[b]) ** 2).sum(-1, dtype=F32))
        out[b, np.argsort(d, kind="stable")[:num]] = True
    return out
