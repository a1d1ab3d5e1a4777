"""Tensor-level wrappers around the C ABI (include/gm3d.h): check arguments the way the reference's
extensions do (CUDA, contiguous, dtype -> RuntimeError / ValueError), allocate outputs with torch, pass raw
device pointers and the current stream.  PyTorch is plumbing here (memory + streams); all arithmetic runs in
libgm3d_sm100.so.  No CPU path: a CPU tensor raises.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

__all__ = [
    "furthest_point_sample", "fps_centers", "gather", "gather_grad", "knn", "group", "chamfer_forward",
    "chamfer_fused", "chamfer_backward", "select_patches", "hard_mask", "feature_mse", "loss_stats", "learning_loss",
    "scale_translate_", "gather_points",
]


def _req(t: torch.Tensor, name: str, dtype: torch.dtype, ndim: Optional[int] = None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (gm3d_b200 has no CPU fallback)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name} must have {ndim} dimensions, got shape {tuple(t.shape)}")
    return t


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _ws(op: int, B: int, N: int, G: int, k: int, like: torch.Tensor) -> Optional[torch.Tensor]:
    n = _lib.load().gm3d_workspace_bytes(op, B, N, G, k)
    return torch.empty(n, dtype=torch.uint8, device=like.device) if n else None


def furthest_point_sample(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    """xyz (B,N,3) f32 -> (B,npoint) int32.  pointnet2_utils.furthest_point_sample (utils/miscc.py:18)."""
    idx, _ = fps_centers(xyz, npoint, want_centers=False)
    return idx


def fps_centers(xyz: torch.Tensor, npoint: int, want_centers: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """FPS indices and (fused) the gathered centres (B,npoint,3) -- miscc.fps in one launch."""
    _req(xyz, "xyz", torch.float32, 3)
    if xyz.shape[2] != 3:
        raise ValueError(f"xyz must be (B,N,3), got {tuple(xyz.shape)}")
    B, N, _ = xyz.shape
    npoint = int(npoint)
    with torch.cuda.device(xyz.device):
        idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        centers = torch.empty((B, npoint, 3), dtype=torch.float32, device=xyz.device) if want_centers else None
        if B == 0 or npoint == 0:
            return idx, centers
        ws = _ws(_lib.OP_FPS, B, N, npoint, 0, xyz)
        rc = _lib.load().gm3d_fps_f32(_p(xyz), B, N, npoint, _p(idx), _p(centers), _p(ws), _stream(xyz))
    _lib.check("gm3d_fps_f32", rc)
    return idx, centers


def gather(features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """features (B,C,N) f32, idx (B,G) int32 -> (B,C,G).  pointnet2_utils.gather_operation forward."""
    _req(features, "features", torch.float32, 3)
    _req(idx, "idx", torch.int32, 2)
    B, C, N = features.shape
    G = idx.shape[1]
    if idx.shape[0] != B:
        raise ValueError("features and idx disagree on the batch size")
    with torch.cuda.device(features.device):
        out = torch.empty((B, C, G), dtype=torch.float32, device=features.device)
        if out.numel() == 0:
            return out
        rc = _lib.load().gm3d_gather_f32(_p(features), _p(idx), B, C, N, G, _p(out), _stream(features))
    _lib.check("gm3d_gather_f32", rc)
    return out


def gather_grad(grad_out: torch.Tensor, idx: torch.Tensor, N: int) -> torch.Tensor:
    """grad_out (B,C,G), idx (B,G) -> grad_features (B,C,N); deterministic scatter-add."""
    _req(grad_out, "grad_out", torch.float32, 3)
    _req(idx, "idx", torch.int32, 2)
    B, C, G = grad_out.shape
    with torch.cuda.device(grad_out.device):
        g = torch.empty((B, C, N), dtype=torch.float32, device=grad_out.device)
        if g.numel() == 0:
            return g
        if G == 0:
            return g.zero_()
        rc = _lib.load().gm3d_gather_grad_f32(_p(grad_out), _p(idx), B, C, N, G, _p(g), _stream(grad_out))
    _lib.check("gm3d_gather_grad_f32", rc)
    return g


def knn(ref: torch.Tensor, query: torch.Tensor, k: int, want_dist: bool = True):
    """ref (B,N,3), query (B,G,3) f32 -> D (B,G,k) f32 euclidean (or None), I (B,G,k) int64."""
    _req(ref, "ref", torch.float32, 3)
    _req(query, "query", torch.float32, 3)
    if ref.shape[2] != query.shape[2]:
        raise ValueError("ref and query disagree on the point dimension")
    if ref.shape[0] != query.shape[0]:
        raise ValueError("ref and query disagree on the batch size")
    B, N, dim = ref.shape
    G = query.shape[1]
    k = int(k)
    if k > N or k <= 0:
        raise ValueError(f"k={k} must satisfy 1 <= k <= N={N}")
    with torch.cuda.device(ref.device):
        I = torch.empty((B, G, k), dtype=torch.int64, device=ref.device)
        D = torch.empty((B, G, k), dtype=torch.float32, device=ref.device) if want_dist else None
        if I.numel() == 0:
            return D, I
        if dim == 3 and k <= _lib.KNN_MAX_K:  # every reference configuration: the specialised kernels
            rc = _lib.load().gm3d_knn_f32(_p(ref), _p(query), B, N, G, k, _p(D), _p(I), None, _stream(ref))
            _lib.check("gm3d_knn_f32", rc)
        else:  # no shape limits, like upstream: the general selection kernel
            rc = _lib.load().gm3d_knn_general_f32(_p(ref), _p(query), B, N, G, dim, k, _p(D), _p(I), _stream(ref))
            _lib.check("gm3d_knn_general_f32", rc)
    return D, I


def group(xyz: torch.Tensor, num_group: int, group_size: int, want_org: bool = False, want_idx: bool = False):
    """Fused Group.forward.  xyz (B,N,3) -> dict(neighborhood, center[, neighborhood_org, knn_idx], fps_idx)."""
    _req(xyz, "xyz", torch.float32, 3)
    if xyz.shape[2] != 3:
        raise ValueError(f"xyz must be (B,N,3), got {tuple(xyz.shape)}")
    B, N, _ = xyz.shape
    G, k = int(num_group), int(group_size)
    if G > N or k > N or G <= 0 or k <= 0:
        raise ValueError(f"need 1 <= num_group <= N and 1 <= group_size <= N (N={N}, G={G}, k={k})")
    dev = xyz.device
    with torch.cuda.device(dev):
        fps_idx = torch.empty((B, G), dtype=torch.int32, device=dev)
        center = torch.empty((B, G, 3), dtype=torch.float32, device=dev)
        nb = torch.empty((B, G, k, 3), dtype=torch.float32, device=dev)
        nb_org = torch.empty((B, G, k, 3), dtype=torch.float32, device=dev) if want_org else None
        knn_idx = torch.empty((B, G, k), dtype=torch.int64, device=dev) if want_idx else None
        if B > 0:
            ws = _ws(_lib.OP_GROUP, B, N, G, k, xyz)
            rc = _lib.load().gm3d_group_f32(_p(xyz), B, N, G, k, _p(fps_idx), _p(center), _p(knn_idx), _p(nb),
                                            _p(nb_org), _p(ws), _stream(xyz))
            _lib.check("gm3d_group_f32", rc)
    return {"neighborhood": nb, "center": center, "neighborhood_org": nb_org, "knn_idx": knn_idx, "fps_idx": fps_idx}


def _chamfer_ws(P: int, dev) -> torch.Tensor:
    """Zero-initialised workspace of the Chamfer forward (ticket + per-patch scratch); the kernel leaves the
    ticket zeroed, so a cached workspace could be reused -- a fresh torch.zeros keeps the wrapper stateless."""
    n = _lib.load().gm3d_workspace_bytes(_lib.OP_CHAMFER_FWD, P, 0, 0, 0)
    return torch.zeros(n, dtype=torch.uint8, device=dev)


def _chamfer_args(xyz1, xyz2, xyz2_index):
    _req(xyz1, "xyz1", torch.float32, 3)
    _req(xyz2, "xyz2", torch.float32, 3)
    if xyz1.shape[2] != 3 or xyz2.shape[2] != 3:
        raise ValueError("chamfer inputs must be (P,n,3) and (P,m,3)")
    P, n, _ = xyz1.shape
    m = xyz2.shape[1]
    if xyz2_index is not None:
        _req(xyz2_index, "xyz2_index", torch.int32, 1)
        if xyz2_index.shape[0] != P:
            raise ValueError("xyz2_index must have one entry per xyz1 patch")
    elif xyz2.shape[0] != P:
        raise ValueError("xyz1 and xyz2 disagree on the number of patches")
    return P, n, m


def chamfer_forward(xyz1: torch.Tensor, xyz2: torch.Tensor, norm: int = 2, want_per_patch: bool = False,
                    want_total: bool = False, xyz2_index: Optional[torch.Tensor] = None, want_stats: bool = False):
    """xyz1 (P,n,3), xyz2 (P,m,3) [or the patch pool + xyz2_index (P,) int32] ->
    dist1 (P,n), dist2 (P,m), idx1, idx2 (int32), per_patch (P,) or None, total (1,) or None[, stats (8,)]."""
    P, n, m = _chamfer_args(xyz1, xyz2, xyz2_index)
    dev = xyz1.device
    with torch.cuda.device(dev):
        d1 = torch.empty((P, n), dtype=torch.float32, device=dev)
        d2 = torch.empty((P, m), dtype=torch.float32, device=dev)
        i1 = torch.empty((P, n), dtype=torch.int32, device=dev)
        i2 = torch.empty((P, m), dtype=torch.int32, device=dev)
        reduce = want_total or want_stats
        pp = torch.empty((P,), dtype=torch.float32, device=dev) if (want_per_patch or reduce) else None
        tot = torch.empty((1,), dtype=torch.float32, device=dev) if want_total else None
        stats = torch.empty((_lib.LOSS_STATS_LEN,), dtype=torch.float32, device=dev) if want_stats else None
        if P > 0 and n > 0 and m > 0:
            ws = _chamfer_ws(P, dev) if reduce else None
            rc = _lib.load().gm3d_chamfer_fwd_f32(_p(xyz1), _p(xyz2), _p(xyz2_index), P, n, m, _p(d1), _p(d2), _p(i1),
                                                  _p(i2), _p(pp), _p(tot), _p(stats), int(norm), _p(ws), _stream(xyz1))
            _lib.check("gm3d_chamfer_fwd_f32", rc)
    if want_stats:
        return d1, d2, i1, i2, pp, tot, stats
    return d1, d2, i1, i2, pp, tot


def chamfer_fused(xyz1: torch.Tensor, xyz2: torch.Tensor, gscale1: float, gscale2: float, norm: int = 2,
                  want_grad2: bool = False, xyz2_index: Optional[torch.Tensor] = None, want_dist: bool = False):
    """Forward + backward of the mean-reduced Chamfer loss in one launch (n, m <= 32).
    Returns dict(per_patch (P,), total (1,), stats (8,), grad1 (P,n,3), grad2 or None[, dist1, dist2, idx1, idx2])."""
    P, n, m = _chamfer_args(xyz1, xyz2, xyz2_index)
    if xyz2_index is not None and want_grad2:
        raise ValueError("grad_xyz2 is not defined for an indexed (shared) xyz2 pool")
    dev = xyz1.device
    e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
    with torch.cuda.device(dev):
        out = {"per_patch": e((P,), torch.float32), "total": e((1,), torch.float32),
               "stats": e((_lib.LOSS_STATS_LEN,), torch.float32), "grad1": e((P, n, 3), torch.float32),
               "grad2": e((P, m, 3), torch.float32) if want_grad2 else None,
               "dist1": e((P, n), torch.float32) if want_dist else None,
               "dist2": e((P, m), torch.float32) if want_dist else None,
               "idx1": e((P, n), torch.int32) if want_dist else None,
               "idx2": e((P, m), torch.int32) if want_dist else None}
        if P > 0 and n > 0 and m > 0:
            ws = _chamfer_ws(P, dev)
            rc = _lib.load().gm3d_chamfer_fused_f32(_p(xyz1), _p(xyz2), _p(xyz2_index), P, n, m, float(gscale1),
                                                    float(gscale2), _p(out["dist1"]), _p(out["dist2"]), _p(out["idx1"]),
                                                    _p(out["idx2"]), _p(out["per_patch"]), _p(out["total"]),
                                                    _p(out["stats"]), int(norm), _p(out["grad1"]), _p(out["grad2"]),
                                                    None, 0, _p(ws), _stream(xyz1))
            _lib.check("gm3d_chamfer_fused_f32", rc)
    return out


def chamfer_backward(xyz1, xyz2, idx1, idx2, gdist1, gdist2, want_grad2: bool = True,
                     xyz2_index: Optional[torch.Tensor] = None, gscale1: float = 1.0, gscale2: float = 1.0):
    """Atomics-free Chamfer backward -> grad_xyz1 (P,n,3), grad_xyz2 (P,m,3) or None.
    gdist1 / gdist2 may be None: the upstream gradient is then the scalar gscale1 / gscale2 everywhere."""
    _req(xyz1, "xyz1", torch.float32, 3)
    _req(xyz2, "xyz2", torch.float32, 3)
    _req(idx1, "idx1", torch.int32, 2)
    _req(idx2, "idx2", torch.int32, 2)
    if gdist1 is not None:
        _req(gdist1, "grad_dist1", torch.float32, 2)
    if gdist2 is not None:
        _req(gdist2, "grad_dist2", torch.float32, 2)
    P, n, _ = xyz1.shape
    m = xyz2.shape[1]
    if xyz2_index is not None:
        _req(xyz2_index, "xyz2_index", torch.int32, 1)
        if want_grad2:
            raise ValueError("grad_xyz2 is not defined for an indexed (shared) xyz2 pool")
    dev = xyz1.device
    with torch.cuda.device(dev):
        g1 = torch.empty((P, n, 3), dtype=torch.float32, device=dev)
        g2 = torch.empty((P, m, 3), dtype=torch.float32, device=dev) if want_grad2 else None
        if P > 0 and n > 0 and m > 0:
            rc = _lib.load().gm3d_chamfer_bwd_f32(_p(xyz1), _p(xyz2), _p(xyz2_index), _p(idx1), _p(idx2), _p(gdist1),
                                                  _p(gdist2), float(gscale1), float(gscale2), P, n, m, _p(g1), _p(g2),
                                                  _stream(xyz1))
            _lib.check("gm3d_chamfer_bwd_f32", rc)
    return g1, g2


def select_patches(nbhd: Optional[torch.Tensor], mask: torch.Tensor, num_selected: int, invert: bool = False,
                   want_out: bool = True, want_index: bool = False, status: Optional[torch.Tensor] = None):
    """`nbhd[mask]` for a (B,G) mask with exactly num_selected ones per row.  nbhd (B,G,...) f32 ->
    out (B*M, ...) [, patch_index (B*M,) int32].  mask may be bool or uint8 (viewed, not copied)."""
    if mask.dtype == torch.bool:
        mask = mask.view(torch.uint8)
    _req(mask, "mask", torch.uint8, 2)
    B, G = mask.shape
    M = int(num_selected)
    dev = mask.device
    out = idx = None
    row = 1
    if want_out:
        _req(nbhd, "nbhd", torch.float32)
        if nbhd.shape[0] != B or nbhd.shape[1] != G:
            raise ValueError("nbhd must be (B,G,...) matching the mask")
        row = nbhd[0, 0].numel()
    with torch.cuda.device(dev):
        if want_out:
            out = torch.empty((B * M, *nbhd.shape[2:]), dtype=torch.float32, device=dev)
        if want_index:
            idx = torch.empty((B * M,), dtype=torch.int32, device=dev)
        if B > 0 and M > 0:
            rc = _lib.load().gm3d_select_patches_f32(_p(nbhd) if want_out else None, _p(mask), B, G, row, M,
                                                     int(bool(invert)), _p(out), _p(idx), _p(status), _stream(mask))
            _lib.check("gm3d_select_patches_f32", rc)
    return out, idx


def hard_mask(loss_pred: Optional[torch.Tensor], B: int, L: int, len_keep: int, len_loss: int,
              rand_keys: Optional[torch.Tensor] = None, seed: int = 0, offset: int = 0,
              device: Optional[torch.device] = None, want_index: bool = False):
    """(B,L) uint8 mask, 1 = masked, exactly L - len_keep ones per row (include/gm3d.h).
    want_index: also return patch_index (B*(L-len_keep),) int32 = flat ids b*L+i of the masked patches in order."""
    if loss_pred is not None:
        _req(loss_pred, "loss_pred", torch.float32, 2)
        device = loss_pred.device
    if rand_keys is not None:
        _req(rand_keys, "rand_keys", torch.float32, 2)
        device = rand_keys.device
    if device is None:
        raise ValueError("hard_mask needs loss_pred, rand_keys or an explicit device")
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("hard_mask runs on CUDA only (gm3d_b200 has no CPU fallback)")
    with torch.cuda.device(device):
        mask = torch.empty((B, L), dtype=torch.uint8, device=device)
        index = torch.empty((B * (L - int(len_keep)),), dtype=torch.int32, device=device) if want_index else None
        if B > 0 and L > 0:
            rc = _lib.load().gm3d_hard_mask_f32(_p(loss_pred), B, L, int(len_keep), int(len_loss), _p(rand_keys),
                                                int(seed) & (2**64 - 1), int(offset) & (2**64 - 1), _p(mask), _p(index),
                                                0, torch.cuda.current_stream(device).cuda_stream)
            _lib.check("gm3d_hard_mask_f32", rc)
    return (mask, index) if want_index else mask


def feature_mse(pred: torch.Tensor, target: torch.Tensor, index: Optional[torch.Tensor] = None,
                gloss: Optional[torch.Tensor] = None, want_loss: bool = True, want_grad: bool = False):
    """Normalised-feature MSE rows (include/gm3d.h: gm3d_feature_mse_f32).  pred (R,D), target (T,D) f32, index (R) int32
    or None -> (loss (R,) or None, grad (R,D) = gloss[:,None] * d loss / d pred or None)."""
    _req(pred, "pred", torch.float32, 2)
    _req(target, "target", torch.float32, 2)
    R, D = pred.shape
    if target.shape[1] != D:
        raise ValueError("pred and target disagree on the feature dimension")
    if index is not None:
        _req(index, "index", torch.int32, 1)
        if index.shape[0] != R:
            raise ValueError("index must have one entry per pred row")
    elif target.shape[0] != R:
        raise ValueError("pred and target disagree on the number of rows")
    if gloss is not None:
        _req(gloss, "gloss", torch.float32, 1)
    with torch.cuda.device(pred.device):
        loss = torch.empty((R,), dtype=torch.float32, device=pred.device) if want_loss else None
        grad = torch.empty((R, D), dtype=torch.float32, device=pred.device) if want_grad else None
        if R > 0 and D > 0:
            rc = _lib.load().gm3d_feature_mse_f32(_p(pred), _p(target), _p(index), R, D, _p(loss), _p(gloss), _p(grad),
                                                  _stream(pred))
            _lib.check("gm3d_feature_mse_f32", rc)
    return loss, grad


def loss_stats(per_patch: torch.Tensor) -> torch.Tensor:
    """per_patch (P,) -> stats (8,) = [sum, sum_sq, count, min, max, mean, 0, 0] for the step's one all-reduce."""
    _req(per_patch, "per_patch", torch.float32)
    with torch.cuda.device(per_patch.device):
        stats = torch.empty((_lib.LOSS_STATS_LEN,), dtype=torch.float32, device=per_patch.device)
        rc = _lib.load().gm3d_loss_stats_f32(_p(per_patch), per_patch.numel(), _p(stats), _stream(per_patch))
    _lib.check("gm3d_loss_stats_f32", rc)
    return stats


# ------------------------------------------------------------------ either side of the hot path
_LL_WS = {}


def learning_loss(loss_pred: torch.Tensor, loss_target: torch.Tensor, relative: bool, gscale: float = 1.0,
                  want_grad: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """forward_learning_loss value (0-d) and gscale * d loss / d loss_pred in one launch
    (..._feature_besed.py:1111-1135)."""
    _req(loss_pred, "loss_pred", torch.float32, 2)
    _req(loss_target, "loss_target", torch.float32, 2)
    if loss_pred.shape != loss_target.shape:
        raise ValueError(f"loss_pred {tuple(loss_pred.shape)} and loss_target {tuple(loss_target.shape)} differ")
    B, L = loss_pred.shape
    with torch.cuda.device(loss_pred.device):
        lib = _lib.load()
        loss = torch.empty((), dtype=torch.float32, device=loss_pred.device)
        grad = torch.empty_like(loss_pred) if want_grad else None
        need = lib.gm3d_workspace_bytes(_lib.OP_LEARNING_LOSS, B, 0, 0, 0)
        key = (loss_pred.device.index, torch.cuda.current_stream(loss_pred.device).cuda_stream)
        ws = _LL_WS.get(key)
        if ws is None or ws.numel() < need:  # one workspace per (device, stream): the ticket resets itself
            ws = _LL_WS[key] = torch.zeros(max(need, 4096), dtype=torch.uint8, device=loss_pred.device)
        _lib.check("gm3d_learning_loss_f32", lib.gm3d_learning_loss_f32(
            _p(loss_pred), _p(loss_target), B, L, int(bool(relative)), float(gscale), _p(loss), _p(grad), _p(ws), _stream(loss_pred)))
    return loss, grad


def scale_translate_(pc: torch.Tensor, scale_shift: torch.Tensor) -> torch.Tensor:
    """In place pc[b,:,0:3] = pc * scale[b] + shift[b]; scale_shift (B,6) f32 on the same device
    (datasets/data_transforms.py:20-35 in one launch)."""
    _req(pc, "pc", torch.float32, 3)
    _req(scale_shift, "scale_shift", torch.float32, 2)
    B, N, C = pc.shape
    if scale_shift.shape != (B, 6):
        raise ValueError(f"scale_shift must be ({B}, 6), got {tuple(scale_shift.shape)}")
    with torch.cuda.device(pc.device):
        _lib.check("gm3d_scale_translate_f32", _lib.load().gm3d_scale_translate_f32(_p(pc), _p(scale_shift), B, N, C, _stream(pc)))
    return pc


def gather_points(xyz: torch.Tensor, idx: torch.Tensor, choice: Optional[torch.Tensor] = None,
                  validate: bool = True) -> torch.Tensor:
    """out[b,j,:] = xyz[b, idx[b, choice[j]], :] -- `fps_idx[:, choice]` + gather_operation + both transposes of
    engine_finetune.py:132-134 as one gather.  xyz (B,N,3) f32, idx (B,G) int32, choice (K) int64 or None."""
    _req(xyz, "xyz", torch.float32, 3)
    _req(idx, "idx", torch.int32, 2)
    B, N, D = xyz.shape
    if D != 3 or idx.shape[0] != B:
        raise ValueError(f"xyz must be (B,N,3) and idx (B,G), got {tuple(xyz.shape)}, {tuple(idx.shape)}")
    G = idx.shape[1]
    if choice is not None:
        _req(choice, "choice", torch.int64, 1)
        # (range check reads the tensor back: pass validate=False inside a CUDA-graph capture)
        if validate and int(choice.numel()) and (int(choice.min()) < 0 or int(choice.max()) >= G):
            raise ValueError("choice holds a column outside [0, G)")
    K = int(choice.numel()) if choice is not None else G
    with torch.cuda.device(xyz.device):
        out = torch.empty((B, K, 3), dtype=torch.float32, device=xyz.device)
        _lib.check("gm3d_gather_points_f32", _lib.load().gm3d_gather_points_f32(_p(xyz), _p(idx), _p(choice), B, N, G, K, _p(out), _stream(xyz)))
    return out
