"""ctypes binding of oracle/gm3d_oracle.c (the CPU restatement).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (gm3d_b200/) never imports this.
PARITY UNPINNED for operator arithmetic (see gm3d_oracle.c header); the reference's call sites
are pinned by tests/golden/.

All functions take / return NumPy arrays (C-contiguous, exact dtypes) so nothing here depends on
torch.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgm3d_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile gm3d_oracle.c with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "gm3d_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libgm3d_oracle.so"], check=True, capture_output=True)
    return _SO


def _host_has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in line + " "
    except OSError:
        pass
    return True


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    elif not _host_has_fma():
        # the shipped .so may have been built with -mfma on another host
        build(force=True)
    L = ctypes.CDLL(_SO)
    f32p = ctypes.POINTER(ctypes.c_float)
    i32p = ctypes.POINTER(ctypes.c_int32)
    i64p = ctypes.POINTER(ctypes.c_int64)
    u8p = ctypes.POINTER(ctypes.c_uint8)
    f64p = ctypes.POINTER(ctypes.c_double)
    I = ctypes.c_int
    L.orc_num_threads.restype = I
    L.orc_set_num_threads.argtypes = [I]
    L.orc_fmaf_array.argtypes = [f32p, f32p, f32p, ctypes.c_int64, f32p]
    L.orc_fps.argtypes = [f32p, I, I, I, I, I, I, i32p]
    L.orc_gather.argtypes = [f32p, i32p, I, I, I, I, f32p]
    L.orc_gather_grad.argtypes = [f32p, i32p, I, I, I, I, f32p]
    L.orc_knn.argtypes = [f32p, f32p, I, I, I, I, f32p, i64p]
    L.orc_knn.restype = I
    L.orc_group.argtypes = [f32p, I, I, I, I, i32p, f32p, i64p, f32p, f32p]
    L.orc_group.restype = I
    L.orc_chamfer_fwd.argtypes = [f32p, f32p, I, I, I, f32p, f32p, i32p, i32p]
    L.orc_chamfer_bwd.argtypes = [f32p, f32p, i32p, i32p, f32p, f32p, I, I, I, f32p, f32p]
    L.orc_chamfer_per_patch.argtypes = [f32p, f32p, I, I, I, I, f64p]
    L.orc_hard_mask.argtypes = [f32p, I, I, I, I, f32p, u8p]
    L.orc_hard_mask.restype = I
    _lib = L
    return L


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def fmaf(a, b, c) -> np.ndarray:
    a, b, c = _f32(a).ravel(), _f32(b).ravel(), _f32(c).ravel()
    out = np.empty_like(a)
    lib().orc_fmaf_array(_p(a, ctypes.c_float), _p(b, ctypes.c_float), _p(c, ctypes.c_float), a.size,
                         _p(out, ctypes.c_float))
    return out


def fps(xyz, G: int, tie: str = "lowest_index", block: int = 0, skip_near_origin: bool = True) -> np.ndarray:
    """furthest_point_sample: xyz (B,N,3) f32 -> (B,G) int32."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    idx = np.zeros((B, G), dtype=np.int32)
    tie_mode = {"lowest_index": 0, "pointnet2_thread_order": 1}[tie]
    if tie_mode == 1 and block <= 0:
        block = 1
        while block * 2 <= min(N, 512):  # opt_n_threads(N): largest power of two <= min(N, 512)
            block *= 2
    lib().orc_fps(_p(xyz, ctypes.c_float), B, N, G, tie_mode, block, int(skip_near_origin), _p(idx, ctypes.c_int32))
    return idx


def gather(features, idx) -> np.ndarray:
    """gather_operation: features (B,C,N), idx (B,G) -> (B,C,G)."""
    features, idx = _f32(features), _i32(idx)
    B, C, N = features.shape
    G = idx.shape[1]
    out = np.empty((B, C, G), dtype=np.float32)
    lib().orc_gather(_p(features, ctypes.c_float), _p(idx, ctypes.c_int32), B, C, N, G, _p(out, ctypes.c_float))
    return out


def gather_grad(gout, idx, N: int) -> np.ndarray:
    gout, idx = _f32(gout), _i32(idx)
    B, C, G = gout.shape
    gfeat = np.empty((B, C, N), dtype=np.float32)
    lib().orc_gather_grad(_p(gout, ctypes.c_float), _p(idx, ctypes.c_int32), B, C, N, G, _p(gfeat, ctypes.c_float))
    return gfeat


def knn(ref, query, k: int):
    """KNN(k, transpose_mode=True): ref (B,N,3), query (B,G,3) -> D (B,G,k) f32 euclidean, I (B,G,k) int64."""
    ref, query = _f32(ref), _f32(query)
    B, N, _ = ref.shape
    G = query.shape[1]
    dist = np.empty((B, G, k), dtype=np.float32)
    idx = np.empty((B, G, k), dtype=np.int64)
    rc = lib().orc_knn(_p(ref, ctypes.c_float), _p(query, ctypes.c_float), B, N, G, k, _p(dist, ctypes.c_float),
                       _p(idx, ctypes.c_int64))
    if rc:
        raise ValueError(f"orc_knn: invalid arguments (k={k}, N={N})")
    return dist, idx


def group(xyz, G: int, k: int):
    """Group.forward: xyz (B,N,3) -> dict(fps_idx, center, knn_idx, neighborhood, neighborhood_org)."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    fps_idx = np.empty((B, G), dtype=np.int32)
    center = np.empty((B, G, 3), dtype=np.float32)
    knn_idx = np.empty((B, G, k), dtype=np.int64)
    nb = np.empty((B, G, k, 3), dtype=np.float32)
    nb_org = np.empty((B, G, k, 3), dtype=np.float32)
    rc = lib().orc_group(_p(xyz, ctypes.c_float), B, N, G, k, _p(fps_idx, ctypes.c_int32), _p(center, ctypes.c_float),
                         _p(knn_idx, ctypes.c_int64), _p(nb, ctypes.c_float), _p(nb_org, ctypes.c_float))
    if rc:
        raise ValueError(f"orc_group: invalid arguments (N={N}, G={G}, k={k})")
    return {"fps_idx": fps_idx, "center": center, "knn_idx": knn_idx, "neighborhood": nb, "neighborhood_org": nb_org}


def chamfer_fwd(xyz1, xyz2):
    """ChamferFunction.forward: (P,n,3),(P,m,3) -> dist1 (P,n), dist2 (P,m), idx1, idx2 (int32)."""
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    P, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.empty((P, n), dtype=np.float32)
    d2 = np.empty((P, m), dtype=np.float32)
    i1 = np.empty((P, n), dtype=np.int32)
    i2 = np.empty((P, m), dtype=np.int32)
    lib().orc_chamfer_fwd(_p(xyz1, ctypes.c_float), _p(xyz2, ctypes.c_float), P, n, m, _p(d1, ctypes.c_float),
                          _p(d2, ctypes.c_float), _p(i1, ctypes.c_int32), _p(i2, ctypes.c_int32))
    return d1, d2, i1, i2


def chamfer_bwd(xyz1, xyz2, idx1, idx2, gdist1, gdist2):
    xyz1, xyz2, gdist1, gdist2 = _f32(xyz1), _f32(xyz2), _f32(gdist1), _f32(gdist2)
    idx1, idx2 = _i32(idx1), _i32(idx2)
    P, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.empty((P, n, 3), dtype=np.float32)
    g2 = np.empty((P, m, 3), dtype=np.float32)
    lib().orc_chamfer_bwd(_p(xyz1, ctypes.c_float), _p(xyz2, ctypes.c_float), _p(idx1, ctypes.c_int32),
                          _p(idx2, ctypes.c_int32), _p(gdist1, ctypes.c_float), _p(gdist2, ctypes.c_float), P, n, m,
                          _p(g1, ctypes.c_float), _p(g2, ctypes.c_float))
    return g1, g2


def chamfer_per_patch(dist1, dist2, norm: int = 2) -> np.ndarray:
    dist1, dist2 = _f32(dist1), _f32(dist2)
    P, n = dist1.shape
    m = dist2.shape[1]
    out = np.empty((P,), dtype=np.float64)
    lib().orc_chamfer_per_patch(_p(dist1, ctypes.c_float), _p(dist2, ctypes.c_float), P, n, m, norm,
                                _p(out, ctypes.c_double))
    return out


def hard_mask(loss_pred, len_keep: int, len_loss: int, rand_keys) -> np.ndarray:
    """generate_mask with explicit random keys: (B,L) -> (B,L) uint8, 1 = masked."""
    loss_pred, rand_keys = _f32(loss_pred), _f32(rand_keys)
    B, L = loss_pred.shape
    mask = np.empty((B, L), dtype=np.uint8)
    rc = lib().orc_hard_mask(_p(loss_pred, ctypes.c_float), B, L, len_keep, len_loss, _p(rand_keys, ctypes.c_float),
                             _p(mask, ctypes.c_uint8))
    if rc:
        raise ValueError("orc_hard_mask: invalid len_keep / len_loss")
    return mask
