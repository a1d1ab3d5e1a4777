"""ctypes binding of libgm3d_sm100.so (the C ABI in include/gm3d.h).

There is no fallback: if the shared library is missing or does not export the ABI the import of any
operator fails with an ImportError that says how to build it (`python -m gm3d_b200.build`).
"""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_PKG, "libgm3d_sm100.so")

GM3D_ABI_VERSION = 5
GM3D_EINVAL, GM3D_ENOSUP, GM3D_EALIGN = -1, -2, -3
OP_FPS, OP_KNN, OP_GROUP, OP_CHAMFER_FWD, OP_CHAMFER_BWD, OP_HARD_MASK, OP_LOSS_STATS, OP_CLOUD_STEP, OP_LEARNING_LOSS = range(1, 10)
KNN_MAX_K = 32
LOSS_STATS_LEN = 8
STEP_OVERLAP_NEXT, STEP_OVERLAP_PREV, STEP_AFTER_PREV, STEP_SHARED_SMS = 1, 2, 4, 8
MAX_PEERS = 8
INBOX_DEPTH = 4
INBOX_BYTES = INBOX_DEPTH * MAX_PEERS * 32


class StepReduce(ctypes.Structure):
    """gm3d_step_reduce_t (include/gm3d.h): where the tail of a loss launch publishes / all-reduces the step's
    {sum, sum_sq, count}.  Passed by HOST pointer; the library copies it into the kernel parameters."""
    _fields_ = [("head", ctypes.c_void_p), ("world", ctypes.c_int), ("rank", ctypes.c_int),
                ("inbox", ctypes.c_void_p * MAX_PEERS), ("epoch", ctypes.c_void_p),
                ("timeout_us", ctypes.c_uint), ("defer", ctypes.c_int), ("status", ctypes.c_void_p),
                ("collected", ctypes.c_void_p)]


_vp, _i, _u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64

# name -> (restype, argtypes); every symbol include/gm3d.h declares
SIGNATURES = {
    "gm3d_abi_version": (_i, []),
    "gm3d_strerror": (ctypes.c_char_p, [_i]),
    "gm3d_workspace_bytes": (ctypes.c_size_t, [_i, _i, _i, _i, _i]),
    "gm3d_fps_f32": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "gm3d_gather_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "gm3d_gather_grad_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "gm3d_knn_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "gm3d_knn_general_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "gm3d_knn_group_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "gm3d_group_f32": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gm3d_chamfer_fwd_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "gm3d_chamfer_fused_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    "gm3d_chamfer_bwd_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, _i, _i, _i, _vp, _vp, _vp]),
    "gm3d_select_patches_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "gm3d_feature_mse_f32": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "gm3d_hard_mask_f32": (_i, [_vp, _i, _i, _i, _i, _vp, _u64, _u64, _vp, _vp, _i, _vp]),
    "gm3d_loss_stats_f32": (_i, [_vp, _i, _vp, _vp]),
    "gm3d_learning_loss_f32": (_i, [_vp, _vp, _i, _i, _i, ctypes.c_float, _vp, _vp, _vp, _vp]),
    "gm3d_scale_translate_f32": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "gm3d_gather_points_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "gm3d_encoder_fwd_bf16": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "gm3d_cloud_step_f32": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _u64, _u64, _vp, _vp, _vp,
                                 ctypes.c_float, ctypes.c_float, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "gm3d_step_reduce_collect": (_i, [_vp, _i, _vp]),
    "gm3d_peer_alloc": (_i, [ctypes.c_size_t, ctypes.POINTER(_vp), ctypes.c_char_p]),
    "gm3d_peer_open": (_i, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "gm3d_peer_close": (_i, [_vp]),
    "gm3d_peer_free": (_i, [_vp]),
}

_lib = None


class Gm3dError(RuntimeError):
    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed: {msg} (code {code})")
        self.code = code


def load() -> ctypes.CDLL:
    """Load the library once; raise ImportError (never fall back) when it is unusable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: the gm3d_b200 operators have no CPU or PyTorch fallback. "
            "Build the sm_100a library with `python -m gm3d_b200.build`.")
    try:
        lib = ctypes.CDLL(SO_PATH)
    except OSError as e:  # e.g. libcudart not resolvable
        raise ImportError(f"cannot load {SO_PATH}: {e}") from e
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError(f"{SO_PATH} does not export {name}; rebuild with `python -m gm3d_b200.build --force`") from e
        fn.restype, fn.argtypes = res, args
    got = lib.gm3d_abi_version()
    if got != GM3D_ABI_VERSION:
        raise ImportError(f"{SO_PATH} has ABI version {got}, this package needs {GM3D_ABI_VERSION}; rebuild it")
    _lib = lib
    return lib


def strerror(code: int) -> str:
    return load().gm3d_strerror(code).decode()


def check(fn: str, code: int) -> None:
    if code == 0:
        return
    msg = strerror(code)
    if code == GM3D_EINVAL:
        raise ValueError(f"{fn}: {msg}")
    if code == GM3D_ENOSUP:
        raise NotImplementedError(f"{fn}: {msg}")
    raise Gm3dError(fn, code, msg)
