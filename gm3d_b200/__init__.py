"""gm3d_b200 -- B200-native (sm_100a) point-grouping + reconstruction-loss path of GeoMask3D.

Drop-in operator surface (same names / arguments as the extensions the reference imports):
    gm3d_b200.pointnet2_utils.furthest_point_sample / gather_operation
    gm3d_b200.knn.KNN(k, transpose_mode)
    gm3d_b200.chamfer.ChamferDistanceL1 / ChamferDistanceL2 / ChamferDistanceL2_split / ChamferFunction
    gm3d_b200.group.Group / GroupGM3D
    gm3d_b200.masking.generate_mask / mask_center_rand
    gm3d_b200.loss.forward_loss_usual / forward_loss_feature
    gm3d_b200.dist.all_reduce_mean / all_reduce_stats
`install_shims()` makes the reference's own import lines (`from knn_cuda import KNN`, ...) resolve here.

All arithmetic runs in libgm3d_sm100.so through the C ABI of include/gm3d.h; there is no CPU, Triton or
PyTorch fallback -- importing an operator without the built library raises ImportError.
"""
from __future__ import annotations

import os
import sys

__version__ = "0.1.0"


def install_shims() -> str:
    """Put the import shims (`pointnet2_ops`, `knn_cuda`, `extensions.chamfer_dist`) on sys.path."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compat")
    if path not in sys.path:
        sys.path.insert(0, path)
    return path


def library_path() -> str:
    from . import _lib
    return _lib.SO_PATH
