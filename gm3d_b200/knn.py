"""Drop-in for `knn_cuda.KNN` (KNN_CUDA 0.2), as constructed at
/root/reference/Point-MAE_SA3D/models/Point_MAE.py:55 and called at :68.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class KNN(nn.Module):
    """k nearest references per query by Euclidean distance, ascending, ties -> lower reference index.

    transpose_mode=True : ref (B, N, dim), query (B, G, dim) -> D, I of shape (B, G, k)
    transpose_mode=False: ref (B, dim, N), query (B, dim, G) -> D, I of shape (B, k, G)
    D is float32 (sqrt applied), I is int64, 0-based.  Runs under no_grad like the original.
    The whole batch is one kernel launch (the original loops over the batch in Python).  dim == 3 and k <= 32 (every
    reference configuration) run the specialised kernels, anything else the general selection kernel.
    """

    def __init__(self, k: int, transpose_mode: bool = False):
        super().__init__()
        self.k = k
        self._t = transpose_mode

    def forward(self, ref: torch.Tensor, query: torch.Tensor):
        assert ref.size(0) == query.size(0), "ref.shape={} != query.shape={}".format(ref.shape, query.shape)
        with torch.no_grad():
            r, q = ref.float(), query.float()
            if not self._t:
                r, q = r.transpose(1, 2), q.transpose(1, 2)
            D, I = ops.knn(r.contiguous(), q.contiguous(), self.k, want_dist=True)
            if not self._t:
                D, I = D.transpose(1, 2).contiguous(), I.transpose(1, 2).contiguous()
        return D, I
