"""Drop-in for the models' `Group` module (FPS centres -> kNN patches -> centre-normalisation):
/root/reference/Point-MAE_SA3D/models/Point_MAE.py:50-78 (returns neighborhood, center) and the GM3D copies
models_mae_learn_loss_Classifier_SVM_feature_besed.py:1222-1260 (additionally returns neighborhood_org).

One C-ABI call (two kernel launches) instead of the reference's FPS + gather + 2 transposes + a Python
loop of 3 launches per cloud + ATen index + subtract.  Stateless: `state_dict()` is empty, so reference
checkpoints load unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .knn import KNN


class Group(nn.Module):  # FPS + KNN
    def __init__(self, num_group: int, group_size: int, return_org: bool = False):
        super().__init__()
        self.num_group = num_group
        self.group_size = group_size
        self.return_org = return_org
        self.knn = KNN(k=self.group_size, transpose_mode=True)  # kept for attribute parity; has no state

    def fps(self, data: torch.Tensor, number: int) -> torch.Tensor:
        """data (B,N,3) -> centres (B,number,3)  (Group.fps, ..._feature_besed.py:1229-1236)."""
        return ops.fps_centers(data, number)[1]

    def forward(self, xyz: torch.Tensor):
        """
            input: B N 3
            ---------------------------
            output: B G M 3   (centre-subtracted)
            center : B G 3
            [neighborhood_org : B G M 3 when return_org]
        """
        with torch.no_grad():  # FPS / kNN are non-differentiable upstream as well; xyz is raw data
            r = ops.group(xyz.float().contiguous(), self.num_group, self.group_size, want_org=self.return_org)
        if self.return_org:
            return r["neighborhood"], r["center"], r["neighborhood_org"]
        return r["neighborhood"], r["center"]


class GroupGM3D(Group):
    """The GM3D variant: forward returns (neighborhood, center, neighborhood_org)."""

    def __init__(self, num_group: int, group_size: int):
        super().__init__(num_group, group_size, return_org=True)


class GroupM2AE(nn.Module):
    """The hierarchical grouping of Point-M2AE(+GM3D): `Group` levels chained on the previous level's centres
    (/root/reference/Point-M2AE_SA3D/cfgs/config_Point_M2AE.yaml:57-69 -- num_groups [512, 256, 64], group_sizes
    [16, 8, 8]; fine-tune: sizes [32, 16, 16], config_finetune_scan_hardest_PointM2AE.yaml:58-69).  The model code
    of Point-M2AE+GM3D is not in the reference tree; the chaining rule is Point-M2AE's published one:
    level 0 groups the raw cloud, level l > 0 groups centres[l-1].
        forward(xyz (B,N,3)) -> (neighborhoods [ (B,G_l,k_l,3) ], centers [ (B,G_l,3) ], idxs [ (B,G_l,k_l) int64 ])
    idxs[l] index into level l's input cloud (xyz for l = 0, centers[l-1] otherwise)."""

    def __init__(self, num_groups=(512, 256, 64), group_sizes=(16, 8, 8)):
        super().__init__()
        self.num_groups, self.group_sizes = tuple(num_groups), tuple(group_sizes)

    @torch.no_grad()
    def forward(self, xyz):
        neighborhoods, centers, idxs = [], [], []
        cloud = xyz
        for g, k in zip(self.num_groups, self.group_sizes):
            r = ops.group(cloud.float().contiguous(), g, k, want_idx=True)
            neighborhoods.append(r["neighborhood"])
            centers.append(r["center"])
            idxs.append(r["knn_idx"])
            cloud = r["center"]
        return neighborhoods, centers, idxs
