"""Drop-in for `extensions.chamfer_dist` (Point-MAE lineage, package `chamfer` 2.0.0):
/root/reference/Point-MAE_SA3D/models/Point_MAE.py:13,390-397,426 and
models_mae_learn_loss_Classifier_SVM_feature_besed.py:26,937,996.

`ChamferFunction`, `ChamferDistanceL2`, `ChamferDistanceL2_split`, `ChamferDistanceL1` keep the stock
behaviour.  The extra keyword `reduction` serves GM3D's per-point / per-patch use of the loss (the locally
modified extension GM3D used is not in the reference tree -- SURVEY F5):
    'mean'  (default) stock scalar
    'none'            (dist1, dist2) per point
    'dist1'           dist1 (P, n)            -- candidate A for GM3D's per-point tensor
    'sum'             dist1 + dist2 (P, n)    -- candidate B (needs n == m)
    'patch'           (P,) = mean_n dist1 + mean_m dist2 (L2) or (mean sqrt + mean sqrt)/2 (L1), fused in-kernel
"""
from __future__ import annotations

import torch

from . import ops


class ChamferFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2, _, _ = ops.chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        grad_xyz1, grad_xyz2 = ops.chamfer_backward(xyz1, xyz2, idx1, idx2, grad_dist1.contiguous(),
                                                    grad_dist2.contiguous(), want_grad2=ctx.needs_input_grad[1])
        return grad_xyz1, grad_xyz2


class _ChamferReduced(torch.autograd.Function):
    """Forward with the per-patch / scalar reduction fused into the kernel; backward rebuilds the per-point
    upstream gradients of that reduction and calls the same atomics-free backward kernel."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, norm: int, scalar: bool):
        d1, d2, i1, i2, pp, tot = ops.chamfer_forward(xyz1, xyz2, norm=norm, want_per_patch=True, want_total=scalar)
        ctx.save_for_backward(xyz1, xyz2, i1, i2, d1, d2)
        ctx.norm, ctx.scalar = norm, scalar
        return tot.reshape(()) if scalar else pp

    @staticmethod
    def backward(ctx, grad):
        xyz1, xyz2, i1, i2, d1, d2 = ctx.saved_tensors
        P, n = d1.shape
        m = d2.shape[1]
        g = (grad / P).reshape(1, 1).expand(P, 1) if ctx.scalar else grad.reshape(P, 1)
        if ctx.norm == 2:
            g1 = (g / n).expand(P, n).contiguous()
            g2 = (g / m).expand(P, m).contiguous()
        else:  # d/dx sqrt(x) = 1 / (2 sqrt x); the /2 of the L1 mean folded in
            g1 = (g * (0.5 / n)) * (0.5 / torch.sqrt(d1))
            g2 = (g * (0.5 / m)) * (0.5 / torch.sqrt(d2))
        gx1, gx2 = ops.chamfer_backward(xyz1, xyz2, i1, i2, g1.contiguous(), g2.contiguous(),
                                        want_grad2=ctx.needs_input_grad[1])
        return gx1, gx2, None, None


def _ignore_zeros(xyz1, xyz2):
    non_zeros1 = torch.sum(xyz1, dim=2).ne(0)
    non_zeros2 = torch.sum(xyz2, dim=2).ne(0)
    return xyz1[non_zeros1].unsqueeze(dim=0), xyz2[non_zeros2].unsqueeze(dim=0)


class _ChamferBase(torch.nn.Module):
    norm = 2

    def __init__(self, ignore_zeros: bool = False, reduction: str = "mean"):
        super().__init__()
        if reduction not in ("mean", "none", "dist1", "sum", "patch"):
            raise ValueError(f"unknown reduction {reduction!r}")
        self.ignore_zeros = ignore_zeros
        self.reduction = reduction

    def _prep(self, xyz1, xyz2):
        if xyz1.size(0) == 1 and self.ignore_zeros:
            xyz1, xyz2 = _ignore_zeros(xyz1, xyz2)
        return xyz1.contiguous(), xyz2.contiguous()

    def _per_point(self, xyz1, xyz2):
        dist1, dist2 = ChamferFunction.apply(xyz1, xyz2)
        if self.norm == 1:
            dist1, dist2 = torch.sqrt(dist1), torch.sqrt(dist2)
        return dist1, dist2

    def forward(self, xyz1, xyz2):
        xyz1, xyz2 = self._prep(xyz1, xyz2)
        r = self.reduction
        if r == "mean":
            return _ChamferReduced.apply(xyz1, xyz2, self.norm, True)
        if r == "patch":
            return _ChamferReduced.apply(xyz1, xyz2, self.norm, False)
        dist1, dist2 = self._per_point(xyz1, xyz2)
        if r == "none":
            return dist1, dist2
        if r == "dist1":
            return dist1
        return dist1 + dist2  # 'sum'


class ChamferDistanceL2(_ChamferBase):
    """mean(dist1) + mean(dist2) (stock)."""
    norm = 2


class ChamferDistanceL1(_ChamferBase):
    """(mean(sqrt dist1) + mean(sqrt dist2)) / 2 (stock)."""
    norm = 1


class ChamferDistanceL2_split(torch.nn.Module):
    """(mean(dist1), mean(dist2)) (stock)."""

    def __init__(self, ignore_zeros: bool = False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2):
        if xyz1.size(0) == 1 and self.ignore_zeros:
            xyz1, xyz2 = _ignore_zeros(xyz1, xyz2)
        dist1, dist2 = ChamferFunction.apply(xyz1.contiguous(), xyz2.contiguous())
        return torch.mean(dist1), torch.mean(dist2)
