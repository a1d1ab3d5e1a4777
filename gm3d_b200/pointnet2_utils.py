"""Drop-in for `pointnet2_ops.pointnet2_utils` -- the two functions the GM3D path calls
(/root/reference/Point-MAE_SA3D/utils/miscc.py:18-19, engine_finetune.py:132-134).

Same names, argument meaning and error behaviour as erikwijmans/Pointnet2_PyTorch: inputs must be CUDA,
contiguous, float32 / int32; `furthest_point_sample` is non-differentiable, `gather_operation` is
differentiable w.r.t. `features`.
"""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import ops


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        r"""xyz (B, N, 3) f32 -> (B, npoint) int32 indices; starts from point 0."""
        out = ops.furthest_point_sample(xyz, npoint)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        r"""features (B, C, N) f32, idx (B, npoint) int32 -> (B, C, npoint)."""
        ctx.save_for_backward(idx)
        ctx.n = features.size(2)
        return ops.gather(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return ops.gather_grad(grad_out.contiguous(), idx, ctx.n), None


gather_operation = GatherOperation.apply


def fps(data: torch.Tensor, number: int) -> torch.Tensor:
    """miscc.fps / Group.fps (utils/miscc.py:13-20): (B,N,3) -> centres (B,number,3).
    One launch: the FPS kernel writes the gathered centres itself (no transposes, no gather kernel)."""
    _, centers = ops.fps_centers(data, number, want_centers=True)
    return centers
