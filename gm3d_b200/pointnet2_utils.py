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


def fps_subsample(points: torch.Tensor, npoints: int, point_all: int = None, choice=None) -> torch.Tensor:
    """The fine-tune / vote-test sub-sampling block (engine_finetune.py:118-134, tools/runner_finetune.py:127-143):
    FPS down to `point_all` points (1200 / 2400 / 4800 / 8192 for npoints 1024 / 2048 / 4096 / 8192, capped at N),
    a random `npoints`-column subset of the FPS order (`np.random.choice(point_all, npoints, False)`, drawn here
    with the same call unless `choice` is given), gather -> (B, npoints, 3).  Two launches, no transposes."""
    import numpy as np
    table = {1024: 1200, 2048: 2400, 4096: 4800, 8192: 8192}
    if point_all is None:
        if npoints not in table:
            raise NotImplementedError()
        point_all = table[npoints]
    point_all = min(int(point_all), points.size(1))
    fps_idx = ops.furthest_point_sample(points, point_all)
    if choice is None:
        choice = np.random.choice(point_all, npoints, False)
    choice = torch.as_tensor(np.asarray(choice), dtype=torch.int64).to(points.device)
    return ops.gather_points(points, fps_idx, choice)
