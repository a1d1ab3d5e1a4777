"""Reconstruction loss glue (`forward_loss`) on top of the fused Chamfer kernels.

usual mode   : /root/reference/Point-MAE_SA3D/models_mae_learn_loss_Classifier_SVM.py:968-982
feature mode : /root/reference/Point-MAE_SA3D/models_mae_learn_loss_Classifier_SVM_feature_besed.py:976-1003
stock scalar : /root/reference/Point-MAE_SA3D/models/Point_MAE.py:422-426

`target[mask]` is never materialised: the Chamfer kernel reads the masked target patches straight out of
the (B,G,k,3) neighbourhood tensor through a patch-index list produced by one small select kernel.
"""
from __future__ import annotations

import torch

from . import ops


class _MaskedChamfer(torch.autograd.Function):
    """per-point Chamfer of pred (P,n,3) against pool[patch_index] with grad w.r.t. pred only."""

    @staticmethod
    def forward(ctx, pred, pool, patch_index, norm, want):
        d1, d2, i1, i2, pp, _ = ops.chamfer_forward(pred, pool, norm=norm, want_per_patch=True,
                                                    xyz2_index=patch_index)
        ctx.save_for_backward(pred, pool, patch_index, i1, i2, d1, d2)
        ctx.want = want
        if want == "patch":
            return pp
        if want == "dist1":
            return d1
        return d1 + d2  # 'sum'

    @staticmethod
    def backward(ctx, grad):
        pred, pool, patch_index, i1, i2, d1, d2 = ctx.saved_tensors
        P, n = d1.shape
        m = d2.shape[1]
        grad = grad.contiguous()
        if ctx.want == "patch":
            g1 = (grad.reshape(P, 1) / n).expand(P, n).contiguous()
            g2 = (grad.reshape(P, 1) / m).expand(P, m).contiguous()
        elif ctx.want == "dist1":
            g1, g2 = grad, torch.zeros_like(d2)
        else:
            g1, g2 = grad, grad
        gx1, _ = ops.chamfer_backward(pred, pool, i1, i2, g1, g2, want_grad2=False, xyz2_index=patch_index)
        return gx1, None, None, None, None


class _FusedMaskedChamfer(torch.autograd.Function):
    """(mean loss, per-patch loss (P,)) of pred (P,n,3) against pool[patch_index] from ONE launch
    (gm3d_chamfer_fused_f32): the upstream gradient of a mean is uniform and known, so the launch that finds the
    arg-mins also emits d mean / d pred; backward() only scales it.  A gradient arriving through the per-patch
    output (the reference detaches it: engine_pretrain_Classifier_SVM.py:205-215) takes the general backward."""

    @staticmethod
    def forward(ctx, pred, pool, patch_index, norm):
        P, n, _ = pred.shape
        m = pool.shape[1]
        gs1, gs2 = ((1.0 / (P * n), 1.0 / (P * m)) if norm == 2 else (0.5 / (P * n), 0.5 / (P * m)))
        r = ops.chamfer_fused(pred, pool, gs1, gs2, norm=norm, xyz2_index=patch_index, want_dist=True)
        ctx.save_for_backward(pred, pool, patch_index, r["grad1"], r["idx1"], r["idx2"], r["dist1"], r["dist2"])
        ctx.norm = norm
        ctx.set_materialize_grads(False)
        mean, pp = r["total"].reshape(()), r["per_patch"]
        return mean, pp

    @staticmethod
    def backward(ctx, g_mean, g_pp):
        pred, pool, patch_index, grad1, i1, i2, d1, d2 = ctx.saved_tensors
        g = None
        if g_mean is not None:
            g = grad1 * g_mean
        if g_pp is not None:
            P, n = d1.shape
            m = d2.shape[1]
            u1 = (g_pp.reshape(P, 1) / n).expand(P, n)
            u2 = (g_pp.reshape(P, 1) / m).expand(P, m)
            if ctx.norm == 1:  # per_patch = (mean sqrt d1 + mean sqrt d2) / 2
                u1, u2 = u1 * 0.25 / d1.sqrt(), u2 * 0.25 / d2.sqrt()
            gx, _ = ops.chamfer_backward(pred, pool, i1, i2, u1.contiguous(), u2.contiguous(), want_grad2=False,
                                         xyz2_index=patch_index)
            g = gx if g is None else g + gx
        return g, None, None, None


class _FeatureMSE(torch.autograd.Function):
    """Rows of the normalised-feature MSE (pred (R,D) vs pool[index]); gradient w.r.t. pred only (the reference
    detaches the feature target).  One launch forward, one backward."""

    @staticmethod
    def forward(ctx, pred, pool, index):
        loss, _ = ops.feature_mse(pred, pool, index)
        ctx.save_for_backward(pred, pool, index)
        return loss

    @staticmethod
    def backward(ctx, g):
        pred, pool, index = ctx.saved_tensors
        _, grad = ops.feature_mse(pred, pool, index, gloss=g.contiguous().float(), want_loss=False, want_grad=True)
        return grad, None, None


def masked_patch_index(mask: torch.Tensor, num_masked: int) -> torch.Tensor:
    """(B,G) 0/1 mask (float, bool or uint8) -> (B*M,) int32 flat ids b*G+g of the masked patches, in order."""
    if mask.dtype not in (torch.bool, torch.uint8):
        mask = mask != 0
    _, idx = ops.select_patches(None, mask.contiguous(), num_masked, want_out=False, want_index=True)
    return idx


def forward_loss_usual(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor, per_point: str = "patch"):
    """pred (N, M, n*3) or (N*M, n, 3); target = neighborhood (N, G, n, 3); mask (N, G) with M ones per row.
    Returns {'MSE_mean', 'Chamfer_mean', 'matrix' (N, M)} like the reference.

    per_point -- GM3D's modified chamfer extension returns a per-point tensor whose exact form is not in the reference
    tree.  ONE convention is used throughout this package (loss.py, pipeline.GroupLossStep,
    gm3d_cloud_step_f32, ChamferDistanceL2(reduction='patch')): 'patch' -- the per-patch loss is
    mean_n dist1 + mean_m dist2, i.e. stock ChamferDistanceL2 per patch, and the whole forward + gradient is one
    launch.  'dist1' / 'sum' keep the two per-point candidates (matrix = mean_n of dist1, or of dist1 + dist2)."""
    N, G, n, D = target.shape
    M = pred.numel() // (N * n * D)
    index = masked_patch_index(mask, M)
    pred_ = pred.reshape(-1, n, D).to(dtype=torch.float32).contiguous()
    pool = target.reshape(N * G, n, D).to(dtype=torch.float32).contiguous()
    if per_point == "patch":
        mean, pp = _FusedMaskedChamfer.apply(pred_, pool, index, 2)
        return {"MSE_mean": mean * 0.0, "Chamfer_mean": mean, "matrix": pp.reshape(N, M)}
    loss = _MaskedChamfer.apply(pred_, pool, index, 2, per_point)
    loss = loss.reshape(N, -1, n)
    return {"MSE_mean": loss.mean() * 0.0, "Chamfer_mean": loss.mean(), "matrix": loss.mean(dim=-1)}


def forward_loss_feature(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor, point_target: torch.Tensor,
                         point_reconstructed: torch.Tensor, per_point: str = "patch"):
    """Feature mode: normalised-feature MSE (N, M) + per-patch Chamfer (N, M)."""
    N, P_, D = target.shape
    bmask = mask if mask.dtype == torch.bool else mask != 0
    PP = pred.numel() // (N * D)
    index = masked_patch_index(bmask, PP)  # serves both halves: `target[mask]` and `point_target[mask]`
    loss_mse = _FeatureMSE.apply(pred.reshape(N * PP, D).to(dtype=torch.float32).contiguous(),
                                 target.reshape(N * P_, D).to(dtype=torch.float32).contiguous(), index).reshape(N, PP)

    n = point_target.shape[2]
    rec = point_reconstructed.reshape(N * PP, -1, 3).to(dtype=torch.float32).contiguous()
    pool = point_target.reshape(N * P_, n, 3).to(dtype=torch.float32).contiguous()
    if per_point == "patch":
        loss_chamfer = _MaskedChamfer.apply(rec, pool, index, 2, "patch").reshape(N, PP)
    else:
        loss_chamfer = _MaskedChamfer.apply(rec, pool, index, 2, per_point).reshape(N, PP, -1).mean(-1)
    return {"MSE_mean": loss_mse.mean(), "Chamfer_mean": loss_chamfer.mean(), "matrix": loss_mse + loss_chamfer}


class _LearningLoss(torch.autograd.Function):
    """value + gradient w.r.t. loss_pred from ONE launch (the kernel that forms the pairwise terms also emits
    their derivatives; backward only scales by the upstream scalar)."""

    @staticmethod
    def forward(ctx, loss_pred, loss_target, relative):
        loss, grad = ops.learning_loss(loss_pred, loss_target, relative, 1.0, want_grad=loss_pred.requires_grad)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g if grad is not None else None), None, None


def forward_learning_loss(loss_pred: torch.Tensor, mask, loss_target: torch.Tensor, relative: bool = False):
    """Drop-in for MaskedAutoencoderViT.forward_learning_loss
    (/root/reference/Point-MAE_SA3D/models_mae_learn_loss_Classifier_SVM_feature_besed.py:1111-1135; called at
    engine_pretrain_Classifier_SVM.py:205-215 with the (N, M) per-patch Chamfer matrix as `loss_target`).
    `mask` is accepted and ignored, as in the reference.  loss_pred (N, L) or (N, L, 1)."""
    if loss_pred.dim() == 3 and loss_pred.shape[-1] == 1:
        loss_pred = loss_pred.squeeze(-1)
    return _LearningLoss.apply(loss_pred.float().contiguous(), loss_target.detach().float().contiguous(), bool(relative))
