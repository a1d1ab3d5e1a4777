"""Drop-ins for GeoMask3D's hard-patch mask selection and Point-MAE's random mask.

generate_mask: /root/reference/Point-MAE_SA3D/models_mae_learn_loss_Classifier_SVM_feature_besed.py:1062-1109
               (ratio cap 0.8) and models_mae_learn_loss_Classifier_SVM.py:1037-1080 (ratio cap 0.5).
mask_center_rand: models/Point_MAE.py:297-320.

The reference loops over the batch on the host with two device syncs per sample and a NumPy shuffle; here it
is one kernel launch with no sync.  The random part is distributionally identical (uniform without
replacement over the non-top patches) and exactly reproducible from (seed, offset) or from explicit
`rand_keys`; the top-`len_loss` set and the per-row cardinality are exact.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

_PHILOX_OFFSET = 0  # advanced per call so successive masks differ when the caller does not pass an offset


def mask_lengths(L: int, mask_ratio: float, epoch: int, total_epoch: int, guide: bool = True,
                 after_200_epoch=None, ratio_cap: float = 0.8):
    """len_keep and len_loss with the reference's own float arithmetic and int() truncation."""
    len_keep = int(L * (1 - mask_ratio))
    keep_ratio = 0.5
    if guide:
        if after_200_epoch:
            keep_ratio = min(float((epoch + 1) / (total_epoch / 2)) * 0.5, 0.5)
        else:
            keep_ratio = float((epoch + 1) / total_epoch) * ratio_cap
    len_loss = int((L - len_keep) * keep_ratio)
    return len_keep, max(len_loss, 0)


def _next_offset(count: int) -> int:
    global _PHILOX_OFFSET
    off = _PHILOX_OFFSET
    _PHILOX_OFFSET += count
    return off


@torch.no_grad()
def generate_mask(loss_pred: torch.Tensor, mask_ratio: float = 0.75, images=None, guide: bool = True, epoch: int = 0,
                  total_epoch: int = 200, after_200_epoch=None, ratio_cap: float = 0.8,
                  rand_keys: Optional[torch.Tensor] = None, seed: Optional[int] = None,
                  offset: Optional[int] = None) -> torch.Tensor:
    """loss_pred (N, L) -> float mask (N, L), 0 is keep, 1 is remove, exactly L - int(L*(1-mask_ratio)) ones per row."""
    N, L = loss_pred.shape
    len_keep, len_loss = mask_lengths(L, mask_ratio, epoch, total_epoch, guide, after_200_epoch, ratio_cap)
    if seed is None:
        seed = torch.initial_seed()
    if offset is None:
        offset = _next_offset(N * L)
    lp = loss_pred.float().contiguous() if len_loss > 0 else None
    m = ops.hard_mask(lp, N, L, len_keep, len_loss, rand_keys=rand_keys, seed=seed, offset=offset,
                      device=loss_pred.device)
    return m.to(torch.float32)


@torch.no_grad()
def mask_center_rand(center: torch.Tensor, mask_ratio: float, noaug: bool = False,
                     rand_keys: Optional[torch.Tensor] = None, seed: Optional[int] = None,
                     offset: Optional[int] = None) -> torch.Tensor:
    """center (B, G, 3) -> bool mask (B, G) with exactly int(mask_ratio * G) ones per row."""
    B, G, _ = center.shape
    if noaug or mask_ratio == 0:
        return torch.zeros(center.shape[:2], dtype=torch.bool, device=center.device)
    num_mask = int(mask_ratio * G)
    if seed is None:
        seed = torch.initial_seed()
    if offset is None:
        offset = _next_offset(B * G)
    m = ops.hard_mask(None, B, G, G - num_mask, 0, rand_keys=rand_keys, seed=seed, offset=offset,
                      device=center.device)
    return m.view(torch.bool)


@torch.no_grad()
def mask_center_block(center: torch.Tensor, mask_ratio: float, noaug: bool = False,
                      index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Drop-in for Point_MAE's `_mask_center_block` (/root/reference/Point-MAE_SA3D/models/Point_MAE.py:268-295):
    per cloud one random centre and the int(mask_ratio * G) centres nearest to it (itself included) are masked.
    center (B, G, 3) -> bool mask (B, G).  `index` (B,) chooses the seed centres; None draws them with Python's
    `random.randint(0, G - 1)` per cloud in batch order -- the reference's own RNG stream.  The selection is the
    hard-mask kernel on -|c - c_seed|^2 (the reference's argsort of the norm is not a stable sort either, so exact
    distance ties are unordered in both)."""
    import random
    B, G, _ = center.shape
    if noaug or mask_ratio == 0:
        return torch.zeros(center.shape[:2], dtype=torch.bool, device=center.device)
    num_mask = int(mask_ratio * G)
    if index is None:
        index = torch.tensor([random.randint(0, G - 1) for _ in range(B)], dtype=torch.long)
    index = index.to(device=center.device, dtype=torch.long)
    c = center.float()
    seed = c.gather(1, index.view(B, 1, 1).expand(B, 1, 3))
    key = -((c - seed) ** 2).sum(-1)  # largest key = nearest centre
    m = ops.hard_mask(key.contiguous(), B, G, G - num_mask, num_mask, device=center.device)
    return m.view(torch.bool)
