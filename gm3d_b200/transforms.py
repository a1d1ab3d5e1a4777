"""Drop-in for the batch augmentation the pre-training and fine-tuning loops apply right before `Group`
(/root/reference/Point-MAE_SA3D/datasets/data_transforms.py:20-35; engine_pretrain_Classifier_SVM.py:99-100,
engine_finetune.py:136).

The reference loops over the batch in Python: per sample two NumPy draws, two host->device copies and three
small kernels.  Here the 6*B random numbers are drawn with the SAME NumPy calls in the same order (so a seeded
run consumes the identical RNG stream and produces bit-identical clouds), cross PCIe once, and one kernel
scales and translates the whole batch in place.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


class PointcloudScaleAndTranslate(object):
    def __init__(self, scale_low=2. / 3., scale_high=3. / 2., translate_range=0.2):
        self.scale_low = scale_low
        self.scale_high = scale_high
        self.translate_range = translate_range

    def draw(self, bsize: int) -> np.ndarray:
        """(bsize, 6) float32: scale xyz, shift xyz per sample -- the reference's draws, in its order."""
        ss = np.empty((bsize, 6), dtype=np.float64)
        for i in range(bsize):
            ss[i, :3] = np.random.uniform(low=self.scale_low, high=self.scale_high, size=[3])
            ss[i, 3:] = np.random.uniform(low=-self.translate_range, high=self.translate_range, size=[3])
        return ss.astype(np.float32)  # the reference casts each draw with .float()

    def __call__(self, pc: torch.Tensor) -> torch.Tensor:
        ss = torch.from_numpy(self.draw(pc.size(0))).to(pc.device, non_blocking=True)
        return ops.scale_translate_(pc, ss)
