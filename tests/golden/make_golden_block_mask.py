"""Golden vectors for `_mask_center_block` (/root/reference/Point-MAE_SA3D/models/Point_MAE.py:268-295), produced by
the reference's own method imported unmodified (stubs for the absent third-party packages: make_golden.install_stubs).
Run in the build container (needs /root/reference); writes tests/golden/reference_block_mask.npz.

    python tests/golden/make_golden_block_mask.py
"""
from __future__ import annotations

import importlib
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    mg.install_stubs()
    sys.path.insert(0, mg.REF)
    pm = importlib.import_module("models.Point_MAE")
    out = {}
    for tag, (B, G, ratio, seed) in {"c1": (8, 64, 0.6, 11), "m2ae_l2": (5, 64, 0.8, 12), "small": (3, 40, 0.25, 13)}.items():
        class _Self:
            mask_ratio = ratio

        centers = mg.synthetic_clouds(B, G, 1260 + seed)
        random.seed(seed)  # the method draws the seed centre of every cloud with random.randint
        m = pm.MaskTransformer._mask_center_block(_Self(), centers)
        random.seed(seed)
        picks = [random.randint(0, G - 1) for _ in range(B)]
        out[f"{tag}_centers"] = centers.numpy()
        out[f"{tag}_mask"] = m.numpy()
        out[f"{tag}_ratio"] = np.array([ratio])
        out[f"{tag}_seed"] = np.array([seed])
        out[f"{tag}_picks"] = np.array(picks)
    out["cases"] = np.array(["c1", "m2ae_l2", "small"])
    path = os.path.join(HERE, "reference_block_mask.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
