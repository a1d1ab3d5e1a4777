"""Generate tests/golden/reference_next.npz: goldens for the SURVEY 8(f) "next" rows, produced by the
REFERENCE'S OWN functions imported from /root/reference (build container only; see make_golden.py for the
stubbing of the absent third-party packages -- none of the stubbed code runs for these rows):

  * ..._feature_besed.py:1111-1135   forward_learning_loss (relative=True / False), value + autograd gradient
  * datasets/data_transforms.py:20-35  PointcloudScaleAndTranslate (NumPy RNG seeded; `.cuda()` redirected to CPU)
  * models/Point_MAE.py:16-47        Encoder (mini-PointNet), eval mode, seeded weights
  * engine_finetune.py:118-134       FPS -> random column subset -> gather (inline code restated with the
                                     reference's own operator calls, backed by the oracle stubs)
    python tests/golden/make_golden_next.py
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

REF = mg.REF


def main():
    mg.install_stubs()
    sys.path.insert(0, REF)
    fb = importlib.import_module("models_mae_learn_loss_Classifier_SVM_feature_besed")
    pm = importlib.import_module("models.Point_MAE")
    dt = importlib.import_module("datasets.data_transforms")
    pn2 = sys.modules["pointnet2_ops.pointnet2_utils"]
    out = {}

    # ---- forward_learning_loss: (N, L) predictions vs the per-patch Chamfer matrix (ties included)
    gen = torch.Generator().manual_seed(2001)
    for tag, (n, L) in {"c2": (16, 39), "small": (3, 7), "m2ae": (4, 52)}.items():
        pred = torch.randn(n, L, generator=gen)
        target = torch.rand(n, L, generator=gen) * 0.05
        target[0, 1] = target[0, 0]  # an exact tie: neither positive nor negative
        out[f"ll_{tag}_pred"], out[f"ll_{tag}_target"] = pred.numpy(), target.numpy()
        for rel in (True, False):
            p = pred.clone().requires_grad_(True)
            loss = fb.MaskedAutoencoderViT.forward_learning_loss(None, p, None, target, relative=rel)
            loss.backward()
            out[f"ll_{tag}_rel{int(rel)}_loss"] = np.array(loss.item(), dtype=np.float64)
            out[f"ll_{tag}_rel{int(rel)}_grad"] = p.grad.numpy()

    # ---- PointcloudScaleAndTranslate (in place; two NumPy draws of 3 per sample)
    torch.Tensor.cuda = lambda self, *a, **k: self  # the transform calls .cuda() on its host constants
    pc = mg.synthetic_clouds(6, 128, 2002)
    out["sat_in"] = pc.numpy().copy()
    np.random.seed(2003)
    res = dt.PointcloudScaleAndTranslate()(pc.clone())
    out["sat_out"] = res.numpy()
    np.random.seed(2003)
    draws = np.stack([np.concatenate([np.random.uniform(2. / 3., 3. / 2., 3), np.random.uniform(-0.2, 0.2, 3)]) for _ in range(6)])
    out["sat_draws"] = draws  # (B, 6) float64: scale xyz, shift xyz -- what the module consumed

    # ---- Encoder, eval mode (BatchNorm with seeded running statistics)
    torch.manual_seed(2004)
    enc = pm.Encoder(384)
    with torch.no_grad():
        for m in enc.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.2)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.1)
    enc.eval()
    xyz = mg.synthetic_clouds(2, 512, 2005)
    nb, _ = pm.Group(16, 32)(xyz)
    with torch.no_grad():
        feat = enc(nb)
    out["enc_neighborhood"] = nb.numpy()
    out["enc_out"] = feat.numpy()
    for name, t in enc.state_dict().items():
        out["enc_w_" + name.replace(".", "_")] = t.numpy()

    # ---- fine-tune sub-sampling (engine_finetune.py:118-134): FPS to point_all, random npoints columns, gather
    pts = mg.synthetic_clouds(3, 2048, 2006)
    npoints, point_all = 1024, 1200
    fps_idx = pn2.furthest_point_sample(pts, point_all)
    np.random.seed(2007)
    choice = np.random.choice(point_all, npoints, False)
    fps_sel = fps_idx[:, choice]
    sub = pn2.gather_operation(pts.transpose(1, 2).contiguous(), fps_sel).transpose(1, 2).contiguous()
    out["ft_points"], out["ft_choice"], out["ft_out"] = pts.numpy(), choice.astype(np.int64), sub.numpy()
    out["ft_point_all"] = np.array([point_all])

    path = os.path.join(HERE, "reference_next.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
