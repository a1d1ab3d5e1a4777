"""Generate tests/golden/*.npz by running the REFERENCE'S OWN Python glue code from /root/reference.

Run once in the build container (it cannot run on the GPU box: /root/reference is absent there):

    python tests/golden/make_golden.py

What this pins.  The reference's call-site code is imported unmodified and executed on CPU:
  * models/Point_MAE.py:50-78              Group.forward            (Point-MAE variant)
  * ..._feature_besed.py:1222-1260         Group.forward            (GM3D variant, + neighborhood_org)
  * utils/miscc.py:13-20                   fps
  * models/Point_MAE.py:297-320            MaskTransformer._mask_center_rand
  * ..._feature_besed.py:1062-1109         generate_mask            (ratio cap 0.8)
  * ..._Classifier_SVM.py:1037-1080        generate_mask            (ratio cap 0.5)
  * ..._Classifier_SVM.py:968-982          forward_loss             (usual mode)
  * ..._feature_besed.py:976-1003          forward_loss             (feature mode)
  * datasets/ModelNetDataset.py:25-46      farthest_point_sample    (NumPy CPU FPS)
What it cannot pin.  The three CUDA extensions those files import (pointnet2_ops, knn_cuda,
extensions.chamfer_dist) are not in the tree, so they are stubbed here with oracle/c_oracle.py.  The
goldens therefore pin the reference's index arithmetic, reshapes, selection and reductions AROUND the
operators -- not the operators' own arithmetic, which stays "parity unpinned" (see DESIGN.md).
Other absent third-party imports (timm, easydict, matplotlib, h5py ...) are stubbed with empty modules;
none of their code runs.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/Point-MAE_SA3D"
sys.path.insert(0, ROOT)

from oracle import c_oracle as co  # noqa: E402

CHAMFER_PER_POINT = {"mode": "scalar"}  # scalar | dist1 | sum  


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []  # behave as a package
        sys.modules[name] = m
        return m

    class _Anything(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    # --- absent third-party packages whose code never runs here
    mod("timm")
    mod("timm.models")
    mod("timm.models.layers", DropPath=_Anything, trunc_normal_=lambda *a, **k: None)
    mod("timm.models.vision_transformer", PatchEmbed=_Anything, Block=_Anything, DropPath=_Anything, Mlp=_Anything)
    mod("matplotlib")
    mod("matplotlib.pyplot")
    mod("mpl_toolkits")
    mod("mpl_toolkits.mplot3d", Axes3D=object)
    mod("easydict", EasyDict=dict)
    mod("h5py")
    mod("termcolor", colored=lambda s, *a, **k: s)

    # --- the three operator extensions, backed by the oracle
    def furthest_point_sample(xyz, npoint):
        return _t(co.fps(xyz.detach().numpy(), int(npoint)))

    def gather_operation(features, idx):
        return _t(co.gather(features.detach().numpy(), idx.numpy()))

    pn2 = mod("pointnet2_ops")
    pn2.pointnet2_utils = mod("pointnet2_ops.pointnet2_utils", furthest_point_sample=furthest_point_sample,
                              gather_operation=gather_operation)

    class KNN(nn.Module):
        def __init__(self, k, transpose_mode=False):
            super().__init__()
            self.k, self._t = k, transpose_mode

        def forward(self, ref, query):
            assert self._t
            D, I = co.knn(ref.detach().numpy(), query.detach().numpy(), self.k)
            return _t(D), _t(I)

    mod("knn_cuda", KNN=KNN)

    class ChamferDistanceL2(nn.Module):
        def __init__(self, ignore_zeros=False):
            super().__init__()

        def forward(self, xyz1, xyz2):
            d1, d2, _, _ = co.chamfer_fwd(xyz1.detach().numpy(), xyz2.detach().numpy())
            d1, d2 = _t(d1), _t(d2)
            if CHAMFER_PER_POINT["mode"] == "dist1":
                return d1
            if CHAMFER_PER_POINT["mode"] == "sum":
                return d1 + d2
            return torch.mean(d1) + torch.mean(d2)

    class ChamferDistanceL1(nn.Module):
        def __init__(self, ignore_zeros=False):
            super().__init__()

        def forward(self, xyz1, xyz2):
            d1, d2, _, _ = co.chamfer_fwd(xyz1.detach().numpy(), xyz2.detach().numpy())
            return (torch.mean(torch.sqrt(_t(d1))) + torch.mean(torch.sqrt(_t(d2)))) / 2

    mod("extensions")
    mod("extensions.chamfer_dist", ChamferDistanceL1=ChamferDistanceL1, ChamferDistanceL2=ChamferDistanceL2)


def synthetic_clouds(B, N, seed):
    """unit-ball normalised gaussian cloud, per-cloud scale U(2/3,3/2) + translate U(-.2,.2)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N, 3, generator=g)
    x = x - x.mean(dim=1, keepdim=True)
    x = x / x.norm(dim=-1).amax(dim=1).view(B, 1, 1)
    scale = torch.rand(B, 1, 3, generator=g) * (3 / 2 - 2 / 3) + 2 / 3
    shift = torch.rand(B, 1, 3, generator=g) * 0.4 - 0.2
    return (x * scale + shift).float().contiguous()


def main():
    install_stubs()
    sys.path.insert(0, REF)
    import importlib

    pm = importlib.import_module("models.Point_MAE")
    fb = importlib.import_module("models_mae_learn_loss_Classifier_SVM_feature_besed")
    sv = importlib.import_module("models_mae_learn_loss_Classifier_SVM")
    miscc = importlib.import_module("utils.miscc")
    mnd = importlib.import_module("datasets.ModelNetDataset")
    out = {}

    # ---- Group.forward, both variants, three shapes (incl. ragged N not a multiple of 32 and G == N)
    for tag, (B, N, G, k, seed) in {"c1": (8, 1024, 64, 32, 1235), "m2ae_l2": (4, 256, 64, 8, 1237),
                                    "ragged": (3, 333, 17, 5, 77), "g_eq_n": (2, 96, 96, 4, 78)}.items():
        xyz = synthetic_clouds(B, N, seed)
        nb, ctr = pm.Group(G, k)(xyz)
        nb2, ctr2, nb_org = fb.Group(G, k)(xyz)
        assert torch.equal(nb, nb2) and torch.equal(ctr, ctr2)
        ctr_fps = miscc.fps(xyz, G)
        assert torch.equal(ctr_fps, ctr)
        out[f"group_{tag}_xyz"] = xyz.numpy()
        out[f"group_{tag}_G_k"] = np.array([G, k])
        out[f"group_{tag}_neighborhood"] = nb.numpy()
        out[f"group_{tag}_center"] = ctr.numpy()
        out[f"group_{tag}_neighborhood_org"] = nb_org.numpy()

    # ---- generate_mask: both ratio caps, several epochs (len_loss from 0 to max)
    gen = torch.Generator().manual_seed(1236)
    loss_pred = torch.randn(16, 64, generator=gen)
    out["mask_loss_pred"] = loss_pred.numpy()
    cases = []
    for cls_tag, cls in (("fb", fb.MaskedAutoencoderViT), ("sv", sv.MaskedAutoencoderViT)):
        for epoch, total, after in ((0, 400, None), (9, 400, None), (199, 400, None), (399, 400, None),
                                    (99, 400, True), (350, 400, True)):
            np.random.seed(1000 + epoch)
            torch.manual_seed(1000 + epoch)
            m = cls.generate_mask(None, loss_pred, mask_ratio=0.6, guide=True, epoch=epoch, total_epoch=total,
                                  after_200_epoch=after)
            key = f"mask_{cls_tag}_e{epoch}_t{total}_a{int(bool(after))}"
            out[key] = m.numpy()
            cases.append(key)
    out["mask_cases"] = np.array(cases)

    # ---- _mask_center_rand (only needs self.mask_ratio)
    class _Self:
        mask_ratio = 0.6

    np.random.seed(5)
    s = _Self()
    rm = pm.MaskTransformer._mask_center_rand(s, torch.zeros(8, 64, 3))
    out["rand_mask_c1"] = rm.numpy()
    out["rand_mask_c1_num_mask"] = np.array([s.num_mask])

    # ---- forward_loss, usual mode (Chamfer only) under both candidate per-point definitions
    xyz = synthetic_clouds(4, 1024, 1240)
    nb, ctr, _ = fb.Group(64, 32)(xyz)
    gen = torch.Generator().manual_seed(1241)
    mask = torch.from_numpy(co.hard_mask(torch.randn(4, 64, generator=gen).numpy(), 25, 15,
                                         torch.rand(4, 64, generator=gen).numpy())).bool()
    gt = nb[mask].reshape(4, 39, 32, 3)
    pred = (gt + 0.02 * torch.randn(gt.shape, generator=gen)).reshape(4, 39, 96)
    out["loss_neighborhood"] = nb.numpy()
    out["loss_mask"] = mask.numpy()
    out["loss_pred_points"] = pred.numpy()

    class _LossSelf:
        loss_func = sys.modules["extensions.chamfer_dist"].ChamferDistanceL2()

    for mode in ("dist1", "sum"):
        CHAMFER_PER_POINT["mode"] = mode
        r = sv.MaskedAutoencoderViT.forward_loss(_LossSelf(), pred, nb, mask)
        out[f"loss_usual_{mode}_chamfer_mean"] = np.array(r["Chamfer_mean"].item(), dtype=np.float64)
        out[f"loss_usual_{mode}_matrix"] = r["matrix"].numpy()
        # feature mode: adds the normalised-feature MSE term; point_target = neighbourhoods, same mask
        feat_t = torch.randn(4, 64, 384, generator=torch.Generator().manual_seed(1242))
        feat_p = torch.randn(4, 39, 384, generator=torch.Generator().manual_seed(1243))
        r2 = fb.MaskedAutoencoderViT.forward_loss(_LossSelf(), feat_p, feat_t, mask, nb, pred)
        out[f"loss_feature_{mode}_matrix"] = r2["matrix"].numpy()
        out[f"loss_feature_{mode}_chamfer_mean"] = np.array(r2["Chamfer_mean"].item(), dtype=np.float64)
        out[f"loss_feature_{mode}_mse_mean"] = np.array(r2["MSE_mean"].item(), dtype=np.float64)
    out["loss_feature_target"] = feat_t.numpy()
    out["loss_feature_pred"] = feat_p.numpy()

    # ---- stock Point-MAE scalar losses (models/Point_MAE.py:390-397,426)
    CHAMFER_PER_POINT["mode"] = "scalar"
    cd = sys.modules["extensions.chamfer_dist"]
    a = pred.reshape(-1, 32, 3)
    b = gt.reshape(-1, 32, 3)
    out["cdl2_scalar"] = np.array(cd.ChamferDistanceL2()(a, b).item(), dtype=np.float64)
    out["cdl1_scalar"] = np.array(cd.ChamferDistanceL1()(a, b).item(), dtype=np.float64)

    # ---- in-tree NumPy FPS (random start made reproducible by seeding NumPy)
    pts = synthetic_clouds(1, 512, 1250)[0].numpy().astype(np.float64)
    np.random.seed(42)
    start = np.random.randint(0, 512)
    np.random.seed(42)
    sel = mnd.farthest_point_sample(pts, 48)
    out["npfps_points"] = pts
    out["npfps_start"] = np.array([start])
    out["npfps_selected"] = sel

    path = os.path.join(HERE, "reference_glue.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
