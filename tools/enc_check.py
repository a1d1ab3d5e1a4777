import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from gm3d_b200.encoder import EncoderB200
g = np.load("tests/golden/reference_next.npz")
sd = {k[len("enc_w_"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("enc_w_")}
def unmangle(k):
    for a in ("first_conv", "second_conv"):
        if k.startswith(a + "_"):
            rest = k[len(a) + 1:]
            i, name = rest.split("_", 1)
            return f"{a}.{i}.{name}"
sd = {unmangle(k): v for k, v in sd.items()}
enc = EncoderB200.from_state_dict(sd).cuda()
nb = torch.from_numpy(g["enc_neighborhood"]).cuda()
out = enc(nb)
torch.cuda.synchronize()
want = g["enc_out"]
got = out.cpu().numpy()
print("status", enc.last_status.item(), "max|want|", np.abs(want).max(), "max abs err", np.abs(got - want).max(),
      "mean abs err", np.abs(got - want).mean(), "rel", np.abs(got - want).max() / np.abs(want).max())
print(got[0, 0, :8]); print(want[0, 0, :8])
