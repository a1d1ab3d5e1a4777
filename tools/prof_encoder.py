"""Launch the tcgen05 Encoder a few times at the pre-training size (for `ncu -k regex:encoder`)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gm3d_b200.encoder import EncoderB200  # noqa: E402

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_next.npz"))
sd = {}
for k in g.files:
    if k.startswith("enc_w_"):
        rest = k[len("enc_w_"):]
        for pre in ("first_conv", "second_conv"):
            if rest.startswith(pre + "_"):
                i, nm = rest[len(pre) + 1:].split("_", 1)
                sd[f"{pre}.{i}.{nm}"] = torch.from_numpy(g[k])
dev = torch.device("cuda", 0)
enc = EncoderB200.from_state_dict(sd).to(dev)
nb = torch.randn(128, 64, 32, 3, device=dev) * 0.08
for _ in range(3):
    out = enc(nb)
torch.cuda.synchronize()
print("ok", enc.last_status.item(), float(out.abs().mean()))
