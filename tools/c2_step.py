"""A few config-2 loss steps (forward + backward) for `ncu`:  python tools/c2_step.py [loss-class-name] [steps]"""
import pathlib
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import xfmr_b200  # noqa: E402

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "InfomationNoiseContrastiveEstimationLoss"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kw = {}
if len(sys.argv) > 3:
    kw["num_negatives"] = int(sys.argv[3])
inp = bench.make_c2(dev, 0, torch.bfloat16)
m = getattr(xfmr_b200, name)(sigma=5.0, margin=0.5, **kw)
q = inp["user_embed"].detach().requires_grad_(True)
v = inp["item_embed"].detach().requires_grad_(True)
for _ in range(steps):
    loss = m(q, v, inp["target"], item_idx=inp["item_idx"], pos_idx=inp["pos_idx"])
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss))
