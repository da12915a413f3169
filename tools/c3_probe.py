"""Manual probe: config 3 on one rank (8,192 users x 24,576 candidates, d=256 bf16), eager steps for a launch list."""
import sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200, bench
from xfmr_b200 import synthetic
dev = torch.device("cuda:0")
b, u, d = 8192, 16384, 256
inp = synthetic.make_loss_inputs(b, b + u, d, 32, n_catalog=200_000, seed=50)
inp = {k: (v.to(dev, torch.bfloat16) if v.is_floating_point() and k.endswith("embed") else v.to(dev)) for k, v in inp.items()}
m = xfmr_b200.InfomationNoiseContrastiveEstimationLoss(sigma=5.0, margin=0.5)
step = bench.loss_step_fn(m, inp)
for _ in range(3): step()
torch.cuda.synchronize()
ts = bench.timed_steps(step, 10, 0, None)
print("C3 eager ms/step", sum(ts) / len(ts))
