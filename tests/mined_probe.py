"""Manual probe: one mined (K=4) forward+backward at config 2, for a per-kernel launch list."""
import sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200, bench
dev = torch.device("cuda:0")
inp = bench.make_c2(dev, 0, torch.bfloat16)
m = xfmr_b200.PairwiseHingeLoss(num_negatives=4, sigma=5.0, margin=0.5)
step = bench.loss_step_fn(m, inp)
for _ in range(3): step()
torch.cuda.synchronize()
