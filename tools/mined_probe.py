"""Manual probe: the reference's default training loss (PairwiseHingeLoss, num_negatives=4) at config 2."""
import sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200, bench
dev = torch.device("cuda:0")
inp = bench.make_c2(dev, 0, torch.bfloat16)
name = sys.argv[1] if len(sys.argv) > 1 else "PairwiseHingeLoss"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
m = getattr(xfmr_b200, name)(num_negatives=k, sigma=5.0, margin=0.5)
step = bench.loss_step_fn(m, inp)
for _ in range(3): step()
torch.cuda.synchronize()
ts = bench.timed_steps(step, 10, 0, None)
print(name, "K", k, "ms/step", sum(ts) / len(ts))
