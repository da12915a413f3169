"""Manual probe: PairwiseLogisticLoss (BPR) fwd+bwd at config 2, graph replay."""
import pathlib
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import bench
import xfmr_b200

dev = torch.device("cuda:0")
inp = bench.make_c2(dev, 0, torch.bfloat16)
for name in ("PairwiseLogisticLoss", "InfomationNoiseContrastiveEstimationLoss"):
    m = getattr(xfmr_b200, name)(sigma=5.0, margin=0.5)
    step = bench.graphed(bench.loss_step_fn(m, inp))
    flush = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
    ts = bench.timed_steps(step, 20, 3, flush)
    print(name, "ms/step", sorted(ts)[len(ts) // 2])
