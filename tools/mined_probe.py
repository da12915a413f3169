"""Manual probe: a mined loss (default PairwiseHingeLoss, num_negatives=4: the reference's default training loss) at config 2.
Prints graph-replay and eager step times, the summed sweep time per step (library events) and the SM clock under load.
    python tools/mined_probe.py [loss-class] [K] [mining]"""
import pathlib
import subprocess
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import xfmr_b200  # noqa: E402
from xfmr_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
inp = bench.make_c2(dev, 0, torch.bfloat16)
name = sys.argv[1] if len(sys.argv) > 1 else "PairwiseHingeLoss"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 4
kw = {"mining": sys.argv[3]} if len(sys.argv) > 3 else {}
m = getattr(xfmr_b200, name)(num_negatives=k, sigma=5.0, margin=0.5, **kw)
step = bench.loss_step_fn(m, inp)
for _ in range(3):
    step()
torch.cuda.synchronize()
eager = bench.timed_steps(step, 10, 0, None)
_lib.sweep_timing(True)
for _ in range(10):
    step()
torch.cuda.synchronize()
sweep_ms, n_sweeps = _lib.sweep_timing_read()
_lib.sweep_timing(False)
replay = bench.graphed(step)
ts = bench.timed_steps(replay, 20, 3, None)
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader"],
                     capture_output=True, text=True, check=False).stdout.strip()
print(f"{name} K={k} {kw}: graph {sum(ts) / len(ts):.4f} ms/step, eager {sum(eager) / len(eager):.4f}, "
      f"sweeps {sweep_ms / 10:.4f} ms/step in {n_sweeps / 10:.0f} launches; clocks after: {clk}")
