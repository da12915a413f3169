"""Manual probe: hash gather at config 4 (4Mi ids, k = 2, 2^22 x 128 bf16 table)."""
import pathlib
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import bench

print(bench.bench_gather(torch.device("cuda:0")))
