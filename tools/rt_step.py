"""A retrieval search for `ncu`:  python tools/rt_step.py [num_items] [num_queries] [repeats]"""
import pathlib
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200  # noqa: E402
from xfmr_b200 import synthetic  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
q = int(sys.argv[2]) if len(sys.argv) > 2 else 65_536
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 2
items = synthetic.make_catalog(n, 128, seed=100, device=dev, dtype=torch.bfloat16)
queries = synthetic.make_catalog(q, 128, seed=7, device=dev, dtype=torch.bfloat16)
for _ in range(rep):
    s, i = xfmr_b200.topk_search(queries, items, 100)
torch.cuda.synchronize()
print("ok", float(s[0, 0]), int(i[0, 0]))
