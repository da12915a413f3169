"""Manual probe: one retrieval call at the 8-GPU shard size of config 5 (for an ncu launch list)."""
import sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200
from xfmr_b200 import synthetic
dev = torch.device("cuda:0")
items = synthetic.make_catalog(12_500_000, 128, seed=1, device=dev, dtype=torch.bfloat16)
q = synthetic.make_catalog(65536, 128, seed=2, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    s, i = xfmr_b200.topk_search(q, items, 100)
    m = xfmr_b200.topk_merge(torch.cat([s] * 8, dim=1), torch.cat([i] * 8, dim=1), 100)
torch.cuda.synchronize()
