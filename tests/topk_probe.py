"""Manual probe: retrieval sweep timing on a mid-sized problem."""
import sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200
from xfmr_b200 import _lib, synthetic
dev = torch.device("cuda:0")
Q, N = 128 * 148, 6_000_000
items = synthetic.make_catalog(N, 128, seed=1, device=dev, dtype=torch.bfloat16)
q = synthetic.make_catalog(Q, 128, seed=2, device=dev, dtype=torch.bfloat16)
xfmr_b200.topk_search(q, items, 100); torch.cuda.synchronize()
_lib.sweep_timing(True)
for _ in range(2): xfmr_b200.topk_search(q, items, 100)
ms, n = _lib.sweep_timing_read(); _lib.sweep_timing(False)
ms /= n
tiles = -(-N // 128)
print(f"Q={Q} N={N}: sweep {ms:.2f} ms, {ms*1e-3*1.965e9/tiles:.0f} cycles/tile, {2*Q*N*128/ms/1e9:.0f} TFLOP/s, {Q/ms*1e3:.0f} q/s")
