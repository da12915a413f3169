"""Times the bf16 retrieval search (library events around the sweep launch):  python tools/rt_time.py [items] [queries] [repeats]"""
import pathlib
import subprocess
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200  # noqa: E402
from xfmr_b200 import _lib, synthetic  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
q = int(sys.argv[2]) if len(sys.argv) > 2 else 65_536
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 5
items = synthetic.make_catalog(n, 128, seed=100, device=dev, dtype=torch.bfloat16)
queries = synthetic.make_catalog(q, 128, seed=7, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    xfmr_b200.topk_search(queries, items, 100)
torch.cuda.synchronize()
_lib.sweep_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(rep):
    s, i = xfmr_b200.topk_search(queries, items, 100)
e1.record()
torch.cuda.synchronize()
ms, cnt = _lib.sweep_timing_read()
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True,
                     check=False).stdout.strip()
tot = e0.elapsed_time(e1) / rep
print(f"{n} items x {q} queries: {tot:.2f} ms/search ({q / tot:.1f} k q/s), sweep {ms / max(cnt, 1):.2f} ms, "
      f"{2.0 * q * n * 128 / (ms / max(cnt, 1) * 1e-3) / 1e12:.0f} TF/s; after: {clk}; checksum {float(s.sum()):.3f} {int(i.sum())}")
