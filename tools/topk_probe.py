"""Manual probe: retrieval sweep timing + clock trace windows (needs the -DXB_TRACE build for the trace part).
`XB_LIB=.../libxfmr_b200_trace.so python tests/topk_probe.py [Q] [N]`"""
import sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200
from xfmr_b200 import _lib, synthetic
dev = torch.device("cuda:0")
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 128 * 296
NI = int(sys.argv[2]) if len(sys.argv) > 2 else 8_000_000
items = synthetic.make_catalog(NI, 128, seed=1, device=dev, dtype=torch.bfloat16)
q = synthetic.make_catalog(Q, 128, seed=2, device=dev, dtype=torch.bfloat16)
xfmr_b200.topk_search(q, items, 100); torch.cuda.synchronize()
_lib.sweep_timing(True)
for _ in range(2): xfmr_b200.topk_search(q, items, 100)
torch.cuda.synchronize()
t_ms, t_n = _lib.sweep_timing_read(); _lib.sweep_timing(False)
rb = -(-Q // 128); waves = -(-rb // 148)
print(f"TOPK Q={Q} N={NI}: sweep {t_ms / t_n:.2f} ms, {t_ms / t_n * 1e-3 * 1.965e9 / (NI / 128) / waves:.0f} cycles/tile/wave, {2 * Q * NI * 128 / (t_ms / t_n * 1e-3) / 1e12:.0f} TFLOP/s, {Q / (t_ms / t_n * 1e-3):.0f} q/s")
NT = NI // 128
trace = torch.zeros(NT + 16, 8, dtype=torch.int64, device=dev)
_lib.lib.xb_debug_set_trace(trace.data_ptr(), NT)
xfmr_b200.topk_search(q, items, 100); torch.cuda.synchronize()
_lib.lib.xb_debug_set_trace(None, 0)
t = trace.cpu()
if int(t[:, 5].max()) > 0:
    for lo in (16, 200, 1000, 4000, 16000, NT - 200):
        hi = lo + 64
        if hi >= NT: continue
        d = (t[hi - 1, 5] - t[lo, 5]).item() / (hi - 1 - lo)
        print(f"  tiles {lo}..{hi}: period {d:.0f}; epi: s_full wait {float((t[lo:hi, 4] - t[lo:hi, 3]).float().mean()):.0f} busy {float((t[lo:hi, 5] - t[lo:hi, 4]).float().mean()):.0f}")
