"""Manual probe (not a pytest file): per-sweep timings at config 2 + a short retrieval sweep + clock trace of CTA (0,0).
`python tests/quick_probe.py [loss-class-name]`"""
import sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200, bench
from xfmr_b200 import _lib, synthetic
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "InfomationNoiseContrastiveEstimationLoss"
inp = bench.make_c2(dev, 0, torch.bfloat16)
m = getattr(xfmr_b200, name)(sigma=5.0, margin=0.5)
def fwd():
    q = inp["user_embed"].detach().requires_grad_(True); v = inp["item_embed"].detach().requires_grad_(True)
    return m(q, v, inp["target"], item_idx=inp["item_idx"], pos_idx=inp["pos_idx"])
for _ in range(3): fwd().backward()
torch.cuda.synchronize()
N = 20
_lib.sweep_timing(True)
for _ in range(N): fwd()
torch.cuda.synchronize()
f_ms, f_n = _lib.sweep_timing_read(); _lib.sweep_timing(False)
_lib.sweep_timing(True)
for _ in range(N): fwd().backward()
torch.cuda.synchronize()
a_ms, a_n = _lib.sweep_timing_read(); _lib.sweep_timing(False)
print(f"{name}: fwd sweep {f_ms / N * 1e3:.1f} us ({f_n // N} launches)  bwd sweeps {(a_ms - f_ms) / N * 1e3:.1f} us  total {a_ms / N * 1e3:.1f} us"
      f"  -> {275.5186 / (a_ms / N):.1f} TFLOP/s algorithmic")
# retrieval
Q, NI = 128 * 148, 2_000_000
items = synthetic.make_catalog(NI, 128, seed=1, device=dev, dtype=torch.bfloat16)
q = synthetic.make_catalog(Q, 128, seed=2, device=dev, dtype=torch.bfloat16)
xfmr_b200.topk_search(q, items, 100); torch.cuda.synchronize()
_lib.sweep_timing(True)
for _ in range(3): xfmr_b200.topk_search(q, items, 100)
torch.cuda.synchronize()
t_ms, t_n = _lib.sweep_timing_read(); _lib.sweep_timing(False)
print(f"TOPK Q={Q} N={NI}: sweep {t_ms / t_n:.2f} ms, {t_ms / t_n * 1e-3 * 1.965e9 / (NI / 128):.0f} cycles/tile, {2 * Q * NI * 128 / (t_ms / t_n * 1e-3) / 1e12:.0f} TFLOP/s")
# trace
NT = 4000
trace = torch.zeros(NT * 2, 8, dtype=torch.int64, device=dev)
def show(nm, lo, hi):
    t = trace.cpu()
    d = (t[hi - 1, 5] - t[lo, 5]).item() / (hi - 1 - lo)
    print(f"{nm}: period {d:.0f}; epi: s_full wait {float((t[lo:hi, 4] - t[lo:hi, 3]).float().mean()):.0f} busy {float((t[lo:hi, 5] - t[lo:hi, 4]).float().mean()):.0f};"
          f" mma: s_empty wait {float((t[lo:hi, 1] - t[lo:hi, 0]).float().mean()):.0f} c_full wait {float((t[lo:hi, 2] - t[lo:hi, 1]).float().mean()):.0f}"
          f" issue+rest {float((t[lo + 1:hi + 1, 0] - t[lo:hi, 2]).float().mean()):.0f};"
          f" epi detail: unit0 {float((t[lo:hi, 7] - t[lo:hi, 4]).float().mean()):.0f} unit1 {float((t[lo:hi, 5] - t[lo:hi, 7]).float().mean()):.0f}"
          f" st_wait+arrive {float((t[lo:hi, 6] - t[lo:hi, 5]).float().mean()):.0f} next-tile-top {float((t[lo + 1:hi + 1, 3] - t[lo:hi, 6]).float().mean()):.0f}")
_lib.lib.xb_debug_set_trace(trace.data_ptr(), NT)
xfmr_b200.topk_search(q, items, 100); torch.cuda.synchronize()
_lib.lib.xb_debug_set_trace(None, 0)
show("TOPK", 3000, 3012)
trace.zero_(); _lib.lib.xb_debug_set_trace(trace.data_ptr(), 77)
loss = fwd(); torch.cuda.synchronize()
show("FWD", 40, 52)
trace.zero_(); _lib.lib.xb_debug_set_trace(trace.data_ptr(), 200)
loss.backward(); torch.cuda.synchronize()
_lib.lib.xb_debug_set_trace(None, 0)
show("GRAD dI", 16, 28)
tt = trace.cpu()
print("GRAD dI tile-end stamps of CTA(0,0), every 8th of the last block's tiles:", [int(x - tt[0, 3]) for x in tt[0:32:4, 5]])
c = tt[200]
print(f"GRAD dI CTA(0,0): launch->setup {int(c[1]-c[0])}, ->first scores {int(c[2]-c[1])}, to last block's loop end {int(c[3]-c[2])}, last write-out {int(c[4]-c[3])}, total {int(c[5]-c[0])}")

