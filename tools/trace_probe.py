"""Manual probe: per-tile clock trace of CTA (0,0) for the retrieval sweep and the forward loss sweep."""
import sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200, bench
from xfmr_b200 import _lib, synthetic
dev = torch.device("cuda:0")
NT = 4000
trace = torch.zeros(NT, 8, dtype=torch.int64, device=dev)
def show(name, lo, hi):
    t = trace.cpu()
    base = t[lo, 3].item()
    print(name, "tiles", lo, "..", hi, " columns: mma[wait_sempty_start, got_sempty, got_cfull] epi[wait_sfull_start, got_sfull, tile_end] prod[wait_cempty_start, got]")
    for i in range(lo, hi):
        print("  ", i, [int(x - base) for x in t[i].tolist()])
    d = (t[hi - 1, 5] - t[lo, 5]).item() / (hi - 1 - lo)
    print("   mean period", d, "cycles;  epilogue s_full wait mean", float((t[lo:hi, 4] - t[lo:hi, 3]).float().mean()),
          " epilogue busy mean", float((t[lo:hi, 5] - t[lo:hi, 4]).float().mean()),
          " mma: s_empty wait", float((t[lo:hi, 1] - t[lo:hi, 0]).float().mean()), " c_full wait", float((t[lo:hi, 2] - t[lo:hi, 1]).float().mean()),
          " producer c_empty wait", float((t[lo:hi, 7] - t[lo:hi, 6]).float().mean()))
# retrieval
Q, N = 128 * 148, 2_000_000
items = synthetic.make_catalog(N, 128, seed=1, device=dev, dtype=torch.bfloat16)
q = synthetic.make_catalog(Q, 128, seed=2, device=dev, dtype=torch.bfloat16)
xfmr_b200.topk_search(q, items, 100); torch.cuda.synchronize()
_lib.lib.xb_debug_set_trace(trace.data_ptr(), NT)
xfmr_b200.topk_search(q, items, 100); torch.cuda.synchronize()
_lib.lib.xb_debug_set_trace(None, 0)
show("TOPK", 3000, 3012)

# loss sweeps at config 2
inp = bench.make_c2(dev, 0, torch.bfloat16)
m = xfmr_b200.InfomationNoiseContrastiveEstimationLoss(sigma=5.0, margin=0.5)
q2 = inp["user_embed"].detach().requires_grad_(True); v2 = inp["item_embed"].detach().requires_grad_(True)
loss = m(q2, v2, inp["target"], item_idx=inp["item_idx"], pos_idx=inp["pos_idx"]); torch.cuda.synchronize()
trace.zero_()
_lib.lib.xb_debug_set_trace(trace.data_ptr(), 77)
loss = m(q2, v2, inp["target"], item_idx=inp["item_idx"], pos_idx=inp["pos_idx"]); torch.cuda.synchronize()
show("FWD InfoNCE", 40, 52)
trace.zero_()
loss.backward(); torch.cuda.synchronize()
_lib.lib.xb_debug_set_trace(None, 0)
show("GRAD dI (item-major)", 16, 28)
