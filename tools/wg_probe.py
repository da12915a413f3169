"""Manual probe (not a test): clock-stamp trace of CTA 0 of the warpgroup-per-tile sweeps at config 2.
Needs the trace build:  make -C matrix-factorization-torch_b200/csrc VARIANT=_trace XB_EXTRA=-DXB_TRACE -j8
Run:  XB_LIB=matrix-factorization-torch_b200/libxfmr_b200_trace.so python tools/wg_probe.py [loss-class-name]
Columns per virtual tile: mmaA[wait start, got s_empty+c_full, issued] epi[wait s_full start, got, units done, arrived] mmaB[issue]"""
import pathlib
import sys

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import xfmr_b200  # noqa: E402
from xfmr_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "InfomationNoiseContrastiveEstimationLoss"
inp = bench.make_c2(dev, 0, torch.bfloat16)
m = getattr(xfmr_b200, name)(sigma=5.0, margin=0.5)
q = inp["user_embed"].detach().requires_grad_(True)
v = inp["item_embed"].detach().requires_grad_(True)


def fwd():
    return m(q, v, inp["target"], item_idx=inp["item_idx"], pos_idx=inp["pos_idx"])


for _ in range(2):
    fwd().backward()
torch.cuda.synchronize()
NT = 160
trace = torch.zeros(NT, 8, dtype=torch.int64, device=dev)


def show(label, lo, hi):
    t = trace.cpu()
    base = int(t[lo, 0])
    print(label)
    for i in range(lo, hi):
        print("  ", i, [int(x) - base if int(x) else 0 for x in t[i].tolist()])
    per = (int(t[hi - 1, 7]) - int(t[lo, 7])) / (hi - 1 - lo)
    ewait = float((t[lo:hi, 4] - t[lo:hi, 3]).float().mean())
    ebusy = float((t[lo:hi, 5] - t[lo:hi, 4]).float().mean())
    earr = float((t[lo:hi, 6] - t[lo:hi, 5]).float().mean())
    s2e = float((t[lo:hi, 4] - t[lo:hi, 2]).float().mean())
    g2b = float((t[lo:hi, 7] - t[lo:hi, 6]).float().mean())
    awt = float((t[lo:hi, 1] - t[lo:hi, 0]).float().mean())
    ais = float((t[lo:hi, 2] - t[lo:hi, 1]).float().mean())
    print(f"   period {per:.0f}; epilogue: wait s_full {ewait:.0f}, units {ebusy:.0f}, st_wait+arrive {earr:.0f}; S issued -> epilogue sees it {s2e:.0f};"
          f" G arrived -> second MMA issued {g2b:.0f}; mmaA wait {awt:.0f} issue {ais:.0f}")


_lib.lib.xb_debug_set_trace(trace.data_ptr(), NT)
loss = fwd()
torch.cuda.synchronize()
show("merged forward + dQ", 40, 58)
trace.zero_()
loss.backward()
torch.cuda.synchronize()
_lib.lib.xb_debug_set_trace(None, 0)
show("dI", 40, 58)
show("dI around a row-block boundary", 60, 72)
