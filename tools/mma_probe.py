"""Manual probe: per-tile cost of the MMA/TMA side (DEBUG mode = trivial epilogue, both MMAs)."""
import sys, pathlib, torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200
from xfmr_b200 import _lib
dev = torch.device("cuda:0")
import os
CFG = [(128 * 148, 128 * 400, 128), (128 * 148, 128 * 400, 64), (128 * 148, 128 * 200, 256)]
if os.environ.get("XB_ONE"): CFG = CFG[:1]
for (R, C, d) in CFG:
    rows = torch.randn(R, d, device=dev).bfloat16(); cols = torch.randn(C, d, device=dev).bfloat16()
    kp = -(-d // 64) * 64
    acc = torch.empty(R, kp, device=dev)
    wsb = _lib.lib.xb_debug_workspace_bytes(R, C, d, 0); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    def run():
        _lib.check(_lib.lib.xb_debug_scores(R, C, d, 1, 0, rows.data_ptr(), cols.data_ptr(), None, acc.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr(dev)), "dbg")
    run(); torch.cuda.synchronize()
    _lib.sweep_timing(True)
    for _ in range(3): run()
    ms, n = _lib.sweep_timing_read(); _lib.sweep_timing(False)
    ms /= n
    tiles_per_cta = C // 128
    print(f"R={R} C={C} d={d}: {ms:.3f} ms per sweep, {ms*1e-3*1.965e9/tiles_per_cta:.0f} cycles/tile (1 CTA/SM, {tiles_per_cta} tiles each), "
          f"{2*2*R*C*d/ms/1e9:.0f} TFLOP/s incl. both MMAs")
