"""Manual GPU probe (not a pytest file): validates the two tcgen05 paths of the sweep kernel against
dense matmuls and prints error statistics.  `python tests/gpu_probe.py`"""
import ctypes
import sys
import pathlib

import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import xfmr_b200 as xb  # noqa: E402
from xfmr_b200 import _lib  # noqa: E402


def debug_scores(rows, cols, compute):
    dev = rows.device
    R, d = rows.shape
    C = cols.shape[0]
    kp = -(-d // 64) * 64
    rp, cp = -(-R // 128) * 128, -(-C // 128) * 128
    s = torch.full((rp, cp), float("nan"), device=dev)
    acc = torch.full((rp, kp), float("nan"), device=dev)
    wsb = _lib.lib.xb_debug_workspace_bytes(R, C, d, compute)
    assert wsb > 0, _lib.lib.xb_last_error_string()
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = _lib.lib.xb_debug_scores(R, C, d, _lib.dtype_code(rows.dtype), compute, rows.data_ptr(), cols.data_ptr(),
                                  s.data_ptr(), acc.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr(dev))
    _lib.check(st, "xb_debug_scores")
    torch.cuda.synchronize()
    return s, acc


def main():
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0), _lib.lib.xb_version())
    torch.manual_seed(0)
    for (R, C, d, dtype, compute) in [(128, 128, 64, torch.bfloat16, 0), (128, 256, 128, torch.bfloat16, 0),
                                      (200, 1000, 128, torch.bfloat16, 0), (130, 300, 48, torch.float32, 1),
                                      (256, 640, 256, torch.bfloat16, 0), (128, 384, 128, torch.float32, 1)]:
        rows = torch.randn(R, d, device=dev).to(dtype)
        cols = torch.randn(C, d, device=dev).to(dtype)
        s, acc = debug_scores(rows, cols, compute)
        ref = rows.double() @ cols.double().T
        got = s[:R, :C].double()
        err = (got - ref).abs().max().item()
        # second MMA: acc = bf16(S) @ cols (hi [+ lo])
        g = got.float().to(torch.bfloat16).double()
        if compute == 1:
            hi = cols.to(torch.bfloat16)
            lo = (cols - hi.float()).to(torch.bfloat16)
            cv = hi.double() + lo.double()
        else:
            cv = cols.double()
        ref_acc = g @ cv
        got_acc = acc[:R, :d].double()
        err2 = (got_acc - ref_acc).abs().max().item()
        print(f"R={R} C={C} d={d} {dtype} compute={compute}: S max|err|={err:.3e} (|ref|max={ref.abs().max():.2f}) "
              f"nan={torch.isnan(got).sum().item()}  ACC max|err|={err2:.3e} (|ref|max={ref_acc.abs().max():.1f}) "
              f"nan={torch.isnan(got_acc).sum().item()}")
        if err > 1e-2 * ref.abs().max().item():
            # locate the error pattern
            bad = ((got - ref).abs() > 1e-2 * ref.abs().max()).nonzero()
            print("  first bad S entries:", bad[:8].tolist(), " count", bad.shape[0])
        if err2 > 1e-2 * ref_acc.abs().max().item():
            bad = ((got_acc - ref_acc).abs() > 1e-2 * ref_acc.abs().max()).nonzero()
            print("  first bad ACC entries:", bad[:8].tolist(), " count", bad.shape[0])


if __name__ == "__main__":
    main()
