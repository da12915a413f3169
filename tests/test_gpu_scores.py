"""The tcgen05 score tile and the second (gradient) MMA against dense float64 matmuls."""

from __future__ import annotations

import pytest
import torch

pytestmark = pytest.mark.gpu


def debug_scores(rows: torch.Tensor, cols: torch.Tensor, compute: int) -> tuple[torch.Tensor, torch.Tensor]:
    from xfmr_b200 import _lib  # noqa: PLC0415

    dev = rows.device
    nr, d = rows.shape
    nc = cols.shape[0]
    kp = -(-d // 64) * 64
    rp, cp = -(-nr // 128) * 128, -(-nc // 128) * 128
    s = torch.full((rp, cp), float("nan"), device=dev)
    acc = torch.full((rp, kp), float("nan"), device=dev)
    wsb = _lib.lib.xb_debug_workspace_bytes(nr, nc, d, compute)
    assert wsb > 0, _lib.lib.xb_last_error_string()
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    status = _lib.lib.xb_debug_scores(nr, nc, d, _lib.dtype_code(rows.dtype), compute, rows.data_ptr(), cols.data_ptr(),
                                      s.data_ptr(), acc.data_ptr(), ws.data_ptr(), wsb, _lib.stream_ptr(dev))
    _lib.check(status, "xb_debug_scores")
    torch.cuda.synchronize()
    return s, acc


CASES = [
    (128, 128, 64, torch.bfloat16, 0),
    (1, 1, 8, torch.bfloat16, 0),          # smallest possible: one row, one column, d < 64
    (200, 1000, 128, torch.bfloat16, 0),   # ragged rows and columns
    (130, 300, 48, torch.float32, 1),      # split-bf16, d not a multiple of 64
    (256, 640, 256, torch.bfloat16, 0),    # widest embedding
    (128, 384, 128, torch.float32, 1),     # split-bf16 at the shared-memory limit
    (384, 4096, 32, torch.float32, 0),     # fp32 inputs rounded to bf16
]


@pytest.mark.parametrize(("nr", "nc", "d", "dtype", "compute"), CASES)
def test_score_tile_and_grad_mma(nr: int, nc: int, d: int, dtype: torch.dtype, compute: int) -> None:
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(nr * 7 + nc)
    rows = torch.randn(nr, d, device=dev, generator=gen).to(dtype)
    cols = torch.randn(nc, d, device=dev, generator=gen).to(dtype)
    s, acc = debug_scores(rows, cols, compute)

    def seen(x: torch.Tensor) -> torch.Tensor:  # the values the tensor cores see
        hi = x.to(torch.bfloat16)
        if compute == 1:
            lo = (x.float() - hi.float()).to(torch.bfloat16)
            return hi.double() + lo.double()
        return hi.double()

    ref = seen(rows) @ seen(cols).T
    got = s[:nr, :nc].double()
    assert not torch.isnan(got).any()
    scale = ref.abs().max().item() + 1e-6
    if compute == 1:
        # dropped lo*lo term: ~2^-16 relative per product
        exact = rows.double() @ cols.double().T
        assert (got - exact).abs().max().item() < 3e-5 * scale
    else:
        assert (got - ref).abs().max().item() < 2e-6 * scale  # fp32 accumulation of exact bf16 products
    # second MMA: acc = bf16(S) @ cols(as seen)
    g = got.float().to(torch.bfloat16).double()
    ref_acc = g @ seen(cols)
    got_acc = acc[:nr, :d].double()
    assert not torch.isnan(got_acc).any()
    assert (got_acc - ref_acc).abs().max().item() < 1e-5 * (ref_acc.abs().max().item() + 1e-6)
    # padded embedding columns of the accumulator are exactly zero
    kp = acc.shape[1]
    if kp > d:
        assert (acc[:nr, d:] == 0).all()
