"""Hashed ("Bloom") embedding gather — the feeder of the score path.

The reference only cites the technique (``README.md:32-36``; a vestigial ``num_hashes: 2`` at
``xfmr_rec/ray.py:97``); there is no reference code.  Definition used here:

    idx[n, h] = XXH32(le64(ids[n]), seed = seed0 + h)  mod  2**log2_rows
    out[n]    = sum_h table[idx[n, h]]          (fp32 sum, one rounding to bf16)

The integer hashing is bit-exact XXH32 (python-xxhash 3.7.0 is the external oracle).
"""

from __future__ import annotations

import torch

from . import _lib


def hash_indices(ids: torch.Tensor, num_hashes: int, log2_rows: int, *, seed0: int = 0) -> torch.Tensor:
    """``[n, num_hashes]`` int32 table rows for each id (``xb_hash_indices``)."""
    device = _lib.require_cuda(ids)
    ids = ids.to(torch.int64).contiguous()
    with torch.cuda.device(device):
        out = torch.empty(ids.numel(), num_hashes, dtype=torch.int32, device=device)
        status = _lib.lib.xb_hash_indices(
            ids.data_ptr(), ids.numel(), num_hashes, seed0, log2_rows, out.data_ptr(), _lib.stream_ptr(device)
        )
    _lib.check(status, "xb_hash_indices")
    return out


def _check_table(table: torch.Tensor) -> int:
    if table.dtype != torch.bfloat16 or table.dim() != 2:  # noqa: PLR2004
        msg = f"table must be a 2-d bfloat16 tensor, got {table.dtype} {tuple(table.shape)}"
        raise TypeError(msg)
    rows = table.size(0)
    if rows & (rows - 1):
        msg = f"table rows must be a power of two, got {rows}"
        raise ValueError(msg)
    return rows.bit_length() - 1


class _HashGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table: torch.Tensor, ids: torch.Tensor, num_hashes: int, seed0: int) -> torch.Tensor:  # noqa: ANN001
        device = _lib.require_cuda(table, ids)
        log2_rows = _check_table(table)
        ids = ids.to(torch.int64).contiguous()
        table = table.contiguous()
        with torch.cuda.device(device):
            out = torch.empty(ids.numel(), table.size(1), dtype=torch.bfloat16, device=device)
            status = _lib.lib.xb_hash_gather(
                ids.data_ptr(),
                ids.numel(),
                num_hashes,
                seed0,
                table.data_ptr(),
                log2_rows,
                table.size(1),
                out.data_ptr(),
                None,
                _lib.stream_ptr(device),
            )
        _lib.check(status, "xb_hash_gather")
        ctx.save_for_backward(ids)
        ctx.meta = (num_hashes, seed0, log2_rows, table.shape)
        return out

    @staticmethod
    def backward(ctx, d_out: torch.Tensor):  # noqa: ANN001, ANN205
        (ids,) = ctx.saved_tensors
        num_hashes, seed0, log2_rows, shape = ctx.meta
        device = d_out.device
        d_out = d_out.to(torch.bfloat16).contiguous()
        with torch.cuda.device(device):
            d_table = torch.zeros(shape, dtype=torch.float32, device=device)
            status = _lib.lib.xb_hash_scatter_grad(
                ids.data_ptr(),
                ids.numel(),
                num_hashes,
                seed0,
                d_out.data_ptr(),
                log2_rows,
                shape[1],
                d_table.data_ptr(),
                _lib.stream_ptr(device),
            )
        _lib.check(status, "xb_hash_scatter_grad")
        return d_table.to(torch.bfloat16), None, None, None


def hash_embedding_gather(table: torch.Tensor, ids: torch.Tensor, num_hashes: int = 2, *, seed0: int = 0) -> torch.Tensor:
    """``out[n] = sum_h table[hash_h(ids[n])]`` as bf16 ``[n, d]`` (``xb_hash_gather``), differentiable in ``table``."""
    return _HashGather.apply(table, ids, num_hashes, seed0)


class HashEmbeddingBag(torch.nn.Module):
    """A ``2**log2_rows x dim`` bf16 table addressed by ``num_hashes`` XXH32 hashes of an int64 id."""

    def __init__(self, log2_rows: int, dim: int, *, num_hashes: int = 2, seed0: int = 0, init_std: float = 0.02) -> None:
        super().__init__()
        self.num_hashes = num_hashes
        self.seed0 = seed0
        self.weight = torch.nn.Parameter(torch.empty(1 << log2_rows, dim, dtype=torch.bfloat16))
        torch.nn.init.normal_(self.weight, std=init_std)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        shape = ids.shape
        out = hash_embedding_gather(self.weight, ids.reshape(-1), self.num_hashes, seed0=self.seed0)
        return out.reshape(*shape, -1)
