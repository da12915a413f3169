"""Seeded synthetic workloads shaped like the reference's data.

Used by the tests and by ``bench.py``; there is no network for MovieLens, so shapes and id statistics
are reproduced instead: unit-norm embeddings (the model normalises, ``xfmr_rec/models.py:59``), item ids
``1..n_catalog`` with 0 reserved for padding (``xfmr_rec/data/prepare.py:85``, ``data/load.py:58-72``),
Zipf-distributed in-batch positives (which creates the accidental hits ``losses.py:103`` removes),
uniform negatives without replacement (``data/lightning.py:349-354``), integer ratings 1..5 as targets
(``params.py:8``) and zero-padded ``pos_idx`` rows.
"""

from __future__ import annotations

import torch


def unit_rows(n: int, d: int, gen: torch.Generator, *, normalize: bool = True, scale: float = 1.0) -> torch.Tensor:
    x = torch.randn(n, d, generator=gen, dtype=torch.float32, device=gen.device)
    if normalize:
        x = torch.nn.functional.normalize(x, dim=-1)
    return x * scale


def zipf_ids(n: int, n_catalog: int, gen: torch.Generator, alpha: float = 1.0) -> torch.Tensor:
    """n ids in 1..n_catalog with P(id = r) proportional to r**-alpha."""
    ranks = torch.arange(1, n_catalog + 1, dtype=torch.float64, device=gen.device)
    pmf = ranks.pow(-alpha)
    return torch.multinomial(pmf / pmf.sum(), n, replacement=True, generator=gen) + 1


def zipf_log_pmf(ids: torch.Tensor, n_catalog: int, alpha: float = 1.0) -> torch.Tensor:
    ranks = torch.arange(1, n_catalog + 1, dtype=torch.float64, device=ids.device)
    log_norm = ranks.pow(-alpha).sum().log()
    return (-alpha * ids.to(torch.float64).log() - log_norm).to(torch.float32)


def make_loss_inputs(  # noqa: PLR0913
    batch: int,
    num_items: int,
    dim: int,
    num_pos: int,
    *,
    n_catalog: int | None = None,
    seed: int = 0,
    device: str | torch.device = "cpu",
    normalize: bool = True,
    scale: float = 1.0,
    signed_targets: bool = False,
    mean_extra_pos: float = 16.0,
) -> dict[str, torch.Tensor]:
    """Inputs of one loss call: rows ``0..B-1`` of ``item_embed`` are the in-batch positives."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n_catalog = n_catalog or max(num_items, 2)
    user = unit_rows(batch, dim, gen, normalize=normalize, scale=scale)
    item = unit_rows(num_items, dim, gen, normalize=normalize, scale=scale)
    pos_ids = zipf_ids(batch, n_catalog, gen)
    n_neg = num_items - batch
    if n_neg > 0:
        if n_neg <= n_catalog:
            neg_ids = torch.randperm(n_catalog, generator=gen, device=device)[:n_neg] + 1
        else:
            neg_ids = torch.randint(1, n_catalog + 1, (n_neg,), generator=gen, device=device)
        item_idx = torch.cat([pos_ids, neg_ids])
    else:
        item_idx = pos_ids
    target = torch.randint(1, 6, (batch,), generator=gen, device=device).to(torch.float32)
    if signed_targets:
        flip = torch.rand(batch, generator=gen, device=device)
        target = torch.where(flip < 0.15, -target, target)  # noqa: PLR2004
        target = torch.where((flip >= 0.15) & (flip < 0.2), torch.zeros_like(target), target)  # noqa: PLR2004
    pos_idx = torch.zeros(batch, max(num_pos, 0), dtype=torch.int64, device=device)
    if num_pos > 0:
        pos_idx[:, 0] = pos_ids
        if num_pos > 1:
            p = 1.0 / (1.0 + mean_extra_pos)
            u = torch.rand(batch, generator=gen, device=device).clamp_min(1e-12)
            extra = (u.log() / torch.log1p(torch.tensor(-p, device=device))).floor().to(torch.int64).clamp(0, num_pos - 1)
            cand = torch.randint(1, n_catalog + 1, (batch, num_pos - 1), generator=gen, device=device)
            keep = torch.arange(num_pos - 1, device=device).unsqueeze(0) < extra.unsqueeze(1)
            pos_idx[:, 1:] = torch.where(keep, cand, torch.zeros_like(cand))
    return {
        "user_embed": user,
        "item_embed": item,
        "target": target,
        "item_idx": item_idx.to(torch.int64),
        "pos_idx": pos_idx,
        "log_q": zipf_log_pmf(item_idx, n_catalog),
    }


def make_catalog(num_items: int, dim: int, *, seed: int = 0, device: str | torch.device = "cpu",
                 dtype: torch.dtype = torch.float32, chunk: int = 1 << 22) -> torch.Tensor:
    """Unit-norm catalog generated chunk by chunk on the target device (C5: 25.6 GB per 10^8 x 128 bf16)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = torch.empty(num_items, dim, dtype=dtype, device=device)
    for start in range(0, num_items, chunk):
        stop = min(start + chunk, num_items)
        out[start:stop] = unit_rows(stop - start, dim, gen).to(dtype)
    return out
