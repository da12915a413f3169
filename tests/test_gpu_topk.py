"""Exact top-k retrieval (``ItemProcessor.search`` semantics) against the C oracle.

fp32 mode: ids bit-exact, ties broken by the lower item id, scores bit-exact (the oracle defines the
summation order).  bf16 mode: recall@k >= 0.999 against the oracle run on the bf16-rounded inputs.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

from helpers import bf16_round

pytestmark = pytest.mark.gpu


def make(nq: int, n: int, d: int, seed: int) -> tuple[torch.Tensor, torch.Tensor]:
    gen = torch.Generator().manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn(nq, d, generator=gen), dim=-1)
    it = torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=-1)
    return q, it


@pytest.mark.parametrize(("nq", "n", "d", "k"), [
    (1, 1, 8, 1),
    (3, 50, 16, 10),        # fewer items than one tile
    (130, 1000, 64, 20),    # TOP_K default of the reference
    (1024, 3706, 64, 10),   # BASELINE config 1
    (64, 20000, 128, 100),  # many column tiles per CTA, k = 100
    (5, 300, 32, 256),      # k larger than... most of the catalog
])
def test_fp32_ids_and_scores_bit_exact(nq: int, n: int, d: int, k: int) -> None:
    import xfmr_b200  # noqa: PLC0415
    from oracle import native  # noqa: PLC0415

    q, it = make(nq, n, d, nq + n)
    ids = torch.arange(1, n + 1, dtype=torch.int64)  # movie_rn style 1-based ids
    k_eff = min(k, 256)
    scores, got = xfmr_b200.topk_search(q.cuda(), it.cuda(), k_eff, item_ids=ids.cuda())
    ref_s, ref_i = native.topk(q.numpy(), it.numpy(), k_eff, item_ids=ids.numpy())
    assert np.array_equal(got.cpu().numpy(), ref_i)
    assert np.array_equal(scores.cpu().numpy(), ref_s)


def test_ties_break_towards_lower_id_and_id_base() -> None:
    import xfmr_b200  # noqa: PLC0415
    from oracle import native  # noqa: PLC0415

    q, it = make(40, 600, 32, 5)
    it[100:130] = it[7]          # 31 identical rows
    it[400] = it[7]
    scores, got = xfmr_b200.topk_search(q.cuda(), it.cuda(), 50, id_base=1000)
    ref_s, ref_i = native.topk(q.numpy(), it.numpy(), 50, id_base=1000)
    assert np.array_equal(got.cpu().numpy(), ref_i)
    assert np.array_equal(scores.cpu().numpy(), ref_s)


def test_exclusions_are_prefiltered() -> None:
    """``exclude_item_ids`` semantics of data/lightning.py:247-252 through the ItemProcessor mirror."""
    import xfmr_b200  # noqa: PLC0415
    from oracle import native  # noqa: PLC0415

    q, it = make(33, 3706, 64, 9)
    ids = torch.randperm(10_000, generator=torch.Generator().manual_seed(1))[:3706] + 1   # arbitrary, unsorted ids
    rng = np.random.default_rng(0)
    excl = np.full((33, 20), native.PAD_ID, dtype=np.int64)
    full = q.numpy() @ it.numpy().T
    for r in range(33):
        top = np.argsort(-full[r])[:12]
        excl[r, :12] = ids.numpy()[top]                     # exclude the 12 best of every query
        excl[r, 12:16] = rng.integers(20_000, 30_000, 4)    # ids that are not in the catalog
    index = xfmr_b200.ItemProcessor(metric="dot").get_index(it, ids)
    scores, got = index.search_batch(q, torch.from_numpy(excl), top_k=20)
    ref_s, ref_i = native.topk(q.numpy(), it.numpy(), 20, item_ids=ids.numpy(), exclude=excl)
    assert np.array_equal(got.cpu().numpy(), ref_i)
    assert np.array_equal(scores.cpu().numpy(), ref_s)
    # single-query DataFrame API of the reference
    frame = index.search(q[0].numpy(), exclude_item_ids=[int(x) for x in excl[0] if x != native.PAD_ID], top_k=20)
    assert list(frame.columns) == ["movie_id", "embedding", "score"]   # lance's to_pandas(): table columns, then the score
    assert np.array_equal(np.stack(frame["embedding"].to_numpy()), it.numpy()[[ids.tolist().index(i) for i in frame["movie_id"]]])
    assert frame["movie_id"].tolist() == ref_i[0].tolist()
    assert frame["score"].is_monotonic_decreasing
    # no exclusions at all (the reference substitutes [0])
    frame0 = index.search(q[0].numpy(), None, top_k=5)
    assert frame0["movie_id"].tolist() == native.topk(q[:1].numpy(), it.numpy(), 5, item_ids=ids.numpy())[1][0].tolist()


def test_bf16_mode_recall() -> None:
    import xfmr_b200  # noqa: PLC0415
    from oracle import native  # noqa: PLC0415

    q, it = make(256, 50_000, 128, 21)
    scores, got = xfmr_b200.topk_search(q.cuda().bfloat16(), it.cuda().bfloat16(), 100)
    ref_s, ref_i = native.topk(bf16_round(q).numpy(), bf16_round(it).numpy(), 100)
    got = got.cpu().numpy()
    hits = sum(len(set(got[r]) & set(ref_i[r])) for r in range(256))
    assert hits / (256 * 100) >= 0.999
    assert np.allclose(scores.cpu().numpy(), ref_s, rtol=1e-5, atol=1e-6)
    assert (np.diff(scores.cpu().numpy(), axis=1) <= 0).all()


def test_merge_is_deterministic_across_shardings() -> None:
    """Row-sharding the catalog G ways and merging gives the single-shard answer for every G."""
    import xfmr_b200  # noqa: PLC0415

    q, it = make(70, 9000, 64, 31)
    it[4000] = it[10]
    qc, itc = q.cuda(), it.cuda()
    base_s, base_i = xfmr_b200.topk_search(qc, itc, 30)
    for shards in (2, 3, 8):
        bounds = np.linspace(0, 9000, shards + 1).astype(int)
        parts = [xfmr_b200.topk_search(qc, itc[a:b].contiguous(), 30, id_base=int(a)) for a, b in zip(bounds[:-1], bounds[1:])]
        cat_s = torch.cat([p[0] for p in parts], dim=1)
        cat_i = torch.cat([p[1] for p in parts], dim=1)
        s, i = xfmr_b200.topk_merge(cat_s, cat_i, 30)
        assert torch.equal(i, base_i)
        assert torch.equal(s, base_s)


def test_search_scores_are_cosine_similarities_by_default() -> None:
    """The reference indexes with ``metric="cosine"`` and reports ``1 - distance`` (data/lightning.py:222-229, :257):
    with un-normalised catalog rows and queries the default ``ItemProcessor`` must rank by cosine similarity, where raw
    inner products rank by norm."""
    import xfmr_b200  # noqa: PLC0415

    gen = torch.Generator().manual_seed(3)
    items = torch.randn(500, 32, generator=gen) * (0.2 + 3.0 * torch.rand(500, 1, generator=gen))
    query = torch.randn(1, 32, generator=gen) * 7.0
    ids = torch.arange(100, 600)
    want = torch.nn.functional.cosine_similarity(query.double(), items.double())
    order = want.argsort(descending=True)[:10]
    frame = xfmr_b200.ItemProcessor().get_index(items, ids).search(query.numpy(), None, top_k=10)
    assert frame["movie_id"].tolist() == ids[order].tolist()
    assert frame["score"].tolist() == pytest.approx(want[order].tolist(), abs=1e-5)
    assert list(frame.columns) == ["movie_id", "embedding", "score"]
    dot = xfmr_b200.ItemProcessor(metric="dot").get_index(items, ids).search(query.numpy(), None, top_k=10)
    raw = (items.double() @ query.double().t()).squeeze(1)
    assert dot["movie_id"].tolist() == ids[raw.argsort(descending=True)[:10]].tolist()


def _brute_force_must_have(q: torch.Tensor, it: torch.Tensor, k: int, allowed: torch.Tensor | None = None):  # noqa: ANN202
    """fp32 scores of the bf16 values; per query the rows that MUST be returned (score above the k-th by a margin that
    covers fp32 accumulation order) and the k-th score itself."""
    sc = q.float() @ it.float().t()
    if allowed is not None:
        sc = sc.masked_fill(~allowed, float("-inf"))
    kk = min(k, it.size(0))
    top_s, _ = sc.topk(kk, dim=1)
    kth = top_s[:, -1:]
    return sc, (sc > kth + 2e-6) & torch.isfinite(sc), kth


@pytest.mark.parametrize(("nq", "n", "d", "k"), [
    (1, 1, 64, 1),            # a single pair
    (33, 300, 64, 7),         # ragged query tile, three item tiles
    (257, 5000, 96, 100),     # d not a multiple of 64 (zero-padded K block), one row into a second query-tile pair
    (700, 128, 128, 128),     # exactly one item tile, k = every item
    (512, 70_000, 128, 256),  # k at the ABI maximum
])
def test_bf16_retrieval_kernel_edge_shapes(nq: int, n: int, d: int, k: int) -> None:
    """The bf16 retrieval kernel (``csrc/sweep_rt.cuh``) on ragged / minimal / maximal shapes: every row that beats the k-th
    score clearly is returned, nothing below the k-th score is, scores are the fp32 dot products of the bf16 values,
    lists are sorted, slots beyond the catalog are (-inf, -1)."""
    import xfmr_b200  # noqa: PLC0415

    q, it = make(nq, n, d, 3 * nq + n)
    q, it = q.cuda().bfloat16(), it.cuda().bfloat16()
    scores, got = xfmr_b200.topk_search(q, it, k)
    sc, must, kth = _brute_force_must_have(q, it, k)
    kk = min(k, n)
    assert bool((got[:, :kk] >= 0).all()) and bool((got[:, kk:] == -1).all())
    assert bool(torch.isinf(scores[:, kk:]).all())
    returned = torch.zeros_like(must)
    returned.scatter_(1, got[:, :kk], True)
    assert bool((returned | ~must).all()), "a row above the k-th score is missing"
    picked = sc.gather(1, got[:, :kk])
    assert bool((picked >= kth - 2e-6).all()), "a row below the k-th score was returned"
    assert torch.allclose(scores[:, :kk], picked, rtol=0, atol=2e-6)
    assert bool((scores[:, 1:kk] <= scores[:, : kk - 1]).all())


def test_bf16_retrieval_kernel_with_exclusion_mask() -> None:
    """bf16 search with per-query exclusion lists through the dense bit mask (``rt_kernel<HAS_MASK>``): no excluded id comes
    back and the allowed rows above the k-th allowed score all do."""
    import xfmr_b200  # noqa: PLC0415

    nq, n, d, k = 130, 3000, 64, 20
    q, it = make(nq, n, d, 77)
    ids = torch.randperm(50_000, generator=torch.Generator().manual_seed(2))[:n] + 1
    sc_all = bf16_round(q) @ bf16_round(it).t()
    excl = torch.full((nq, 30), -1, dtype=torch.int64)
    for r in range(nq):
        excl[r, :25] = ids[sc_all[r].topk(25).indices]           # the 25 best of every query are excluded
        excl[r, 25:28] = torch.tensor([60_001, 60_002, 60_003])  # ids that are not in the catalog
    index = xfmr_b200.ItemProcessor(metric="dot", compute="bf16").get_index(it.bfloat16(), ids)
    scores, got = index.search_batch(q.bfloat16(), excl, top_k=k)
    got_c = got.cpu()
    assert not bool((got_c[:, :, None] == excl[:, None, :]).any())
    allowed = ~(ids[None, :, None] == excl[:, None, :]).any(dim=2)
    sc, must, kth = _brute_force_must_have(q.cuda().bfloat16(), it.cuda().bfloat16(), k, allowed.cuda())
    pos = {int(v): j for j, v in enumerate(ids.tolist())}
    cols = torch.tensor([[pos[int(v)] for v in row] for row in got_c.tolist()], device="cuda")
    returned = torch.zeros_like(must)
    returned.scatter_(1, cols, True)
    assert bool((returned | ~must).all())
    assert bool((sc.gather(1, cols) >= kth - 2e-6).all())
