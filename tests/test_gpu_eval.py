"""Sparse exclusion lists and batched evaluation against the CPU oracles.

Ids are bit-exact (integer work); metrics are fp32 against the float64 restatement of torchmetrics'
definitions (``oracle/metrics_oracle.py``, unpinned: torchmetrics is not installable here), atol 1e-5.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make(nq: int, n: int, d: int, seed: int) -> tuple[torch.Tensor, torch.Tensor]:
    gen = torch.Generator().manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn(nq, d, generator=gen), dim=-1)
    it = torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=-1)
    return q, it


def exclusion_lists(q: torch.Tensor, it: torch.Tensor, ids: np.ndarray, width: int, n_best: int, seed: int) -> np.ndarray:
    from oracle import native  # noqa: PLC0415

    rng = np.random.default_rng(seed)
    excl = np.full((q.size(0), width), native.PAD_ID, dtype=np.int64)
    full = q.numpy() @ it.numpy().T
    for r in range(q.size(0)):
        m = int(rng.integers(0, n_best + 1))             # ragged: 0..n_best of the best items
        excl[r, :m] = ids[np.argsort(-full[r])[:m]]
        if width > n_best:
            excl[r, n_best] = 10**9 + r                  # an id that is not in the catalog
    return excl


@pytest.mark.parametrize(("nq", "n", "k", "width", "n_best"), [
    (33, 3706, 20, 24, 20),
    (200, 20000, 100, 150, 140),
    (5, 40, 30, 16, 15),         # k + E larger than the catalog: empty slots at the end
])
def test_sparse_exclusions_equal_the_prefilter(nq: int, n: int, k: int, width: int, n_best: int) -> None:
    import xfmr_b200  # noqa: PLC0415
    from oracle import native  # noqa: PLC0415

    q, it = make(nq, n, 64, nq + n)
    ids = (torch.randperm(3 * n, generator=torch.Generator().manual_seed(1))[:n] + 1).numpy()
    excl = exclusion_lists(q, it, ids, width, n_best, 0)
    index = xfmr_b200.ItemProcessor(metric="dot").get_index(it, torch.from_numpy(ids))
    index.DENSE_MASK_BYTES = 0    # force the sparse path; the dense one is covered by test_gpu_topk.py
    scores, got = index.search_batch(q, torch.from_numpy(excl), top_k=k)
    ref_s, ref_i = native.topk(q.numpy(), it.numpy(), k, item_ids=ids, exclude=excl)
    assert np.array_equal(got.cpu().numpy(), ref_i)
    assert np.array_equal(scores.cpu().numpy(), ref_s)
    # and it agrees with the dense-mask path
    dense = xfmr_b200.ItemProcessor(metric="dot").get_index(it, torch.from_numpy(ids))
    s2, i2 = dense.search_batch(q, torch.from_numpy(excl), top_k=k)
    assert torch.equal(i2, got)
    assert torch.equal(s2, scores)


def test_exclusions_too_long_for_either_path_raise() -> None:
    import xfmr_b200  # noqa: PLC0415

    q, it = make(4, 500, 32, 3)
    index = xfmr_b200.ItemProcessor(metric="dot").get_index(it)
    index.DENSE_MASK_BYTES = 0
    with pytest.raises(ValueError, match="exclusion lists"):
        index.search_batch(q, torch.zeros(4, 250, dtype=torch.int64), top_k=20)


def targets_for(q: torch.Tensor, it: torch.Tensor, ids: np.ndarray, seed: int) -> list[dict[int, float]]:
    rng = np.random.default_rng(seed)
    full = q.numpy() @ it.numpy().T
    out = []
    for r in range(q.size(0)):
        kind = r % 5
        if kind == 0:
            out.append({})                                                       # user without targets
            continue
        order = np.argsort(-full[r])
        near = order[rng.choice(min(60, len(ids)), size=int(rng.integers(1, 12)), replace=False)]   # some will be retrieved
        far = rng.choice(len(ids), size=int(rng.integers(0, 8)), replace=False)
        t = {int(ids[j]): float(rng.integers(1, 6)) for j in np.concatenate([near, far])}
        if kind == 1:
            t = {i: 0.0 for i in t}                                              # only irrelevant targets
        if kind == 2:
            t[int(ids[order[0]])] = 0.0                                          # a zero-rated hit
        out.append(t)
    return out


@pytest.mark.parametrize(("nq", "n", "k"), [(64, 3706, 20), (301, 9000, 100), (7, 50, 10)])
def test_batched_evaluation_matches_the_metric_oracle(nq: int, n: int, k: int) -> None:
    import xfmr_b200  # noqa: PLC0415
    from oracle import metrics_oracle, native  # noqa: PLC0415

    q, it = make(nq, n, 64, 7 * nq)
    ids = (torch.randperm(2 * n, generator=torch.Generator().manual_seed(2))[:n] + 1).numpy()
    excl = exclusion_lists(q, it, ids, 8, 6, 1)
    targets = targets_for(q, it, ids, 4)
    index = xfmr_b200.ItemProcessor(metric="dot").get_index(it, torch.from_numpy(ids))
    out = index.evaluate(q, [list(t) for t in targets], [list(t.values()) for t in targets], torch.from_numpy(excl), top_k=k)
    _, ref_i = native.topk(q.numpy(), it.numpy(), k, item_ids=ids, exclude=excl)
    per_query, mean = metrics_oracle.batch_metrics(ref_i.tolist(), targets, k)
    assert xfmr_b200.METRIC_NAMES == metrics_oracle.METRIC_NAMES
    np.testing.assert_allclose(out["per_query"].cpu().numpy(), np.array(per_query), atol=1e-5, rtol=0)
    for m, name in enumerate(xfmr_b200.METRIC_NAMES):
        assert abs(float(out[name]) - mean[m]) < 1e-5, name
    assert float(out["RetrievalHitRate"]) > 0   # the construction guarantees some hits


def test_metrics_known_answers() -> None:
    """Hand-computed case: ranked [7, 3, 9, 4], targets {3: 2, 4: 1, 8: 3}, k = 4."""
    import math  # noqa: PLC0415

    import xfmr_b200  # noqa: PLC0415

    pad = -(2**63)
    ids = torch.tensor([[7, 3, 9, 4], [1, 2, -1, -1]], device="cuda")
    tid = torch.tensor([[3, 4, 8], [5, pad, pad]], device="cuda")
    tval = torch.tensor([[2.0, 1.0, 3.0], [1.0, 0.0, 0.0]], device="cuda")
    per_query, mean = xfmr_b200.retrieval_metrics(ids, tid, tval)
    dcg = 2 / math.log2(3) + 1 / math.log2(5)
    idcg = 3 / math.log2(2) + 2 / math.log2(3) + 1 / math.log2(4)
    want0 = [dcg / idcg, 2 / 3, 2 / 4, (1 / 2 + 2 / 4) / 2, 1.0, 1 / 2]
    np.testing.assert_allclose(per_query[0].cpu().numpy(), want0, atol=1e-6)
    np.testing.assert_allclose(per_query[1].cpu().numpy(), [0.0] * 6, atol=0)
    np.testing.assert_allclose(mean.cpu().numpy(), np.array(want0) / 2, atol=1e-6)


def test_index_bundle_round_trip_and_sharded_load(tmp_path) -> None:  # noqa: ANN001
    """``save`` / ``load`` through the reference's table layout; a 2-way sharded load merges to the same answer."""
    import xfmr_b200  # noqa: PLC0415
    from oracle import native  # noqa: PLC0415

    q, it = make(20, 1500, 32, 77)
    ids = (torch.randperm(5000, generator=torch.Generator().manual_seed(3))[:1500] + 1)
    texts = [f"item {int(i)}" for i in ids]
    index = xfmr_b200.ItemProcessor(id_col="movie_id", text_col="movie_text", metric="dot").get_index(it, ids, texts)
    index.save(tmp_path / "bundle")
    loaded = xfmr_b200.ItemProcessor.load(tmp_path / "bundle")
    assert loaded.id_col == "movie_id" and loaded.item_text == texts
    s1, i1 = loaded.search_batch(q, None, top_k=15)
    ref_s, ref_i = native.topk(q.numpy(), it.numpy(), 15, item_ids=ids.numpy())
    assert np.array_equal(i1.cpu().numpy(), ref_i)
    assert np.array_equal(s1.cpu().numpy(), ref_s)
    frame = loaded.search(q[0].numpy(), None, top_k=3)
    assert frame["movie_text"].tolist() == [f"item {int(i)}" for i in ref_i[0][:3]]
    parts = [xfmr_b200.ItemProcessor.load(tmp_path / "bundle", rank=r, world_size=2).search_batch(q, None, top_k=15) for r in range(2)]
    s2, i2 = xfmr_b200.topk_merge(torch.cat([p[0] for p in parts], dim=1), torch.cat([p[1] for p in parts], dim=1), 15)
    assert torch.equal(i2, i1) and torch.equal(s2, s1)
