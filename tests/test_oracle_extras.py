"""CPU checks of the oracles that have no reference code to be pinned against (uniformity family, ranking
metrics): internal consistency and hand-computed known answers."""

from __future__ import annotations

import math

import pytest
import torch


def test_uniformity_oracle_equals_the_ordered_pair_form() -> None:
    from oracle import losses_oracle  # noqa: PLC0415

    x = torch.nn.functional.normalize(torch.randn(50, 16, dtype=torch.float64, generator=torch.Generator().manual_seed(0)), dim=-1)
    d2 = torch.cdist(x, x).pow(2)
    off = ~torch.eye(50, dtype=torch.bool)
    for t in (0.5, 2.0):
        want = (d2[off] * -t).exp().mean().log()
        assert abs(float(losses_oracle.uniformity(x, t)) - float(want)) < 1e-9   # cdist's sqrt-then-square residual
    # two antipodal unit vectors: |x - y|^2 = 4
    pair = torch.tensor([[1.0, 0.0], [-1.0, 0.0]], dtype=torch.float64)
    assert abs(float(losses_oracle.uniformity(pair, 2.0)) + 8.0) < 1e-12


def test_directau_and_mawu_oracles_agree_without_margins() -> None:
    from oracle import losses_oracle  # noqa: PLC0415

    gen = torch.Generator().manual_seed(1)
    q = torch.nn.functional.normalize(torch.randn(20, 8, dtype=torch.float64, generator=gen), dim=-1)
    v = torch.nn.functional.normalize(torch.randn(30, 8, dtype=torch.float64, generator=gen), dim=-1)
    t = torch.ones(20, dtype=torch.float64)
    a = losses_oracle.directau(q, v, t, gamma=1.0)
    b = losses_oracle.mawu(q, v, t, gamma_user=0.5, gamma_item=0.5)
    assert abs(float(a) - float(b)) < 1e-12
    zero = torch.zeros(20, dtype=torch.float64)
    c = losses_oracle.mawu(q, v, t, gamma_user=0.5, gamma_item=0.5, user_margin=zero, item_margin=zero)
    assert abs(float(a) - float(c)) < 1e-6   # 2 - 2 cos(theta) = |u - v|^2 for unit rows (acos clamp residual)


def test_metric_oracle_known_answers() -> None:
    from oracle import metrics_oracle  # noqa: PLC0415

    got = metrics_oracle.query_metrics([7, 3, 9, 4], {3: 2.0, 4: 1.0, 8: 3.0}, 4)
    dcg = 2 / math.log2(3) + 1 / math.log2(5)
    idcg = 3 / math.log2(2) + 2 / math.log2(3) + 1 / math.log2(4)
    want = [dcg / idcg, 2 / 3, 2 / 4, (1 / 2 + 2 / 4) / 2, 1.0, 1 / 2]
    assert got == pytest.approx(want, abs=1e-12)
    assert metrics_oracle.query_metrics([1, 2], {}, 2) == [0.0] * 6                # no targets
    assert metrics_oracle.query_metrics([1, 2], {1: 0.0}, 2) == [0.0] * 6          # nothing relevant
    assert metrics_oracle.query_metrics([1, 2, -1], {1: 1.0, 2: 1.0}, 3) == pytest.approx([1.0, 1.0, 2 / 3, 1.0, 1.0, 1.0])
    per_query, mean = metrics_oracle.batch_metrics([[1], [2]], [{1: 1.0}, {1: 1.0}], 1)
    assert mean == pytest.approx([0.5] * 6)


def test_new_workspace_queries_need_no_gpu() -> None:
    import ctypes  # noqa: PLC0415

    from xfmr_b200 import _lib  # noqa: PLC0415

    ok = _lib.UniformityDesc(n=4096, dim=128, in_dtype=1, compute=0, t=2.0, reserved=0)
    assert _lib.lib.xb_uniformity_workspace_bytes(ctypes.byref(ok)) > 0
    for bad in (_lib.UniformityDesc(n=1, dim=128, in_dtype=1, compute=0, t=2.0, reserved=0),
                _lib.UniformityDesc(n=64, dim=128, in_dtype=1, compute=0, t=0.0, reserved=0),
                _lib.UniformityDesc(n=64, dim=512, in_dtype=1, compute=0, t=2.0, reserved=0)):
        assert _lib.lib.xb_uniformity_workspace_bytes(ctypes.byref(bad)) == 0


def test_metric_oracle_ndcg_matches_scikit_learn() -> None:
    """torchmetrics is not installable here; scikit-learn's ``ndcg_score`` (linear gain, log2 discount, top-k
    truncation - the same definition) pins the NDCG column of the oracle on random cases."""
    import numpy as np  # noqa: PLC0415

    sklearn_metrics = pytest.importorskip("sklearn.metrics")
    from oracle import metrics_oracle  # noqa: PLC0415

    rng = np.random.default_rng(0)
    for _ in range(50):
        n_docs = int(rng.integers(5, 40))
        k = int(rng.integers(1, 12))
        doc_ids = rng.permutation(1000)[:n_docs]
        relevance = rng.integers(0, 6, n_docs).astype(float) * (rng.random(n_docs) < 0.5)
        if relevance.sum() == 0:
            relevance[0] = 3.0
        scores = rng.standard_normal(n_docs)                       # distinct with probability 1: no ties
        order = np.argsort(-scores)
        ranked = doc_ids[order][:k].tolist()
        targets = {int(d): float(r) for d, r in zip(doc_ids, relevance) if r > 0}
        # documents that are not targets but are ranked count as irrelevant: the oracle sees them only in `ranked`
        got = metrics_oracle.query_metrics(ranked, targets, k)
        want = sklearn_metrics.ndcg_score(relevance[None, :], scores[None, :], k=k)
        assert got[0] == pytest.approx(want, abs=1e-12)
        hits = sum(1 for d in ranked if d in targets)
        assert got[1] == pytest.approx(hits / len(targets))
        assert got[2] == pytest.approx(hits / k)
