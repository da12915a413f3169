"""CPU ORACLE (test infrastructure, not product) — the six ranking metrics of ``update_metrics``.

Reference call site: ``xfmr_rec/lightning.py:149-187`` feeds, per user, the retrieved ``(item id, score)`` list and
the graded ``target`` dict into ``torchmetrics.retrieval`` metrics built with ``top_k`` (``:289-306``).  The
arithmetic lives in torchmetrics 1.8.2 (``uv.lock``), which is neither under ``/root/reference`` nor installed
here: **parity unpinned**; the functions restate torchmetrics' functional definitions
(``retrieval_normalized_dcg``, ``retrieval_recall``, ``retrieval_precision``, ``retrieval_average_precision``,
``retrieval_hit_rate``, ``retrieval_reciprocal_rank``) for a ranked list without score ties, with
``empty_target_action="neg"`` (a user without a relevant target scores 0) and the mean over users that
``RetrievalMetric.compute`` takes.  Plain Python loops on purpose.  ``tests/test_oracle_extras.py`` checks hand-computed
known answers and, for NDCG, scikit-learn's ``ndcg_score`` (same definition) on random cases.
"""

from __future__ import annotations

import math

METRIC_NAMES = (
    "RetrievalNormalizedDCG",
    "RetrievalRecall",
    "RetrievalPrecision",
    "RetrievalMAP",
    "RetrievalHitRate",
    "RetrievalMRR",
)


def query_metrics(ranked_ids: list[int], targets: dict[int, float], k: int) -> list[float]:
    """Metrics at k for one user.  ``ranked_ids``: result list, best first (-1 = empty slot);
    ``targets``: item id -> graded relevance (relevant = value > 0)."""
    ranked = [i for i in ranked_ids[:k]]
    n_rel = sum(1 for v in targets.values() if v > 0)
    if n_rel == 0:
        return [0.0] * 6
    gains = [targets.get(i, 0.0) if i >= 0 else 0.0 for i in ranked]
    rel = [g > 0 for g in gains]
    dcg = sum(g / math.log2(r + 2) for r, g in enumerate(gains))
    ideal = sorted(targets.values(), reverse=True)[:k]
    idcg = sum(g / math.log2(r + 2) for r, g in enumerate(ideal))
    hits = sum(rel)
    ndcg = dcg / idcg if idcg > 0 else 0.0
    recall = hits / n_rel
    precision = hits / k
    if hits:
        seen = 0
        ap = 0.0
        for r, is_rel in enumerate(rel):
            if is_rel:
                seen += 1
                ap += seen / (r + 1)
        ap /= hits
        mrr = 1.0 / (rel.index(True) + 1)
    else:
        ap = 0.0
        mrr = 0.0
    return [ndcg, recall, precision, ap, 1.0 if hits else 0.0, mrr]


def batch_metrics(ranked_ids: list[list[int]], targets: list[dict[int, float]], k: int) -> tuple[list[list[float]], list[float]]:
    per_query = [query_metrics(r, t, k) for r, t in zip(ranked_ids, targets, strict=True)]
    mean = [sum(row[m] for row in per_query) / len(per_query) for m in range(6)]
    return per_query, mean
