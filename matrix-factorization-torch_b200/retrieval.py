"""Exact top-k retrieval with the call shape of ``ItemProcessor.search``.

Reference: ``xfmr_rec/data/lightning.py:182-259``.  The reference builds a LanceDB ``IVF_HNSW_PQ``
cosine index (``get_index``, :182-235) and answers one query at a time with an approximate search,
removing ``exclude_item_ids`` before ranking (``prefilter=True``, :247-252) and reporting
``score = 1 - cosine_distance`` (:257).  Here the index is the embedding matrix itself, resident in HBM
in the layout the TMA descriptors read, and ``search`` is an exact brute-force scan on the tensor cores:
same arguments, same result columns, batched over queries.

Ordering contract: score descending, ties broken by the lower item id.
"""

from __future__ import annotations

import ctypes
from typing import TYPE_CHECKING

import torch

from . import _lib

if TYPE_CHECKING:
    from collections.abc import Sequence

    import pandas as pd

TOP_K = 20  # xfmr_rec/params.py:21
_PAD_ID = -(2**63)  # never a real id; ignored by the mask builder


def build_pair_mask(
    col_ids: torch.Tensor,
    row_id_lists: torch.Tensor | None,
    *,
    row_ids0: torch.Tensor | None = None,
    num_rows: int | None = None,
    transpose: bool = False,
) -> tuple[torch.Tensor, torch.Tensor | None]:
    """Bit (r, c) set iff ``col_ids[c] == row_ids0[r]`` or ``col_ids[c] in row_id_lists[r]``.

    Returns ``(mask, mask_t)`` as uint32 bit matrices padded to 128 rows / 128 columns
    (``xb_build_pair_mask``).  This is the complement of ``negative_masks`` (losses.py:92-110) and the
    ``NOT IN`` prefilter of ``search`` (data/lightning.py:247-252).
    """
    device = _lib.require_cuda(col_ids, row_id_lists, row_ids0)
    num_cols = col_ids.numel()
    if row_id_lists is not None:
        num_rows = row_id_lists.size(0)
        list_len = row_id_lists.size(1)
        row_id_lists = row_id_lists.to(torch.int64).contiguous()
    else:
        list_len = 0
    if row_ids0 is not None:
        num_rows = row_ids0.numel()
        row_ids0 = row_ids0.to(torch.int64).contiguous()
    if num_rows is None:
        msg = "num_rows is unknown: give row_id_lists, row_ids0 or num_rows"
        raise ValueError(msg)
    col_ids = col_ids.to(torch.int64).contiguous()
    rows_pad = -(-num_rows // 128) * 128
    cols_pad = -(-num_cols // 128) * 128
    words = _lib.lib.xb_mask_words(num_cols)
    words_t = _lib.lib.xb_mask_words(num_rows)
    with torch.cuda.device(device):
        mask = torch.empty(rows_pad, words, dtype=torch.int32, device=device)
        mask_t = torch.empty(cols_pad, words_t, dtype=torch.int32, device=device) if transpose else None
        ws_bytes = int(_lib.lib.xb_pair_mask_workspace_bytes(num_cols))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        status = _lib.lib.xb_build_pair_mask(
            num_rows,
            num_cols,
            list_len,
            col_ids.data_ptr(),
            _lib.ptr(row_ids0),
            _lib.ptr(row_id_lists) if list_len else None,
            mask.data_ptr(),
            _lib.ptr(mask_t),
            ws.data_ptr(),
            ws_bytes,
            _lib.stream_ptr(device),
        )
    _lib.check(status, "xb_build_pair_mask")
    return mask, mask_t


def topk_search(
    queries: torch.Tensor,
    items: torch.Tensor,
    k: int,
    *,
    item_ids: torch.Tensor | None = None,
    id_base: int = 0,
    excl_mask: torch.Tensor | None = None,
    compute: str | None = None,
) -> tuple[torch.Tensor, torch.Tensor]:
    """``(scores [Q, k] fp32, ids [Q, k] int64)`` of the k best catalog rows per query (``xb_topk_search``)."""
    device = _lib.require_cuda(queries, items, item_ids, excl_mask)
    if queries.dim() != 2 or items.dim() != 2 or queries.size(1) != items.size(1):  # noqa: PLR2004
        msg = f"queries {tuple(queries.shape)} and items {tuple(items.shape)} must be [Q, d] and [N, d]"
        raise ValueError(msg)
    if items.dtype != queries.dtype:
        queries = queries.to(items.dtype)
    queries = queries.contiguous()
    items = items.contiguous()
    desc = _lib.TopkDesc(
        num_queries=queries.size(0),
        num_items=items.size(0),
        dim=queries.size(1),
        k=k,
        in_dtype=_lib.dtype_code(items.dtype),
        compute=_lib.compute_code(compute, items.dtype),
        has_exclusions=int(excl_mask is not None),
        reserved=0,
        id_base=id_base,
    )
    ws_bytes = int(_lib.lib.xb_topk_workspace_bytes(ctypes.byref(desc)))
    if ws_bytes == 0:
        raise _lib.XbError("xb_topk_workspace_bytes: " + _lib.lib.xb_last_error_string().decode())
    if item_ids is not None:
        item_ids = item_ids.to(torch.int64).contiguous()
    with torch.cuda.device(device):
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        scores = torch.empty(queries.size(0), k, dtype=torch.float32, device=device)
        ids = torch.empty(queries.size(0), k, dtype=torch.int64, device=device)
        status = _lib.lib.xb_topk_search(
            ctypes.byref(desc),
            queries.data_ptr(),
            items.data_ptr(),
            _lib.ptr(item_ids),
            _lib.ptr(excl_mask),
            scores.data_ptr(),
            ids.data_ptr(),
            ws.data_ptr(),
            ws_bytes,
            _lib.stream_ptr(device),
        )
    _lib.check(status, "xb_topk_search")
    return scores, ids


def topk_merge(scores: torch.Tensor, ids: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Merge ``[Q, L]`` candidate (score, id) lists into the k best by (score desc, id asc)."""
    device = _lib.require_cuda(scores, ids)
    scores = scores.to(torch.float32).contiguous()
    ids = ids.to(torch.int64).contiguous()
    num_queries, length = scores.shape
    with torch.cuda.device(device):
        out_s = torch.empty(num_queries, k, dtype=torch.float32, device=device)
        out_i = torch.empty(num_queries, k, dtype=torch.int64, device=device)
        status = _lib.lib.xb_topk_merge(
            num_queries, 1, length, k, scores.data_ptr(), ids.data_ptr(), out_s.data_ptr(), out_i.data_ptr(),
            _lib.stream_ptr(device),
        )
    _lib.check(status, "xb_topk_merge")
    return out_s, out_i


def topk_filter(
    scores: torch.Tensor, ids: torch.Tensor, exclude: torch.Tensor, k: int
) -> tuple[torch.Tensor, torch.Tensor]:
    """Drop the ids listed in ``exclude [Q, E]`` (padded with ``PAD_ID``) from ranked lists ``[Q, L]``; keep k.

    ``xb_topk_filter``: with ``L >= k + E`` this equals removing the excluded items BEFORE ranking, the
    ``prefilter=True`` semantics of ``ItemProcessor.search`` (data/lightning.py:247-252), without a dense
    ``Q x N`` exclusion mask.
    """
    device = _lib.require_cuda(scores, ids, exclude)
    scores = scores.to(torch.float32).contiguous()
    ids = ids.to(torch.int64).contiguous()
    exclude = exclude.to(torch.int64).contiguous()
    num_queries, length = scores.shape
    if exclude.dim() != 2 or exclude.size(0) != num_queries:  # noqa: PLR2004
        msg = f"exclude must be [{num_queries}, E], got {tuple(exclude.shape)}"
        raise ValueError(msg)
    with torch.cuda.device(device):
        out_s = torch.empty(num_queries, k, dtype=torch.float32, device=device)
        out_i = torch.empty(num_queries, k, dtype=torch.int64, device=device)
        status = _lib.lib.xb_topk_filter(
            num_queries, length, k, exclude.size(1), scores.data_ptr(), ids.data_ptr(),
            exclude.data_ptr() if exclude.numel() else None, out_s.data_ptr(), out_i.data_ptr(),
            _lib.stream_ptr(device),
        )
    _lib.check(status, "xb_topk_filter")
    return out_s, out_i


METRIC_NAMES = (
    "RetrievalNormalizedDCG",
    "RetrievalRecall",
    "RetrievalPrecision",
    "RetrievalMAP",
    "RetrievalHitRate",
    "RetrievalMRR",
)  # xfmr_rec/lightning.py:296-301, in the column order of xb_retrieval_metrics


def retrieval_metrics(
    ids: torch.Tensor, target_ids: torch.Tensor, target_vals: torch.Tensor
) -> tuple[torch.Tensor, torch.Tensor]:
    """``(per_query [Q, 6], mean [6])`` ranking metrics at ``k = ids.size(1)`` (``xb_retrieval_metrics``).

    ``ids [Q, k]`` ranked result lists (-1 = empty), ``target_ids [Q, T]`` padded with ``PAD_ID``,
    ``target_vals [Q, T]`` graded relevance.  Columns as ``METRIC_NAMES`` — what ``update_metrics``
    (xfmr_rec/lightning.py:149-187) feeds torchmetrics one user at a time.
    """
    device = _lib.require_cuda(ids, target_ids, target_vals)
    ids = ids.to(torch.int64).contiguous()
    target_ids = target_ids.to(torch.int64).contiguous()
    target_vals = target_vals.to(torch.float32).contiguous()
    num_queries, k = ids.shape
    if target_ids.shape != target_vals.shape or target_ids.dim() != 2 or target_ids.size(0) != num_queries:  # noqa: PLR2004
        msg = f"targets must be [{num_queries}, T] ids and values, got {tuple(target_ids.shape)} / {tuple(target_vals.shape)}"
        raise ValueError(msg)
    with torch.cuda.device(device):
        per_query = torch.empty(num_queries, len(METRIC_NAMES), dtype=torch.float32, device=device)
        mean = torch.empty(len(METRIC_NAMES), dtype=torch.float32, device=device)
        status = _lib.lib.xb_retrieval_metrics(
            num_queries, k, target_ids.size(1), ids.data_ptr(), target_ids.data_ptr(), target_vals.data_ptr(),
            per_query.data_ptr(), mean.data_ptr(), _lib.stream_ptr(device),
        )
    _lib.check(status, "xb_retrieval_metrics")
    return per_query, mean


def pad_id_lists(lists: Sequence[Sequence[int]], *, pad: int = _PAD_ID, dtype: torch.dtype = torch.int64) -> torch.Tensor:
    """Ragged python lists -> ``[len(lists), max_len]`` tensor padded with ``pad`` (at least one column)."""
    width = max(max((len(e) for e in lists), default=0), 1)
    return torch.tensor([list(e) + [pad] * (width - len(e)) for e in lists], dtype=dtype)


ITEMS_PARQUET = "items.parquet"        # bundle files written by ItemProcessor.save
PROCESSORS_JSON = "processors.json"    # same file name as the reference's export bundle (xfmr_rec/params.py)


def arrow_to_catalog(table, *, id_col: str, text_col: str | None, embedding_col: str = "embedding"):  # noqa: ANN001, ANN201
    """``(embeddings float32 [N, d] numpy, ids int64 numpy, texts list | None)`` from a pyarrow table with the schema
    the reference writes into LanceDB: ``movie_rn, movie_id, movie_text, embedding: fixed_size_list<float32, d>``
    (``ItemProcessor.get_index``, xfmr_rec/data/lightning.py:189, :208-219) — e.g. ``lance_table.to_arrow()`` or a parquet
    copy of it.  Variable-length list columns are accepted when every row has the same length.
    """
    import numpy as np  # noqa: PLC0415
    import pyarrow as pa  # noqa: PLC0415

    col = table[embedding_col]
    col = col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col
    num_rows = len(col)
    if col.null_count:
        msg = f"{col.null_count} rows have no embedding"
        raise ValueError(msg)
    flat = col.flatten().to_numpy(zero_copy_only=False)
    if num_rows == 0 or flat.size % num_rows != 0:
        msg = f"embedding column is ragged or empty: {flat.size} values in {num_rows} rows"
        raise ValueError(msg)
    if pa.types.is_list(col.type) or pa.types.is_large_list(col.type):
        lengths = np.diff(col.offsets.to_numpy())
        if (lengths != lengths[0]).any():
            msg = "embedding column is ragged"
            raise ValueError(msg)
    emb = np.ascontiguousarray(flat.reshape(num_rows, -1), dtype=np.float32)
    ids = np.ascontiguousarray(table[id_col].to_numpy(), dtype=np.int64)
    texts = table[text_col].to_pylist() if text_col is not None and text_col in table.column_names else None
    return emb, ids, texts


# the fields the reference's ``ItemProcessor`` dumps into ``processors.json`` (pydantic ``model_dump()`` of
# xfmr_rec/data/lightning.py:79-81, :128-133, :154-165; written at xfmr_rec/lightning.py:318-322) with its defaults
REFERENCE_ITEM_ARGS = {
    "batch_size": 32,
    "data_dir": "data",
    "idx_col": "movie_rn",
    "id_col": "movie_id",
    "text_col": "movie_text",
    "lance_table_name": "movies",
    "lance_db_path": "lance_db",
    "num_partitions": None,
    "num_sub_vectors": None,
    "num_probes": 8,
    "refine_factor": 4,
}


class ItemProcessor:
    """Exact-search stand-in for ``xfmr_rec.data.lightning.ItemProcessor`` (search side only).

    ``get_index`` takes the item embeddings directly (the reference encodes them with the model,
    data/lightning.py:182-219, which is outside this path) and keeps them on the GPU; ``search`` has the reference's
    signature and returns the reference's columns: every column of the item table (``idx_col`` when known, ``id_col``,
    ``text_col``, ``embedding``) plus ``score`` - ``recommend`` drops ``embedding`` from that frame
    (xfmr_rec/lightning.py:93-95).

    ``metric="cosine"`` (default, what the reference indexes with, data/lightning.py:222-229: ``score = 1 - cosine
    distance``): catalog rows are L2-normalised when the index is built and queries when they are searched, so scores are
    cosine similarities whatever the norms of the inputs.  ``metric="dot"`` scores raw inner products (identical for the
    unit-norm embeddings the reference's model emits, xfmr_rec/models.py:59).  Item ids must be non-negative: -1 marks
    an empty result slot (fewer than ``top_k`` candidates left after the exclusions).

    The constructor also accepts (and ``save`` writes back) the other fields of the reference's ``processors.json``
    entry, so ``ItemProcessor(**json.load(...)["items"])`` works on a reference export.
    """

    def __init__(self, *, idx_col: str = "movie_rn", id_col: str = "movie_id", text_col: str = "movie_text",
                 compute: str | None = None, metric: str = "cosine", **reference_args: object) -> None:
        if metric not in ("cosine", "dot"):
            msg = f"metric must be 'cosine' or 'dot', got {metric!r}"
            raise ValueError(msg)
        unknown = set(reference_args) - set(REFERENCE_ITEM_ARGS) - {"items_parquet"}
        if unknown:
            msg = f"unknown ItemProcessor arguments: {sorted(unknown)}"
            raise TypeError(msg)
        self.idx_col = idx_col
        self.id_col = id_col
        self.text_col = text_col
        self.compute = compute
        self.metric = metric
        self.reference_args = {k: reference_args.get(k, v) for k, v in REFERENCE_ITEM_ARGS.items()
                               if k not in ("idx_col", "id_col", "text_col")}
        self.embeddings: torch.Tensor | None = None
        self.item_ids: torch.Tensor | None = None
        self.item_idx: torch.Tensor | None = None      # the reference's ``movie_rn`` column (host tensor), when known
        self.item_text: Sequence[str] | None = None
        self._row_of_id: dict[int, int] | None = None   # item id -> catalog row, built on the first lookup

    def get_index(
        self,
        item_embeddings: torch.Tensor,
        item_ids: torch.Tensor | Sequence[int] | None = None,
        item_text: Sequence[str] | None = None,
        *,
        item_idx: torch.Tensor | Sequence[int] | None = None,
        device: torch.device | str = "cuda",
    ) -> ItemProcessor:
        emb = torch.as_tensor(item_embeddings)
        dtype = emb.dtype if emb.dtype in (torch.float32, torch.bfloat16) else torch.float32
        emb = emb.to(device)
        if self.metric == "cosine":
            # (normalised in fp32; a bf16 index rounds afterwards: cosine of the bf16-rounded unit vectors)
            emb = torch.nn.functional.normalize(emb.float(), dim=-1)
        self.embeddings = emb.to(dtype).contiguous()
        if item_ids is None:
            self.item_ids = None
        else:
            ids = torch.as_tensor(item_ids, dtype=torch.int64)
            if ids.numel() and int(ids.min()) < 0:
                msg = "item ids must be non-negative (-1 marks an empty result slot)"
                raise ValueError(msg)
            self.item_ids = ids.to(device).contiguous()
        self.item_idx = None if item_idx is None else torch.as_tensor(item_idx, dtype=torch.int64).cpu()
        self.item_text = item_text
        self._row_of_id = None
        return self

    def get_index_from_arrow(
        self,
        table,  # noqa: ANN001  pyarrow.Table
        *,
        embedding_col: str = "embedding",
        device: torch.device | str = "cuda",
        dtype: torch.dtype | None = None,
        rank: int = 0,
        world_size: int = 1,
    ) -> ItemProcessor:
        """Index from the item table the reference keeps in LanceDB: ``id_col``, ``text_col`` and a
        ``fixed_size_list<float32, d>`` embedding column (data/lightning.py:189, :208-219).  ``rank`` / ``world_size``
        keep only this rank's contiguous row shard (catalog row-sharding of ``distributed.sharded_topk``)."""
        import numpy as np  # noqa: PLC0415

        emb, ids, texts = arrow_to_catalog(table, id_col=self.id_col, text_col=self.text_col, embedding_col=embedding_col)
        idx = None
        if self.idx_col in table.column_names:
            idx = np.ascontiguousarray(table[self.idx_col].to_numpy(), dtype=np.int64)
        if world_size > 1:
            bounds = [len(ids) * r // world_size for r in range(world_size + 1)]
            lo, hi = bounds[rank], bounds[rank + 1]
            emb, ids = emb[lo:hi], ids[lo:hi]
            texts = texts[lo:hi] if texts is not None else None
            idx = idx[lo:hi] if idx is not None else None
        emb_t = torch.from_numpy(np.array(emb, copy=True))
        if dtype is not None:
            emb_t = emb_t.to(dtype)
        return self.get_index(emb_t, torch.from_numpy(np.array(ids, copy=True)), texts,
                              item_idx=None if idx is None else torch.from_numpy(np.array(idx, copy=True)), device=device)

    def to_arrow(self):  # noqa: ANN201
        """The index as a pyarrow table with the reference's column layout (``idx_col`` when known, ``id_col``, ``text_col``,
        ``embedding: fixed_size_list<float32, d>``; data/lightning.py:189, :208-219)."""
        import pyarrow as pa  # noqa: PLC0415

        if self.embeddings is None:
            msg = "index is empty: call get_index first"
            raise RuntimeError(msg)
        emb = self.embeddings.float().cpu().numpy()
        num_items, dim = emb.shape
        ids = self.item_ids.cpu().numpy() if self.item_ids is not None else torch.arange(num_items).numpy()
        columns = {}
        if self.item_idx is not None:
            columns[self.idx_col] = pa.array(self.item_idx.numpy(), type=pa.int64())
        columns[self.id_col] = pa.array(ids, type=pa.int64())
        if self.item_text is not None:
            columns[self.text_col] = pa.array(list(self.item_text), type=pa.string())
        columns["embedding"] = pa.FixedSizeListArray.from_arrays(pa.array(emb.reshape(-1), type=pa.float32()), dim)
        return pa.table(columns)

    def save(self, path) -> None:  # noqa: ANN001
        """Write the index bundle: ``items.parquet`` (reference table layout) + ``processors.json`` with the processor
        arguments under ``"items"`` in the reference's schema (every field of its ``ItemProcessor.model_dump()``,
        xfmr_rec/lightning.py:318-322) plus this class's own ``compute`` / ``metric`` / ``items_parquet`` keys, which the
        reference's pydantic model ignores."""
        import json  # noqa: PLC0415
        import pathlib  # noqa: PLC0415

        import pyarrow.parquet as pq  # noqa: PLC0415

        path = pathlib.Path(path)
        path.mkdir(parents=True, exist_ok=True)
        pq.write_table(self.to_arrow(), path / ITEMS_PARQUET)
        items = {**self.reference_args, "idx_col": self.idx_col, "id_col": self.id_col, "text_col": self.text_col,
                 "compute": self.compute, "metric": self.metric, "items_parquet": ITEMS_PARQUET}
        args = {"items": {k: items[k] for k in [*REFERENCE_ITEM_ARGS, "compute", "metric", "items_parquet"]}}
        (path / PROCESSORS_JSON).write_text(json.dumps(args, indent=2))

    @classmethod
    def load(cls, path, *, device: torch.device | str = "cuda", dtype: torch.dtype | None = None,  # noqa: ANN001
             rank: int = 0, world_size: int = 1) -> ItemProcessor:
        """Read a bundle written by ``save`` (or any directory with an ``items.parquet`` in the reference's layout)."""
        import json  # noqa: PLC0415
        import pathlib  # noqa: PLC0415

        import pyarrow.parquet as pq  # noqa: PLC0415

        path = pathlib.Path(path)
        args = {}
        if (path / PROCESSORS_JSON).exists():
            args = json.loads((path / PROCESSORS_JSON).read_text()).get("items", {})
        known = {k: v for k, v in args.items() if k in REFERENCE_ITEM_ARGS or k in ("compute", "metric")}
        return cls(**known).get_index_from_arrow(
            pq.read_table(path / ITEMS_PARQUET), device=device, dtype=dtype, rank=rank, world_size=world_size
        )

    # exclusion lists go through a dense [Q, N] bit mask (exact for any list length) while that mask is small, and
    # through a post-filter of the k + E best otherwise (exact while k + E <= MAX_K) — see _exclusions
    DENSE_MASK_BYTES = 256 << 20
    MAX_K = 256

    def _exclusions(
        self, exclude: torch.Tensor | Sequence[Sequence[int]] | None, num_queries: int, top_k: int
    ) -> tuple[torch.Tensor | None, torch.Tensor | None]:
        """``(dense bit mask, sparse id lists)`` — at most one of them is not None."""
        if exclude is None:
            return None, None
        assert self.embeddings is not None
        device = self.embeddings.device
        if not isinstance(exclude, torch.Tensor):
            if max((len(e) for e in exclude), default=0) == 0:
                return None, None
            exclude = pad_id_lists(exclude)
        if exclude.numel() == 0:
            return None, None
        if exclude.size(0) != num_queries:
            msg = f"one exclusion list per query expected: {exclude.size(0)} lists for {num_queries} queries"
            raise ValueError(msg)
        exclude = exclude.to(device)
        num_items = self.embeddings.size(0)
        mask_bytes = (-(-num_queries // 128) * 128) * (-(-num_items // 128) * 16)
        sparse_ok = top_k + exclude.size(1) <= self.MAX_K
        if mask_bytes > self.DENSE_MASK_BYTES:
            if not sparse_ok:
                msg = (
                    f"exclusion lists of {exclude.size(1)} ids with top_k={top_k} need a dense {mask_bytes >> 20} MiB mask "
                    f"(limit {self.DENSE_MASK_BYTES >> 20} MiB) or top_k + list length <= {self.MAX_K}; search in smaller "
                    "query batches"
                )
                raise ValueError(msg)
            return None, exclude
        col_ids = self.item_ids
        if col_ids is None:
            col_ids = torch.arange(num_items, dtype=torch.int64, device=device)
        mask, _ = build_pair_mask(col_ids, exclude)
        return mask, None

    def search_batch(
        self,
        embedding: torch.Tensor,
        exclude_item_ids: torch.Tensor | Sequence[Sequence[int]] | None = None,
        top_k: int = TOP_K,
    ) -> tuple[torch.Tensor, torch.Tensor]:
        """Batched search: ``embedding [Q, d]`` -> ``(scores [Q, k], item ids [Q, k])`` on the GPU."""
        if self.embeddings is None:
            msg = "index is empty: call get_index first"
            raise RuntimeError(msg)
        queries = torch.as_tensor(embedding)
        if queries.dim() == 1:
            queries = queries[None, :]
        queries = queries.to(self.embeddings.device)
        if self.metric == "cosine":
            queries = torch.nn.functional.normalize(queries.float(), dim=-1)
        mask, sparse = self._exclusions(exclude_item_ids, queries.size(0), top_k)
        if sparse is not None:
            fetch = top_k + sparse.size(1)
            scores, ids = topk_search(queries, self.embeddings, fetch, item_ids=self.item_ids, compute=self.compute)
            return topk_filter(scores, ids, sparse, top_k)
        return topk_search(
            queries, self.embeddings, top_k, item_ids=self.item_ids, excl_mask=mask, compute=self.compute
        )

    def evaluate(
        self,
        embedding: torch.Tensor,
        target_ids: torch.Tensor | Sequence[Sequence[int]],
        target_vals: torch.Tensor | Sequence[Sequence[float]],
        exclude_item_ids: torch.Tensor | Sequence[Sequence[int]] | None = None,
        top_k: int = TOP_K,
    ) -> dict[str, torch.Tensor]:
        """Batched validation: search + the six ranking metrics of ``update_metrics`` for all users at once.

        The reference evaluates one user per step (``validation_step`` -> ``update_metrics`` -> ``recommend`` ->
        ``search``, xfmr_rec/lightning.py:149-206, history excluded at :88-89).  Returns ``{metric name: mean over
        users}`` (0-d tensors on the GPU, names as torchmetrics' classes, :296-301) plus ``"per_query"`` ``[Q, 6]``.
        """
        _scores, ids = self.search_batch(embedding, exclude_item_ids, top_k)
        device = ids.device
        if not isinstance(target_ids, torch.Tensor):
            if [len(e) for e in target_ids] != [len(e) for e in target_vals]:
                msg = "target_ids and target_vals must be lists of equal lengths, row by row"
                raise ValueError(msg)
            target_vals = pad_id_lists(target_vals, pad=0, dtype=torch.float32)
            target_ids = pad_id_lists(target_ids)
        per_query, mean = retrieval_metrics(ids, target_ids.to(device), torch.as_tensor(target_vals).to(device))
        out = {name: mean[i] for i, name in enumerate(METRIC_NAMES)}
        out["per_query"] = per_query
        return out

    def search(
        self,
        embedding,  # noqa: ANN001  numpy [d] / [1, d] or tensor, as in the reference
        exclude_item_ids: list[int] | None = None,
        top_k: int = TOP_K,
    ) -> pd.DataFrame:
        """One query -> DataFrame sorted by score (data/lightning.py:237-259)."""
        import pandas as pd  # noqa: PLC0415

        exclude = [list(exclude_item_ids)] if exclude_item_ids else None
        scores, ids = self.search_batch(torch.as_tensor(embedding).reshape(1, -1), exclude, top_k)
        scores = scores[0].cpu()
        ids = ids[0].cpu()
        keep = scores > float("-inf")               # empty slots carry score -inf (and id -1)
        ids, scores = ids[keep], scores[keep]
        if self.item_ids is None:
            rows = ids.tolist()
        else:
            if self._row_of_id is None:   # once per index, not once per query
                self._row_of_id = {int(v): r for r, v in enumerate(self.item_ids.cpu().tolist())}
            rows = [self._row_of_id[int(v)] for v in ids.tolist()]
        # the reference returns every column of the item table + score (lance ``to_pandas()``, data/lightning.py:250-259)
        frame: dict[str, object] = {}
        if self.item_idx is not None:
            frame[self.idx_col] = self.item_idx[rows].numpy() if rows else self.item_idx[:0].numpy()
        frame[self.id_col] = ids.numpy()
        if self.item_text is not None:
            frame[self.text_col] = [self.item_text[r] for r in rows]
        emb_rows = self.embeddings[torch.as_tensor(rows, dtype=torch.int64, device=self.embeddings.device)].float().cpu().numpy()
        frame["embedding"] = list(emb_rows)
        frame["score"] = scores.numpy()
        return pd.DataFrame(frame)
