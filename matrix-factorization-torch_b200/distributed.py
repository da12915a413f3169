"""Sharding of the score path across the GPUs of one box (``torch.distributed``, NCCL over NVLink).

The reference has no collective of its own (Lightning DDP with one worker, ``xfmr_rec/ray.py:40, 106``);
both functions here are the north-star extensions specified in SURVEY.md 8(e):

* training — ``global_negatives_losses``: every rank all-gathers the other ranks' in-batch items and
  uniform negatives and scores its own users against the union.  The gathered rows are re-ordered so
  that the rank's own positives stay in rows ``0..B-1``; that keeps every convention of
  ``xfmr_rec/losses.py`` (diagonal = positive, ``item_idx[:B]`` = in-batch ids) and makes the oracle
  literally "the reference loss called once per rank on the re-ordered concatenation".  Gradients of
  the gathered rows flow back to their owners through a sum-reduction (reduce-scatter).
* retrieval — ``sharded_topk``: the catalog is row-sharded, queries are replicated, each rank returns
  its local top-k and the ``[G, Q, k]`` candidates are merged with the (score desc, id asc)
  comparator, which is deterministic for any G.

Host logic only; the loss / search callables are injected so the CPU (gloo) tests can drive the
same code with the oracle.
"""

from __future__ import annotations

from typing import TYPE_CHECKING

import torch
import torch.distributed as dist

if TYPE_CHECKING:
    from collections.abc import Callable


class _AllGatherRows(torch.autograd.Function):
    """All-gather of equally-shaped row blocks; backward sum-reduces each block to its owner."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, group) -> torch.Tensor:  # noqa: ANN001
        ctx.group = group
        ctx.rows = x.size(0)
        return _gather_plain(x, group)  # [G, rows, d]

    @staticmethod
    def backward(ctx, grad: torch.Tensor):  # noqa: ANN001, ANN205
        group = ctx.group
        rank = dist.get_rank(group)
        grad = grad.contiguous()
        if dist.get_backend(group) == "nccl":
            out = torch.empty_like(grad[0])
            dist.reduce_scatter_tensor(out, grad.reshape(-1, *grad.shape[2:]), group=group)
            return out, None
        # gloo has no reduce-scatter: all-reduce and keep the own slice
        dist.all_reduce(grad, group=group)
        return grad[rank].clone(), None


def _gather_plain(x: torch.Tensor, group) -> torch.Tensor:  # noqa: ANN001
    """[rows, ...] per rank -> [G, rows, ...]: ONE collective straight into the stacked buffer (no per-rank list + stack)."""
    world = dist.get_world_size(group)
    x = x.contiguous()
    out = x.new_empty((world, *x.shape))
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, x, group=group)
    else:  # gloo: all_gather_into_tensor is not available for every dtype / version
        dist.all_gather(list(out.unbind(0)), x, group=group)
    return out


def _own_first(stacked: torch.Tensor, rank: int) -> torch.Tensor:
    """[G, rows, ...] -> [G*rows, ...] with block `rank` first, the others in rank order."""
    world = stacked.size(0)
    order = [rank] + [r for r in range(world) if r != rank]
    return torch.cat([stacked[r] for r in order], dim=0)


def global_negatives_inputs(
    item_embed: torch.Tensor,
    neg_embed: torch.Tensor,
    item_idx: torch.Tensor,
    neg_idx: torch.Tensor,
    *,
    log_q_item: torch.Tensor | None = None,
    log_q_neg: torch.Tensor | None = None,
    group=None,  # noqa: ANN001
) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor | None]:
    """Candidate set of this rank: ``[own items | other ranks' items | all ranks' negatives]``.

    Returns ``(item_embed_all [G*(B+U), d], item_idx_all, log_q_all)``; differentiable in
    ``item_embed`` and ``neg_embed`` (gradient contributions from every rank are summed at the owner).
    Mirrors the single-rank concatenation of ``xfmr_rec/lightning.py:125-135``.
    """
    rank = dist.get_rank(group)
    items = _own_first(_AllGatherRows.apply(item_embed, group), rank)
    negs = _own_first(_AllGatherRows.apply(neg_embed, group), rank)
    idx_items = _own_first(_gather_plain(item_idx, group), rank)
    idx_negs = _own_first(_gather_plain(neg_idx, group), rank)
    log_q = None
    if log_q_item is not None and log_q_neg is not None:
        log_q = torch.cat(
            [_own_first(_gather_plain(log_q_item, group), rank), _own_first(_gather_plain(log_q_neg, group), rank)]
        )
    return torch.cat([items, negs], dim=0), torch.cat([idx_items, idx_negs], dim=0), log_q


def global_negatives_losses(
    loss_fn: Callable[..., torch.Tensor],
    user_embed: torch.Tensor,
    item_embed: torch.Tensor,
    neg_embed: torch.Tensor,
    target: torch.Tensor,
    *,
    item_idx: torch.Tensor,
    neg_idx: torch.Tensor,
    pos_idx: torch.Tensor,
    group=None,  # noqa: ANN001
    **loss_kwargs,  # noqa: ANN003
) -> torch.Tensor:
    """Per-rank loss against the global candidate set; the job's loss is the sum over ranks.

    ``loss_fn(user_embed, item_embed_all, target, item_idx=, pos_idx=, **loss_kwargs)`` is any of the
    drop-in modules / ``fused_losses`` (or the oracle in CPU tests).
    """
    items_all, idx_all, _ = global_negatives_inputs(item_embed, neg_embed, item_idx, neg_idx, group=group)
    return loss_fn(user_embed, items_all, target, item_idx=idx_all, pos_idx=pos_idx, **loss_kwargs)


def sharded_topk(
    search_fn: Callable[[torch.Tensor, int], tuple[torch.Tensor, torch.Tensor]],
    merge_fn: Callable[[torch.Tensor, torch.Tensor, int], tuple[torch.Tensor, torch.Tensor]],
    queries: torch.Tensor,
    k: int,
    *,
    group=None,  # noqa: ANN001
    gather: bool = True,
) -> tuple[torch.Tensor, torch.Tensor]:
    """Row-sharded catalog search: local top-k on every rank, exchange, merge.

    ``search_fn(queries, k) -> (scores [Q, k], global ids [Q, k])`` searches the local shard (e.g.
    ``ItemProcessor.search_batch`` with ``item_ids`` / ``id_base`` carrying global ids);
    ``merge_fn(scores [Q', G*k], ids [Q', G*k], k)`` is ``retrieval.topk_merge`` on GPU.

    NCCL: the merge is sharded by query.  One ``all_to_all_single`` per tensor hands every rank the ``G`` per-shard lists
    of ITS ``Q / G`` queries (each rank sends ``(G - 1) / G`` of its ``[Q, k]`` lists: 69 MB at config 5 on 8 GPUs
    instead of the 629 MB an all-gather of everything moves), the rank merges those rows, and - with ``gather=True`` -
    one ``all_gather_into_tensor`` per tensor gives every rank the full ``[Q, k]`` result.  ``gather=False`` returns the
    rank's own query slice ``[ceil(Q / G), k]`` (rows ``rank * ceil(Q / G) ...``), which is all a serving front-end that
    owns those queries needs.  Other backends (gloo in the CPU tests has no all-to-all): all-gather + merge of every row.
    """
    scores, ids = search_fn(queries, k)
    world = dist.get_world_size(group)
    if world == 1:
        return merge_fn(scores, ids, k)
    num_q = scores.size(0)
    if dist.get_backend(group) != "nccl":
        all_scores = _gather_plain(scores, group)  # [G, Q, k]
        all_ids = _gather_plain(ids, group)
        cat_scores = all_scores.permute(1, 0, 2).reshape(num_q, world * k)
        cat_ids = all_ids.permute(1, 0, 2).reshape(num_q, world * k)
        out_scores, out_ids = merge_fn(cat_scores, cat_ids, k)
        if gather:
            return out_scores, out_ids
        per = -(-num_q // world)
        rank = dist.get_rank(group)
        return out_scores[rank * per : (rank + 1) * per], out_ids[rank * per : (rank + 1) * per]
    per = -(-num_q // world)                      # queries merged by one rank
    pad = per * world - num_q
    if pad:
        scores = torch.cat([scores, scores.new_full((pad, k), float("-inf"))])
        ids = torch.cat([ids, ids.new_full((pad, k), -1)])
    recv_scores = torch.empty_like(scores)        # [G (source shard), per, k]
    recv_ids = torch.empty_like(ids)
    dist.all_to_all_single(recv_scores, scores.contiguous(), group=group)
    dist.all_to_all_single(recv_ids, ids.contiguous(), group=group)
    cat_scores = recv_scores.view(world, per, k).transpose(0, 1).reshape(per, world * k)
    cat_ids = recv_ids.view(world, per, k).transpose(0, 1).reshape(per, world * k)
    my_scores, my_ids = merge_fn(cat_scores, cat_ids, k)
    if not gather:
        return my_scores, my_ids
    out_scores = torch.empty(world * per, k, dtype=my_scores.dtype, device=my_scores.device)
    out_ids = torch.empty(world * per, k, dtype=my_ids.dtype, device=my_ids.device)
    dist.all_gather_into_tensor(out_scores, my_scores.contiguous(), group=group)
    dist.all_gather_into_tensor(out_ids, my_ids.contiguous(), group=group)
    return out_scores[:num_q], out_ids[:num_q]


class RetrievalGrid:
    """``world = R x (world / R)``: rank ``r`` owns catalog shard ``r % R`` and query group ``r // R``.

    ``catalog_group`` holds the ``R`` ranks that share a query group (they exchange and merge their per-shard lists);
    ``search`` returns the full ``[Q, k]`` result on every rank (rank-ordered all-gather of the merged query slices).
    ``R = world`` is plain row sharding (``sharded_topk``).  ``R < world`` is for catalogs that need fewer shards than
    there are GPUs; at config 5 it is NOT faster (8 GPUs: 249 k queries/s with 4 shards x 2 query groups against 270 k
    with 8 shards, profiles/r02_layouts_8gpu.txt - under the power cap the longer streams run at lower clocks)."""

    def __init__(self, shards: int, *, group=None) -> None:  # noqa: ANN001
        self.world_group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world % shards:
            msg = f"{shards} catalog shards do not divide {self.world} ranks"
            raise ValueError(msg)
        self.shards = shards
        self.query_groups = self.world // shards
        self.shard = self.rank % shards
        self.query_group = self.rank // shards
        self.catalog_group = group
        if self.query_groups > 1:  # every rank creates every subgroup (torch.distributed requirement), keeps its own
            for g in range(self.query_groups):
                sub = dist.new_group(list(range(g * shards, (g + 1) * shards)))
                if g == self.query_group:
                    self.catalog_group = sub

    def rows_per_group(self, num_queries: int) -> int:
        """Query rows per group, padded so that the ``R`` merging ranks of a group take equal slices."""
        per = -(-num_queries // self.query_groups)
        return -(-per // self.shards) * self.shards

    def query_slice(self, num_queries: int) -> slice:
        """Rows of the query matrix this rank's query group searches (the last groups may hold fewer, or none)."""
        per = self.rows_per_group(num_queries)
        return slice(min(self.query_group * per, num_queries), min((self.query_group + 1) * per, num_queries))

    def search(
        self,
        search_fn: Callable[[torch.Tensor, int], tuple[torch.Tensor, torch.Tensor]],
        merge_fn: Callable[[torch.Tensor, torch.Tensor, int], tuple[torch.Tensor, torch.Tensor]],
        queries: torch.Tensor,
        k: int,
    ) -> tuple[torch.Tensor, torch.Tensor]:
        """``queries`` is the full ``[Q, d]`` matrix (replicated); ``search_fn`` searches this rank's catalog shard."""
        num_q = queries.size(0)
        if self.query_groups == 1:
            return sharded_topk(search_fn, merge_fn, queries, k, group=self.world_group)
        per_group = self.rows_per_group(num_q)
        mine = queries[self.query_slice(num_q)]
        if mine.size(0) < per_group:  # pad so that every rank contributes the same number of rows
            mine = torch.cat([mine, mine.new_zeros(per_group - mine.size(0), mine.size(1))])
        my_scores, my_ids = sharded_topk(search_fn, merge_fn, mine, k, group=self.catalog_group, gather=False)
        out_scores = my_scores.new_empty(self.world * my_scores.size(0), k)
        out_ids = my_ids.new_empty(self.world * my_ids.size(0), k)
        if dist.get_backend(self.world_group) == "nccl":
            dist.all_gather_into_tensor(out_scores, my_scores.contiguous(), group=self.world_group)
            dist.all_gather_into_tensor(out_ids, my_ids.contiguous(), group=self.world_group)
        else:
            dist.all_gather(list(out_scores.view(self.world, -1, k).unbind(0)), my_scores.contiguous(), group=self.world_group)
            dist.all_gather(list(out_ids.view(self.world, -1, k).unbind(0)), my_ids.contiguous(), group=self.world_group)
        return out_scores[:num_q], out_ids[:num_q]
