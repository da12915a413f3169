"""World-size-2 gloo tests of the multi-GPU host logic, with the oracle standing in for the
CUDA kernels: global-negative all-gather + gradient reduction, and catalog-sharded top-k with merge."""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_loss(user, items, target, *, item_idx, pos_idx, **kw):  # noqa: ANN001, ANN003, ANN202
    from oracle import losses_oracle  # noqa: PLC0415

    name = "InfomationNoiseContrastiveEstimationLoss"
    return losses_oracle.all_losses(user, items, target, item_idx=item_idx, pos_idx=pos_idx, names=(name,), **kw)[name]


def _global_negatives_worker(rank: int, world: int, port: int, out: dict) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import xfmr_b200  # noqa: PLC0415
    from xfmr_b200 import synthetic  # noqa: PLC0415

    b, u, d = 12, 20, 16
    data = [synthetic.make_loss_inputs(b, b + u, d, 3, n_catalog=40, seed=50 + r, mean_extra_pos=1.0) for r in range(world)]
    mine = data[rank]
    q = mine["user_embed"].double().requires_grad_(True)
    items = mine["item_embed"][:b].double().requires_grad_(True)
    negs = mine["item_embed"][b:].double().requires_grad_(True)
    loss = xfmr_b200.distributed.global_negatives_losses(
        _oracle_loss, q, items, negs, mine["target"].double(), item_idx=mine["item_idx"][:b], neg_idx=mine["item_idx"][b:],
        pos_idx=mine["pos_idx"], sigma=2.0,
    )
    loss.backward()
    out[rank] = (loss.item(), q.grad.clone(), items.grad.clone(), negs.grad.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_global_negatives_match_single_process_reference() -> None:
    world = 2
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_global_negatives_worker, args=(world, _free_port(), out), nprocs=world, join=True)

    from xfmr_b200 import synthetic  # noqa: PLC0415

    b, u, d = 12, 20, 16
    data = [synthetic.make_loss_inputs(b, b + u, d, 3, n_catalog=40, seed=50 + r, mean_extra_pos=1.0) for r in range(world)]
    leaves_q = [x["user_embed"].double().requires_grad_(True) for x in data]
    leaves_i = [x["item_embed"][:b].double().requires_grad_(True) for x in data]
    leaves_n = [x["item_embed"][b:].double().requires_grad_(True) for x in data]
    total = 0.0
    for r in range(world):
        order = [r] + [o for o in range(world) if o != r]
        items_all = torch.cat([leaves_i[o] for o in order] + [leaves_n[o] for o in order])
        idx_all = torch.cat([data[o]["item_idx"][:b] for o in order] + [data[o]["item_idx"][b:] for o in order])
        loss_r = _oracle_loss(leaves_q[r], items_all, data[r]["target"].double(), item_idx=idx_all, pos_idx=data[r]["pos_idx"], sigma=2.0)
        assert out[r][0] == pytest.approx(loss_r.item(), rel=1e-12)
        total = total + loss_r
    total.backward()
    for r in range(world):
        assert torch.allclose(out[r][1], leaves_q[r].grad, rtol=1e-10, atol=1e-12)
        assert torch.allclose(out[r][2], leaves_i[r].grad, rtol=1e-10, atol=1e-12)   # summed over every rank's loss
        assert torch.allclose(out[r][3], leaves_n[r].grad, rtol=1e-10, atol=1e-12)


def _sharded_topk_worker(rank: int, world: int, port: int, out: dict) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import xfmr_b200  # noqa: PLC0415
    from oracle import native  # noqa: PLC0415

    rng = np.random.default_rng(3)
    queries = rng.standard_normal((9, 16)).astype(np.float32)
    catalog = rng.standard_normal((101, 16)).astype(np.float32)
    catalog[70] = catalog[4]                        # a tie across shards: the lower id must win
    bounds = np.linspace(0, 101, world + 1).astype(int)
    lo, hi = bounds[rank], bounds[rank + 1]

    def search(q: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
        s, i = native.topk(q.numpy(), catalog[lo:hi], k, id_base=int(lo))
        return torch.from_numpy(s), torch.from_numpy(i)

    def merge(scores: torch.Tensor, ids: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
        order = sorted(range(scores.size(1)), key=lambda j: 0)  # placeholder to keep the signature obvious
        del order
        rows_s, rows_i = [], []
        for r in range(scores.size(0)):
            pairs = sorted(((-float(s), int(i)) for s, i in zip(scores[r], ids[r]) if int(i) >= 0))[:k]
            rows_s.append([-p[0] for p in pairs])
            rows_i.append([p[1] for p in pairs])
        return torch.tensor(rows_s), torch.tensor(rows_i)

    s, i = xfmr_b200.distributed.sharded_topk(search, merge, torch.from_numpy(queries), 7)
    ref_s, ref_i = native.topk(queries, catalog, 7)
    out[rank] = bool(np.array_equal(i.numpy(), ref_i) and np.allclose(s.numpy(), ref_s, rtol=0, atol=0))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_topk_equals_unsharded() -> None:
    world = 2
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_sharded_topk_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def _sharded_exclusions_worker(rank: int, world: int, port: int, out: dict) -> None:
    """Exclusion lists are applied shard-locally (rank the k + E best of the shard, drop the listed ids, keep k) and the
    filtered lists are merged: must equal the global pre-filtered top-k (data/lightning.py:247-252)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import xfmr_b200  # noqa: PLC0415
    from oracle import native  # noqa: PLC0415

    rng = np.random.default_rng(11)
    queries = rng.standard_normal((13, 8)).astype(np.float32)
    catalog = rng.standard_normal((90, 8)).astype(np.float32)
    full = queries @ catalog.T
    excl = np.full((13, 6), native.PAD_ID, dtype=np.int64)
    for r in range(13):
        excl[r, : r % 6] = np.argsort(-full[r])[: r % 6]          # 0..5 of the best items, ragged
    bounds = np.linspace(0, 90, world + 1).astype(int)
    lo, hi = bounds[rank], bounds[rank + 1]
    n_excl = excl.shape[1]

    def search(q: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
        # the shard-local sparse path of ItemProcessor.search_batch: rank k + E, then filter (python stand-in for xb_topk_filter)
        s, i = native.topk(q.numpy(), catalog[lo:hi], k + n_excl, id_base=int(lo))
        rows_s, rows_i = [], []
        for r in range(s.shape[0]):
            banned = set(excl[r].tolist())
            kept = [(float(a), int(b)) for a, b in zip(s[r], i[r]) if int(b) >= 0 and int(b) not in banned][:k]
            kept += [(float("-inf"), -1)] * (k - len(kept))
            rows_s.append([p[0] for p in kept])
            rows_i.append([p[1] for p in kept])
        return torch.tensor(rows_s), torch.tensor(rows_i)

    def merge(scores: torch.Tensor, ids: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
        rows_s, rows_i = [], []
        for r in range(scores.size(0)):
            pairs = sorted(((-float(s), int(i)) for s, i in zip(scores[r], ids[r]) if int(i) >= 0))[:k]
            rows_s.append([-p[0] for p in pairs])
            rows_i.append([p[1] for p in pairs])
        return torch.tensor(rows_s), torch.tensor(rows_i)

    s, i = xfmr_b200.distributed.sharded_topk(search, merge, torch.from_numpy(queries), 5)
    ref_s, ref_i = native.topk(queries, catalog, 5, exclude=excl)
    out[rank] = bool(np.array_equal(i.numpy(), ref_i) and np.array_equal(s.numpy().astype(np.float32), ref_s))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_local_exclusions_equal_the_global_prefilter() -> None:
    world = 2
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_sharded_exclusions_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def _retrieval_grid_worker(rank: int, world: int, port: int, out: dict, nq: int = 10) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import xfmr_b200  # noqa: PLC0415
    from oracle import native  # noqa: PLC0415

    rng = np.random.default_rng(5)
    queries = rng.standard_normal((nq, 16)).astype(np.float32)   # 10: 2 query groups x 5 rows -> padded to 6 per group
    catalog = rng.standard_normal((97, 16)).astype(np.float32)
    catalog[60] = catalog[3]                        # a tie across shards: the lower id must win
    grid = xfmr_b200.distributed.RetrievalGrid(2)   # 4 ranks = 2 catalog shards x 2 query groups
    bounds = np.linspace(0, 97, grid.shards + 1).astype(int)
    lo, hi = bounds[grid.shard], bounds[grid.shard + 1]

    def search(q: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
        s, i = native.topk(q.numpy(), catalog[lo:hi], k, id_base=int(lo))
        return torch.from_numpy(s), torch.from_numpy(i)

    def merge(scores: torch.Tensor, ids: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
        rows_s, rows_i = [], []
        for r in range(scores.size(0)):
            pairs = sorted(((-float(s), int(i)) for s, i in zip(scores[r], ids[r]) if int(i) >= 0))[:k]
            rows_s.append([-p[0] for p in pairs])
            rows_i.append([p[1] for p in pairs])
        return torch.tensor(rows_s, dtype=torch.float32), torch.tensor(rows_i, dtype=torch.int64)

    s, i = grid.search(search, merge, torch.from_numpy(queries), 7)
    ref_s, ref_i = native.topk(queries, catalog, 7)
    ok = (grid.shards, grid.query_groups, grid.shard, grid.query_group) == (2, 2, rank % 2, rank // 2)
    out[rank] = bool(ok and np.array_equal(i.numpy(), ref_i) and np.array_equal(s.numpy(), ref_s))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nq", [10, 3, 1])   # 3 and 1: the second query group is short / empty (all padding)
def test_retrieval_grid_two_shards_by_two_query_groups(nq: int) -> None:
    """2-D layout (catalog shards x query groups) on 4 gloo ranks equals the unsharded search, ties and padding included."""
    world = 4
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_retrieval_grid_worker, args=(world, _free_port(), out, nq), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
