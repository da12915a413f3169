"""Multi-GPU (NCCL) tests of SURVEY.md 8(e): global negatives for training and catalog sharding for retrieval,
with the CUDA kernels on every rank and the CPU oracle as the checker.  Needs >= 2 GPUs (skipped otherwise)."""

from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

WORLD = 2


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _need_gpus() -> None:
    if torch.cuda.device_count() < WORLD:
        pytest.skip(f"needs {WORLD} GPUs, found {torch.cuda.device_count()}")


def _inputs(rank: int, b: int, u: int, d: int):  # noqa: ANN202
    from xfmr_b200 import synthetic  # noqa: PLC0415

    return synthetic.make_loss_inputs(b, b + u, d, 4, n_catalog=300, seed=900 + rank, mean_extra_pos=1.5)


def _global_negatives_worker(rank: int, port: int, out: dict) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    import xfmr_b200  # noqa: PLC0415

    b, u, d = 96, 160, 64
    mine = _inputs(rank, b, u, d)
    q = mine["user_embed"].to(dev).requires_grad_(True)
    items = mine["item_embed"][:b].to(dev).requires_grad_(True)
    negs = mine["item_embed"][b:].to(dev).requires_grad_(True)
    loss_fn = xfmr_b200.InfomationNoiseContrastiveEstimationLoss(sigma=3.0)
    loss = xfmr_b200.distributed.global_negatives_losses(
        loss_fn, q, items, negs, mine["target"].to(dev), item_idx=mine["item_idx"][:b].to(dev),
        neg_idx=mine["item_idx"][b:].to(dev), pos_idx=mine["pos_idx"].to(dev),
    )
    loss.backward()
    out[rank] = (loss.item(), q.grad.cpu(), items.grad.cpu(), negs.grad.cpu())
    dist.barrier()
    dist.destroy_process_group()


def test_global_negatives_over_nccl_match_oracle() -> None:
    _need_gpus()
    from oracle import losses_oracle  # noqa: PLC0415

    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_global_negatives_worker, args=(_free_port(), out), nprocs=WORLD, join=True)
    b, u, d = 96, 160, 64
    data = [_inputs(r, b, u, d) for r in range(WORLD)]
    lq = [x["user_embed"].double().requires_grad_(True) for x in data]
    li = [x["item_embed"][:b].double().requires_grad_(True) for x in data]
    ln = [x["item_embed"][b:].double().requires_grad_(True) for x in data]
    name = "InfomationNoiseContrastiveEstimationLoss"
    total = 0.0
    for r in range(WORLD):
        order = [r] + [o for o in range(WORLD) if o != r]
        items_all = torch.cat([li[o] for o in order] + [ln[o] for o in order])
        idx_all = torch.cat([data[o]["item_idx"][:b] for o in order] + [data[o]["item_idx"][b:] for o in order])
        loss_r = losses_oracle.all_losses(lq[r], items_all, data[r]["target"].double(), item_idx=idx_all,
                                          pos_idx=data[r]["pos_idx"], sigma=3.0, names=(name,))[name]
        assert out[r][0] == pytest.approx(loss_r.item(), rel=1e-3)
        total = total + loss_r
    total.backward()
    for r in range(WORLD):
        for got, ref in ((out[r][1], lq[r].grad), (out[r][2], li[r].grad), (out[r][3], ln[r].grad)):
            err = (got.double() - ref).norm() / ref.norm()
            assert err < 1e-3, float(err)


def _sharded_topk_worker(rank: int, port: int, out: dict) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=dev)
    import xfmr_b200  # noqa: PLC0415

    gen = torch.Generator().manual_seed(5)
    queries = torch.nn.functional.normalize(torch.randn(200, 64, generator=gen), dim=-1)
    catalog = torch.nn.functional.normalize(torch.randn(30_001, 64, generator=gen), dim=-1)
    catalog[20_000] = catalog[17]     # a tie across shards: the lower id must win
    bounds = np.linspace(0, catalog.size(0), WORLD + 1).astype(int)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    shard = catalog[lo:hi].to(dev)

    def search(q: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
        return xfmr_b200.topk_search(q, shard, k, id_base=lo)

    scores, ids = xfmr_b200.distributed.sharded_topk(search, xfmr_b200.topk_merge, queries.to(dev), 50)
    out[rank] = (scores.cpu().numpy(), ids.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_topk_over_nccl_is_bit_exact() -> None:
    _need_gpus()
    from oracle import native  # noqa: PLC0415

    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_sharded_topk_worker, args=(_free_port(), out), nprocs=WORLD, join=True)
    gen = torch.Generator().manual_seed(5)
    queries = torch.nn.functional.normalize(torch.randn(200, 64, generator=gen), dim=-1)
    catalog = torch.nn.functional.normalize(torch.randn(30_001, 64, generator=gen), dim=-1)
    catalog[20_000] = catalog[17]
    ref_s, ref_i = native.topk(queries.numpy(), catalog.numpy(), 50)
    for r in range(WORLD):
        assert np.array_equal(out[r][1], ref_i)
        assert np.array_equal(out[r][0], ref_s)
