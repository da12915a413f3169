"""Hashed-embedding gather: indices bit-exact against XXH32, rows bit-exact against an fp32 embedding-bag sum."""

from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_indices_bit_exact() -> None:
    import xfmr_b200  # noqa: PLC0415
    from oracle import native  # noqa: PLC0415

    rng = np.random.default_rng(0)
    ids = np.concatenate([rng.integers(0, 2**63 - 1, 100_000, dtype=np.int64),
                          np.array([0, 1, 2, 3706, 87585, 2**31, 2**40 + 7, 99999999, -1, -(2**63)], dtype=np.int64)])
    for nh, log2 in ((2, 22), (1, 10), (4, 31)):
        got = xfmr_b200.hash_indices(torch.from_numpy(ids).cuda(), nh, log2).cpu().numpy()
        assert np.array_equal(got, native.hash_indices(ids, nh, log2))
    assert xfmr_b200.hash_indices(torch.tensor([3706]).cuda(), 2, 22).tolist() == [[3426034, 1982363]]
    assert xfmr_b200.hash_indices(torch.empty(0, dtype=torch.int64).cuda(), 2, 22).shape == (0, 2)


@pytest.mark.parametrize(("n", "nh", "log2", "d"), [(1000, 2, 12, 128), (4097, 3, 10, 64), (333, 1, 8, 8), (50_000, 2, 16, 256)])
def test_gather_bit_exact(n: int, nh: int, log2: int, d: int) -> None:
    import xfmr_b200  # noqa: PLC0415
    from oracle import native  # noqa: PLC0415

    gen = torch.Generator().manual_seed(n)
    table = (torch.randn(1 << log2, d, generator=gen) * 0.02).to(torch.bfloat16)
    ids = torch.randint(0, 2**62, (n,), generator=gen)
    out = xfmr_b200.hash_embedding_gather(table.cuda(), ids.cuda(), nh)
    acc, _ = native.hash_gather(table.float().numpy(), ids.numpy(), nh, log2)
    ref = torch.from_numpy(acc).to(torch.bfloat16)
    assert out.dtype == torch.bfloat16
    assert torch.equal(out.cpu().view(torch.int16), ref.view(torch.int16))
    # same answer as torch's embedding_bag(sum) on the oracle's indices
    idx = torch.from_numpy(native.hash_indices(ids.numpy(), nh, log2)).long()
    bag = torch.nn.functional.embedding_bag(idx, table.float(), mode="sum").to(torch.bfloat16)
    assert torch.equal(out.cpu().view(torch.int16), bag.view(torch.int16))


def test_gather_feeds_the_loss_and_backpropagates() -> None:
    """End of the feeder: ids -> hashed rows -> fused loss; gradient reaches the table (scatter-add)."""
    import xfmr_b200  # noqa: PLC0415

    dev = torch.device("cuda:0")
    users = xfmr_b200.HashEmbeddingBag(12, 64, num_hashes=2).to(dev)
    items = xfmr_b200.HashEmbeddingBag(12, 64, num_hashes=2, seed0=7).to(dev)
    uid = torch.arange(1, 65, device=dev)
    iid = torch.arange(1, 161, device=dev)
    q = users(uid)
    v = items(iid)
    loss = xfmr_b200.PairwiseLogisticLoss()(q, v, torch.ones(64, device=dev), item_idx=iid,
                                            pos_idx=torch.zeros(64, 1, dtype=torch.int64, device=dev))
    loss.backward()
    g = users.weight.grad
    assert g is not None and g.shape == users.weight.shape and torch.isfinite(g.float()).all()
    touched = torch.unique(xfmr_b200.hash_indices(uid, 2, 12).flatten().long())
    untouched = torch.ones(1 << 12, dtype=torch.bool, device=dev)
    untouched[touched] = False
    assert (g[untouched] == 0).all()
    assert g[touched].float().abs().sum() > 0
