"""Pins the CPU oracle: against the committed outputs of the unmodified reference, against the reference
itself when it is present (build container only), and against external known answers (XXH32)."""

from __future__ import annotations

import pathlib
import struct
import sys

import numpy as np
import pytest
import torch

from helpers import LOSS_NAMES, golden_cases, load_golden, rel_err
from oracle import losses_oracle, native

REFERENCE = pathlib.Path("/root/reference/xfmr_rec/losses.py")


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference_golden(name: str) -> None:
    case = load_golden(name)
    got = losses_oracle.losses_and_grads(
        case["user_embed"].double(), case["item_embed"].double(), case["target"].double(),
        item_idx=case["item_idx"], pos_idx=case["pos_idx"], num_negatives=case["num_negatives"],
        sigma=case["sigma"], margin=case["margin"], mining=case["mining"],
    )
    for n in LOSS_NAMES:
        loss, dq, dv = got[n]
        exp_loss, exp_dq, exp_dv = case["expected"][n]
        assert float(loss) == pytest.approx(exp_loss, rel=1e-9, abs=1e-9), n
        assert rel_err(dq, exp_dq) < 1e-6, n   # golden gradients are stored as float32
        assert rel_err(dv, exp_dv) < 1e-6, n


@pytest.mark.skipif(not REFERENCE.exists(), reason="reference sources only exist in the build container")
@pytest.mark.parametrize("k", [0, 3])
def test_oracle_matches_live_reference(k: int) -> None:
    sys.path.insert(0, "/root/reference")
    import xfmr_rec.losses as ref  # noqa: PLC0415

    from xfmr_b200 import synthetic  # noqa: PLC0415

    inp = synthetic.make_loss_inputs(37, 90, 24, 4, n_catalog=50, seed=77 + k, signed_targets=True, mean_extra_pos=2.0)
    for n in LOSS_NAMES:
        module = getattr(ref, n)(num_negatives=k, sigma=1.7, margin=0.4)
        q = inp["user_embed"].double().requires_grad_(True)
        v = inp["item_embed"].double().requires_grad_(True)
        loss = module(q, v, inp["target"].double(), item_idx=inp["item_idx"], pos_idx=inp["pos_idx"])
        dq, dv = torch.autograd.grad(loss, (q, v), allow_unused=True)
        got = losses_oracle.losses_and_grads(
            inp["user_embed"].double(), inp["item_embed"].double(), inp["target"].double(), item_idx=inp["item_idx"],
            pos_idx=inp["pos_idx"], num_negatives=k, sigma=1.7, margin=0.4, names=(n,),
        )[n]
        assert float(got[0]) == pytest.approx(float(loss), rel=1e-12), n
        assert rel_err(got[1], dq) < 1e-12, n
        if dv is not None:
            assert rel_err(got[2], dv) < 1e-12, n


def test_log_q_none_is_identity_and_log_q_shifts_logits() -> None:
    from xfmr_b200 import synthetic  # noqa: PLC0415

    inp = synthetic.make_loss_inputs(16, 40, 8, 3, n_catalog=30, seed=5)
    kw = {"item_idx": inp["item_idx"], "pos_idx": inp["pos_idx"]}
    base = losses_oracle.all_losses(inp["user_embed"], inp["item_embed"], inp["target"], **kw)
    zero = losses_oracle.all_losses(inp["user_embed"], inp["item_embed"], inp["target"], log_q=torch.zeros(40), **kw)
    const = losses_oracle.all_losses(inp["user_embed"], inp["item_embed"], inp["target"], log_q=torch.full((40,), 0.7), **kw)
    for n in LOSS_NAMES:
        assert torch.allclose(base[n], zero[n])
    # a constant shift of every logit leaves the softmax-type and pairwise losses unchanged
    for n in ("InfomationNoiseContrastiveEstimationLoss", "MutualInformationNeuralEstimationLoss", "PairwiseHingeLoss"):
        assert torch.allclose(base[n], const[n], rtol=1e-5, atol=1e-4)


# SURVEY.md Appendix C known answers: XXH32(le64(id), seed)
XXH_KAT = {
    0: (3736311059, 3521805802),
    1: (149775153, 1416521076),
    2: (3926170682, 4143920284),
    3706: (3480504050, 593379227),
    87585: (120160372, 1113714540),
    2**31: (1658944635, 1566471002),
    2**40 + 7: (2934185114, 3032242776),
    99999999: (4193919821, 4022736060),
}


def test_xxh32_known_answers() -> None:
    for value, (s0, s1) in XXH_KAT.items():
        assert native.xxh32_i64(value, 0) == s0
        assert native.xxh32_i64(value, 1) == s1
    assert native.hash_indices(np.array([3706]), 2, 22).tolist() == [[3426034, 1982363]]


def test_xxh32_matches_xxhash_wheel() -> None:
    xxhash = pytest.importorskip("xxhash")
    rng = np.random.default_rng(0)
    values = rng.integers(-(2**63), 2**63 - 1, 5000, dtype=np.int64)
    for v in values:
        for seed in (0, 1, 12345):
            assert native.xxh32_i64(int(v), seed) == xxhash.xxh32_intdigest(struct.pack("<q", int(v)), seed)


def test_topk_oracle_against_numpy_sort() -> None:
    rng = np.random.default_rng(1)
    q = rng.standard_normal((7, 24)).astype(np.float32)
    items = rng.standard_normal((300, 24)).astype(np.float32)
    items[17] = items[5]  # an exact tie: the lower id must come first
    ids = np.arange(1, 301, dtype=np.int64)
    excl = np.full((7, 3), native.PAD_ID, dtype=np.int64)
    excl[:, 0] = ids[np.argmax(q @ items.T, axis=1)]  # exclude each query's best item
    scores, out = native.topk(q, items, 10, item_ids=ids, exclude=excl)
    full = q.astype(np.float64) @ items.astype(np.float64).T
    for r in range(7):
        s = full[r].astype(np.float32)
        order = sorted((j for j in range(300) if ids[j] != excl[r, 0]), key=lambda j: (-s[j], ids[j]))[:10]
        assert out[r].tolist() == [int(ids[j]) for j in order]
        assert np.array_equal(scores[r], s[order])
    # fewer eligible items than k
    s2, i2 = native.topk(q[:1], items[:3], 5)
    assert i2[0, 3:].tolist() == [-1, -1] and np.isneginf(s2[0, 3:]).all()
