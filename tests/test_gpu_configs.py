"""Parity of the CUDA path at the BASELINE.json configurations that ``bench.py`` times.

Everything benchmarked is checked here at the benchmarked size: config 2 at P = 32 for all seven losses (and the
reference's default training loss, ``PairwiseHingeLoss(num_negatives=4)``, ``xfmr_rec/lightning.py:38-42``), config 3's
d = 256 pipeline, the false-negative mask bit for bit, and the retrieval recall at catalog scale.  The oracle runs in
float64 on the GPU (same restatement of ``xfmr_rec/losses.py``, CUDA tensors) because the ``B x N x P`` accidental-hit
broadcast of ``losses.py:108`` is 11.5 GB at config 2.

Tolerances are the north star's: losses and gradients rel 1e-3 (gradients norm-wise), mask bit-exact,
bf16 retrieval recall@k >= 0.999.
"""

from __future__ import annotations

import pytest
import torch

from helpers import LOSS_NAMES, bf16_round, rel_err
from test_gpu_losses import assert_close, cuda_losses_and_grads, oracle_losses_and_grads

pytestmark = pytest.mark.gpu

RTOL = 1e-3
HINGE = "PairwiseHingeLoss"
INFONCE = "InfomationNoiseContrastiveEstimationLoss"


def _oracle_on_gpu(inp: dict, names, **kw) -> dict:  # noqa: ANN001, ANN003
    """One loss at a time (each autograd graph holds several B x N float64 matrices)."""
    out = {}
    for n in names:
        ref = oracle_losses_and_grads(inp, device="cuda:0", names=(n,), **kw)
        out[n] = tuple(t.cpu() for t in ref[n])
        del ref
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ config 2
@pytest.fixture(scope="module")
def c2_inputs() -> dict:
    from xfmr_b200 import synthetic  # noqa: PLC0415

    inp = synthetic.make_loss_inputs(4096, 87585, 128, 32, n_catalog=87585, seed=2, device="cuda:0")
    inp = {k: v.cpu() for k, v in inp.items()}
    # the tensor cores see bf16: give both sides the bf16-rounded values in fp32 storage, so that the gradients come
    # back in fp32 and the comparison is not blurred by a bf16 output rounding
    inp["user_embed"] = bf16_round(inp["user_embed"])
    inp["item_embed"] = bf16_round(inp["item_embed"])
    return inp


@pytest.mark.parametrize("name", LOSS_NAMES)
def test_config2_p32_every_loss_1e3(c2_inputs: dict, name: str) -> None:
    """BASELINE config 2 exactly as ``bench.py`` runs it (4,096 x 87,585, d=128, P=32, sigma=5, margin=0.5, bf16 tensor
    path), fp32 in/out: losses AND gradients within 1e-3."""
    got = cuda_losses_and_grads(c2_inputs, num_negatives=0, sigma=5.0, margin=0.5, compute="bf16", names=(name,))
    torch.cuda.empty_cache()
    ref = _oracle_on_gpu(c2_inputs, (name,), num_negatives=0, sigma=5.0, margin=0.5)
    assert_close(got, ref, rtol=RTOL, label="C2 P=32")


@pytest.mark.parametrize("mining", ["semi_hard", "hard"])
def test_config2_p32_default_training_loss_k4(c2_inputs: dict, mining: str) -> None:
    """``PairwiseHingeLoss(num_negatives=4)`` - the loss the reference trains with (lightning.py:38-42) - at config 2."""
    kw = {"num_negatives": 4, "sigma": 5.0, "margin": 0.5, "mining": mining}
    got = cuda_losses_and_grads(c2_inputs, compute="bf16", names=(HINGE, INFONCE), **kw)
    torch.cuda.empty_cache()
    ref = _oracle_on_gpu(c2_inputs, (HINGE, INFONCE), **kw)
    assert_close(got, ref, rtol=RTOL, label=f"C2 K=4 {mining}")


# ------------------------------------------------------------------------------------------------ config 3 (d = 256)
@pytest.mark.parametrize("d", [192, 256])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_wide_embeddings_every_loss(d: int, dtype: torch.dtype) -> None:
    """d > 128 takes another pipeline shape (fewer TMA stages, fewer TMEM score buffers, a 256-column accumulator)."""
    from xfmr_b200 import synthetic  # noqa: PLC0415

    inp = synthetic.make_loss_inputs(300, 2000, d, 8, n_catalog=900, seed=d, signed_targets=True)
    inp["user_embed"], inp["item_embed"] = bf16_round(inp["user_embed"]), bf16_round(inp["item_embed"])
    for k in (0, 4):
        got = cuda_losses_and_grads(inp, num_negatives=k, sigma=4.0, margin=0.3, dtype=dtype, compute="bf16")
        ref = oracle_losses_and_grads(inp, num_negatives=k, sigma=4.0, margin=0.3)
        # bf16 gradients carry their own 2^-9 output rounding
        assert_close(got, ref, rtol=RTOL if dtype == torch.float32 else 5e-3, label=f"d{d} K={k} {dtype}")


def test_config3_per_rank_shape_d256() -> None:
    """BASELINE config 3 as one rank sees it: 8,192 users x (8,192 in-batch + 16,384 uniform) candidates, d=256, bf16
    tensor path; sampled softmax and the pairwise hinge loss against the float64 oracle."""
    from xfmr_b200 import synthetic  # noqa: PLC0415

    inp = synthetic.make_loss_inputs(8192, 24576, 256, 32, n_catalog=200_000, seed=50, device="cuda:0")
    inp = {k: v.cpu() for k, v in inp.items()}
    inp["user_embed"], inp["item_embed"] = bf16_round(inp["user_embed"]), bf16_round(inp["item_embed"])
    names = (INFONCE, HINGE)
    got = cuda_losses_and_grads(inp, num_negatives=0, sigma=5.0, margin=0.5, compute="bf16", names=names)
    torch.cuda.empty_cache()
    ref = _oracle_on_gpu(inp, names, num_negatives=0, sigma=5.0, margin=0.5)
    assert_close(got, ref, rtol=RTOL, label="C3 d=256")
    # the dtype the bench runs: bf16 in, bf16 gradients out (2^-9 output rounding on top)
    got16 = cuda_losses_and_grads(inp, num_negatives=0, sigma=5.0, margin=0.5, dtype=torch.bfloat16, names=(INFONCE,))
    loss, dq, dv = got16[INFONCE]
    rloss, rdq, rdv = ref[INFONCE]
    assert abs(float(loss) - float(rloss)) <= RTOL * abs(float(rloss))
    assert rel_err(dq, rdq) < 5e-3
    assert rel_err(dv, rdv) < 5e-3


# ------------------------------------------------------------------------------------------------ mask, bit for bit
def _unpack_bits(words: torch.Tensor, nbits: int) -> torch.Tensor:
    """[rows, words] int32 bit matrix -> [rows, nbits] bool (bit c of a row = word c // 32, bit c % 32)."""
    w = words.to(torch.int64) & 0xFFFFFFFF
    shifts = torch.arange(32, device=words.device, dtype=torch.int64)
    bits = ((w.unsqueeze(-1) >> shifts) & 1).bool().reshape(words.size(0), -1)
    return bits[:, :nbits]


@pytest.mark.parametrize(("b", "n", "p", "n_catalog"), [
    (1, 1, 0, 4),
    (130, 300, 1, 40),         # heavy duplication: every id sits in several columns
    (513, 4000, 32, 900),
    (96, 1000, 512, 3000),     # MovieLens-sized positive lists (up to ~1.8 k per user in the data)
    (1024, 3706, 32, 3706),    # config 1
])
def test_pair_mask_bit_exact_with_negative_masks(b: int, n: int, p: int, n_catalog: int) -> None:
    """``xb_build_pair_mask`` (mask AND transpose, with ``row_ids0``) == ``~negative_masks`` (losses.py:92-110), every bit,
    including the padding: rows >= B and columns >= N are all ones (excluded), pad id 0 matches nothing."""
    import xfmr_b200  # noqa: PLC0415
    from oracle import losses_oracle  # noqa: PLC0415
    from xfmr_b200 import synthetic  # noqa: PLC0415

    dev = torch.device("cuda:0")
    inp = synthetic.make_loss_inputs(b, n, 8, p, n_catalog=n_catalog, seed=b + p, mean_extra_pos=max(p / 2, 0.5))
    item_idx, pos_idx = inp["item_idx"].to(dev), inp["pos_idx"].to(dev)
    mask, mask_t = xfmr_b200.build_pair_mask(item_idx, pos_idx if p > 0 else None, row_ids0=item_idx[:b], transpose=True)
    want = ~losses_oracle.negative_mask(item_idx, pos_idx if p > 0 else None, b)       # [B, N] True = excluded
    rows_pad, cols_pad = mask.size(0), mask_t.size(0)
    got = _unpack_bits(mask, mask.size(1) * 32)
    assert torch.equal(got[:b, :n], want)
    assert bool(got[b:].all()), "padding rows must be excluded"
    assert bool(got[:b, n:].all()), "padding columns must be excluded"
    got_t = _unpack_bits(mask_t, mask_t.size(1) * 32)
    assert torch.equal(got_t[:n, :b], want.t())
    assert bool(got_t[n:].all())
    assert bool(got_t[:n, b:].all())
    assert rows_pad % 128 == 0 and cols_pad % 128 == 0


def test_pair_mask_inside_the_loss_workspace_is_the_same_mask() -> None:
    """The mask the loss sweeps read (workspace region 2 of ``xb_debug_loss_region``) is the standalone builder's."""
    import ctypes  # noqa: PLC0415

    import xfmr_b200  # noqa: PLC0415
    from oracle import losses_oracle  # noqa: PLC0415
    from xfmr_b200 import _lib, synthetic  # noqa: PLC0415
    _loss_fwd, _make_desc = xfmr_b200.losses._loss_fwd, xfmr_b200.losses._make_desc  # noqa: SLF001

    dev = torch.device("cuda:0")
    b, n, p = 200, 1500, 16
    inp = {k: v.to(dev) for k, v in synthetic.make_loss_inputs(b, n, 64, p, n_catalog=500, seed=9).items()}
    slot = xfmr_b200.LOSS_SLOTS[INFONCE]
    _, ws = _loss_fwd(inp["user_embed"], inp["item_embed"], inp["target"], inp["item_idx"], inp["pos_idx"], None, 0, 1.0,
                      1.0, 1 << slot, _lib.XB_COMPUTE_BF16, 0)
    desc = _make_desc(b, n, 64, p, _lib.XB_DTYPE_F32, _lib.XB_COMPUTE_BF16, 0, 1 << slot, 1.0, 1.0, False)
    off, nbytes = ctypes.c_size_t(0), ctypes.c_size_t(0)
    _lib.check(_lib.lib.xb_debug_loss_region(ctypes.byref(desc), 2, ctypes.byref(off), ctypes.byref(nbytes)), "region")
    words = _lib.lib.xb_mask_words(n)
    region = ws[off.value: off.value + nbytes.value].view(torch.int32).reshape(-1, words)
    got = _unpack_bits(region, words * 32)
    want = ~losses_oracle.negative_mask(inp["item_idx"], inp["pos_idx"], b)
    assert torch.equal(got[:b, :n], want)


# ------------------------------------------------------------------------------------------------ retrieval at scale
def test_retrieval_recall_at_ten_million_items() -> None:
    """bf16 retrieval over 10^7 items (a tenth of config 5; candidate buffers, compaction slack and threshold staleness
    all scale with N): recall@100 >= 0.999 against a brute-force ranking of the same bf16-rounded inputs, for a random
    sample of the queries; ties at the k-th score are not counted against either side."""
    import xfmr_b200  # noqa: PLC0415
    from xfmr_b200 import synthetic  # noqa: PLC0415

    dev = torch.device("cuda:0")
    n, q, d, k = 10_000_000, 4096, 128, 100
    items = synthetic.make_catalog(n, d, seed=100, device=dev, dtype=torch.bfloat16)
    queries = synthetic.make_catalog(q, d, seed=7, device=dev, dtype=torch.bfloat16)
    scores, ids = xfmr_b200.topk_search(queries, items, k, id_base=0)
    assert bool((ids >= 0).all()) and bool((ids < n).all())
    assert bool((scores[:, :-1] >= scores[:, 1:]).all()), "result lists must be sorted by score"
    sample = torch.randperm(q, generator=torch.Generator().manual_seed(1))[:64].to(dev)
    hits = total = 0
    items_f = items.float()
    for qi in sample.tolist():
        exact = items_f @ queries[qi].float()                  # fp32 products of the bf16 values
        kth = torch.topk(exact, k).values[-1]
        must = exact > kth                                      # strictly above the k-th score: must be returned
        got = torch.zeros(n, dtype=torch.bool, device=dev)
        got[ids[qi]] = True
        hits += int((must & got).sum())
        total += int(must.sum())
        assert int(got.sum()) == k, "duplicate ids in a result list"
        assert bool((exact[ids[qi]] >= kth - 1e-3).all()), "a returned item is far below the k-th best"
    assert hits >= 0.999 * total, (hits, total)
