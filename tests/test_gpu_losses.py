"""Parity of the fused CUDA losses (through the drop-in modules, i.e. through the C ABI) with the oracle
and with the committed outputs of the unmodified reference.

Tolerances are the north star's: losses and gradients within rel 1e-3 (fp32 accumulation); gradients are
compared norm-wise, ``||got - ref|| / ||ref||``.  In bf16 mode the oracle is fed the bf16-rounded inputs.
"""

from __future__ import annotations

import pytest
import torch

from helpers import LOSS_NAMES, bf16_round, golden_cases, load_golden, rel_err

pytestmark = pytest.mark.gpu

RTOL = 1e-3


def cuda_losses_and_grads(inp: dict, *, num_negatives: int, sigma: float, margin: float, dtype=torch.float32,
                          compute=None, log_q=None, names=LOSS_NAMES, mining="semi_hard") -> dict:
    import xfmr_b200  # noqa: PLC0415

    dev = torch.device("cuda:0")
    out = {}
    for n in names:
        module = getattr(xfmr_b200, n)(num_negatives=num_negatives, sigma=sigma, margin=margin, compute=compute,
                                       mining=mining)
        q = inp["user_embed"].to(dev, dtype).requires_grad_(True)
        v = inp["item_embed"].to(dev, dtype).requires_grad_(True)
        loss = module(q, v, inp["target"].to(dev), item_idx=inp["item_idx"].to(dev), pos_idx=inp["pos_idx"].to(dev),
                      **({"log_q": log_q.to(dev)} if log_q is not None else {}))
        assert loss.dim() == 0 and loss.dtype == torch.float32 and loss.is_cuda
        if torch.isfinite(loss):
            dq, dv = torch.autograd.grad(loss, (q, v))
        else:
            dq, dv = torch.zeros_like(q), torch.zeros_like(v)
        assert dq.dtype == dtype and dv.dtype == dtype
        out[n] = (loss.detach().cpu(), dq.float().cpu(), dv.float().cpu())
    return out


def oracle_losses_and_grads(inp: dict, *, num_negatives: int, sigma: float, margin: float, round_bf16=False,
                            log_q=None, device="cpu", names=LOSS_NAMES, mining="semi_hard") -> dict:
    from oracle import losses_oracle  # noqa: PLC0415

    q, v = inp["user_embed"], inp["item_embed"]
    if round_bf16:
        q, v = bf16_round(q), bf16_round(v)
    return losses_oracle.losses_and_grads(
        q.to(device).double(), v.to(device).double(), inp["target"].to(device).double(),
        item_idx=inp["item_idx"].to(device), pos_idx=inp["pos_idx"].to(device), num_negatives=num_negatives,
        sigma=sigma, margin=margin, log_q=None if log_q is None else log_q.to(device).double(), names=names,
        mining=mining,
    )


def assert_close(got: dict, ref: dict, rtol: float = RTOL, label: str = "") -> None:
    for n, (loss, dq, dv) in got.items():
        rloss, rdq, rdv = ref[n]
        rloss = float(rloss)
        if rloss != rloss or abs(rloss) == float("inf"):
            assert float(loss) != float(loss) or float(loss) == rloss, (label, n)
            continue
        e_loss = abs(float(loss) - rloss) / max(abs(rloss), 1e-6)
        e_dq, e_dv = rel_err(dq, rdq), rel_err(dv, rdv)
        print(f"{label:28s} {n:44s} loss {float(loss):14.6f} rel {e_loss:.2e}  dQ {e_dq:.2e}  dI {e_dv:.2e}")
        assert e_loss < rtol, (label, n, float(loss), rloss)
        assert e_dq < rtol, (label, n, "d_user")
        assert e_dv < rtol, (label, n, "d_item")


@pytest.mark.parametrize("name", golden_cases())
def test_reference_golden_fp32(name: str) -> None:
    case = load_golden(name)
    got = cuda_losses_and_grads(case, num_negatives=case["num_negatives"], sigma=case["sigma"], margin=case["margin"],
                                mining=case["mining"])
    assert_close(got, case["expected"], label=name)


@pytest.mark.parametrize("name", ["losses_dense_unit", "losses_ragged_tile", "losses_mined_k4"])
def test_reference_golden_bf16_io(name: str) -> None:
    """bf16 inputs / bf16 gradients; the oracle sees the bf16-rounded inputs, gradient tolerance is the bf16 ulp."""
    case = load_golden(name)
    got = cuda_losses_and_grads(case, num_negatives=case["num_negatives"], sigma=case["sigma"], margin=case["margin"],
                                dtype=torch.bfloat16)
    ref = oracle_losses_and_grads(case, num_negatives=case["num_negatives"], sigma=case["sigma"], margin=case["margin"],
                                  round_bf16=True)
    for n, (loss, dq, dv) in got.items():
        rloss, rdq, rdv = ref[n]
        assert abs(float(loss) - float(rloss)) <= RTOL * max(abs(float(rloss)), 1e-6), n
        assert rel_err(dq, rdq) < 5e-3, n  # output rounding to bf16: 2^-9 per element
        assert rel_err(dv, rdv) < 5e-3, n


@pytest.mark.parametrize(("b", "n", "d", "p", "k", "sigma", "margin", "signed", "normalize"), [
    (1, 1, 8, 0, 0, 1.0, 1.0, False, True),          # degenerate: a single pair, no negatives at all
    (5, 9, 16, 1, 0, 1.0, 1.0, False, True),
    (128, 128, 64, 4, 0, 1.0, 1.0, False, True),      # exactly one tile
    (129, 257, 64, 4, 0, 2.0, 0.5, True, True),       # one past the tile edge in both directions
    (300, 2000, 128, 16, 0, 10.0, 0.2, True, True),   # several column chunks
    (300, 2000, 128, 16, 8, 10.0, 0.2, False, True),  # mining across chunks
    (300, 9000, 64, 8, 64, 4.0, 0.3, True, True),     # mining at the largest supported K, signed targets, ragged tiles
    (30, 70, 32, 2, 64, 2.0, 0.5, False, True),       # K close to N: nearly every column is kept
    (64, 1500, 96, 8, 0, 1.0, 1.0, False, False),     # un-normalised embeddings exercise the norm terms
    (256, 700, 40, 40, 0, 30.0, -1.0, True, True),    # P > N/20, large sigma, negative margin
    (100, 32768, 32, 4, 0, 2.0, 0.5, False, True),    # one query row block: 64 column chunks (> one factor per lane)
])
def test_against_oracle_fp32(b, n, d, p, k, sigma, margin, signed, normalize) -> None:  # noqa: ANN001, PLR0913
    from xfmr_b200 import synthetic  # noqa: PLC0415

    inp = synthetic.make_loss_inputs(b, n, d, p, n_catalog=max(n // 2, 4), seed=b + n, signed_targets=signed,
                                     normalize=normalize, scale=1.0 if normalize else 0.4, mean_extra_pos=p / 2 + 0.5)
    got = cuda_losses_and_grads(inp, num_negatives=k, sigma=sigma, margin=margin)
    ref = oracle_losses_and_grads(inp, num_negatives=k, sigma=sigma, margin=margin)
    assert_close(got, ref, label=f"B{b} N{n} d{d} K{k}")


@pytest.mark.parametrize(("signed", "dtype"), [(False, torch.float32), (True, torch.float32), (False, torch.bfloat16)])
def test_many_item_row_blocks_and_column_chunks(signed: bool, dtype: torch.dtype) -> None:
    """20,000 items = 157 item row blocks (> 148 SMs): the persistent item-major sweep walks several row blocks per
    CTA; 256 users = 2 query row blocks, so the query-major sweeps split the items into 40 column chunks (> 32: the
    per-chunk scaling of the merged forward + dQ sweep leaves its one-factor-per-lane path)."""
    from xfmr_b200 import synthetic  # noqa: PLC0415

    inp = synthetic.make_loss_inputs(256, 20000, 64, 8, n_catalog=9000, seed=5, signed_targets=signed)
    got = cuda_losses_and_grads(inp, num_negatives=0, sigma=4.0, margin=0.3, dtype=dtype)
    ref = oracle_losses_and_grads(inp, num_negatives=0, sigma=4.0, margin=0.3, round_bf16=dtype == torch.bfloat16)
    if dtype == torch.float32:
        assert_close(got, ref, label=f"N20000 signed={signed}")
    else:
        for n, (loss, dq, dv) in got.items():
            rloss, rdq, rdv = ref[n]
            assert abs(float(loss) - float(rloss)) <= RTOL * max(abs(float(rloss)), 1e-6), n
            assert rel_err(dq, rdq) < 5e-3, n  # output rounding to bf16: 2^-9 per element
            assert rel_err(dv, rdv) < 5e-3, n


@pytest.mark.parametrize("sigma", [60.0, 300.0, 1000.0])
def test_large_sigma_takes_the_fallback_sweeps(sigma: float) -> None:
    """The merged forward + dQ sweep fixes each row's exponent reference after one tile; with logits spread over
    thousands of log2 units (sigma up to 1000 in the reference's HPO range, flaml.py:73-79) later tiles overflow it and
    the device-side fallback (separate forward and dQ sweeps) must deliver the same numbers.  At sigma >= 300 the
    softmax of a row is one-hot to fp32 precision, so the gradient is a single G_ij v_j term and carries the full
    rounding of the bf16 gradient tile (2^-9 relative, uniform): the gradient tolerance there is 2^-9, the loss
    tolerance stays 1e-3."""
    from xfmr_b200 import synthetic  # noqa: PLC0415

    inp = synthetic.make_loss_inputs(260, 3000, 64, 8, n_catalog=1500, seed=11, signed_targets=False)
    names = ("InfomationNoiseContrastiveEstimationLoss", "MutualInformationNeuralEstimationLoss")
    got = cuda_losses_and_grads(inp, num_negatives=0, sigma=sigma, margin=0.5, names=names)
    ref = oracle_losses_and_grads(inp, num_negatives=0, sigma=sigma, margin=0.5, names=names)
    assert_close(got, ref, rtol=RTOL if sigma < 100 else 2.0 ** -9, label=f"sigma{sigma}")
    for n in names:
        assert abs(float(got[n][0]) - float(ref[n][0])) <= RTOL * abs(float(ref[n][0])), n


def test_fused_call_equals_single_loss_calls_and_is_linear_in_upstream() -> None:
    import xfmr_b200  # noqa: PLC0415
    from xfmr_b200 import synthetic  # noqa: PLC0415

    dev = torch.device("cuda:0")
    inp = {k: v.to(dev) for k, v in synthetic.make_loss_inputs(200, 900, 64, 8, n_catalog=400, seed=3).items()}
    kw = {"item_idx": inp["item_idx"], "pos_idx": inp["pos_idx"], "sigma": 2.0, "margin": 0.7}
    q = inp["user_embed"].clone().requires_grad_(True)
    v = inp["item_embed"].clone().requires_grad_(True)
    fused = xfmr_b200.fused_losses(q, v, inp["target"], **kw)
    weights = torch.tensor([0.3, -1.2, 0.5, 2.0, 0.7, 1.1, -0.4], device=dev)
    dq, dv = torch.autograd.grad((fused * weights).sum(), (q, v))
    dq_sum, dv_sum = torch.zeros_like(dq), torch.zeros_like(dv)
    for n, slot in xfmr_b200.LOSS_SLOTS.items():
        q1 = inp["user_embed"].clone().requires_grad_(True)
        v1 = inp["item_embed"].clone().requires_grad_(True)
        single = getattr(xfmr_b200, n)(sigma=2.0, margin=0.7)(q1, v1, inp["target"], item_idx=inp["item_idx"], pos_idx=inp["pos_idx"])
        assert float(single) == pytest.approx(float(fused[slot]), rel=1e-5), n
        g = torch.autograd.grad(single, (q1, v1))
        dq_sum += weights[slot] * g[0]
        dv_sum += weights[slot] * g[1]
    assert rel_err(dq, dq_sum) < 1e-3
    assert rel_err(dv, dv_sum) < 1e-3


def test_log_q_extension_matches_oracle() -> None:
    from xfmr_b200 import synthetic  # noqa: PLC0415

    inp = synthetic.make_loss_inputs(150, 800, 64, 6, n_catalog=300, seed=11)
    got = cuda_losses_and_grads(inp, num_negatives=0, sigma=3.0, margin=0.5, log_q=inp["log_q"])
    ref = oracle_losses_and_grads(inp, num_negatives=0, sigma=3.0, margin=0.5, log_q=inp["log_q"])
    assert_close(got, ref, label="log_q")
    got_k = cuda_losses_and_grads(inp, num_negatives=5, sigma=3.0, margin=0.5, log_q=inp["log_q"])
    ref_k = oracle_losses_and_grads(inp, num_negatives=5, sigma=3.0, margin=0.5, log_q=inp["log_q"])
    assert_close(got_k, ref_k, label="log_q K=5")
    # signed targets flip the logit sign per row; hard mining with LogQ exercises the one-sided vote on -R
    signed = synthetic.make_loss_inputs(150, 800, 64, 6, n_catalog=300, seed=12, signed_targets=True)
    for mining in ("semi_hard", "hard"):
        got_s = cuda_losses_and_grads(signed, num_negatives=7, sigma=3.0, margin=0.5, log_q=signed["log_q"], mining=mining)
        ref_s = oracle_losses_and_grads(signed, num_negatives=7, sigma=3.0, margin=0.5, log_q=signed["log_q"], mining=mining)
        assert_close(got_s, ref_s, label=f"log_q K=7 signed {mining}")


def test_config1_movielens_1m_shape_fp32() -> None:
    """BASELINE config 1 at full size: 1,024 x 3,706, d=64 fp32, P=32, oracle on the GPU in float64."""
    from xfmr_b200 import synthetic  # noqa: PLC0415

    inp = synthetic.make_loss_inputs(1024, 3706, 64, 32, n_catalog=3706, seed=0)
    for k in (0, 4):
        got = cuda_losses_and_grads(inp, num_negatives=k, sigma=1.0, margin=1.0)
        ref = oracle_losses_and_grads(inp, num_negatives=k, sigma=1.0, margin=1.0, device="cuda:0")
        ref = {n: tuple(t.cpu() for t in v) for n, v in ref.items()}
        assert_close(got, ref, label=f"C1 K={k}")


def test_config2_movielens_32m_shape_bf16() -> None:
    """BASELINE config 2 at full size (4,096 x 87,585, d=128, bf16): every loss against the float64 oracle
    evaluated on the GPU over the bf16-rounded inputs (P=8 keeps the reference's B x N x P broadcast in memory)."""
    from xfmr_b200 import synthetic  # noqa: PLC0415

    inp = synthetic.make_loss_inputs(4096, 87585, 128, 8, n_catalog=87585, seed=2, device="cuda:0")
    inp = {k: v.cpu() for k, v in inp.items()}
    names = ("ContrastiveLoss", "InfomationNoiseContrastiveEstimationLoss", "PairwiseLogisticLoss")
    got = cuda_losses_and_grads(inp, num_negatives=0, sigma=5.0, margin=0.5, dtype=torch.bfloat16, names=names)
    torch.cuda.empty_cache()
    for n in names:
        ref = oracle_losses_and_grads(inp, num_negatives=0, sigma=5.0, margin=0.5, round_bf16=True, device="cuda:0", names=(n,))
        rloss, rdq, rdv = (t.cpu() for t in ref[n])
        loss, dq, dv = got[n]
        print(f"C2 {n}: loss {float(loss):.4f} vs {float(rloss):.4f}; dQ {rel_err(dq, rdq):.2e} dI {rel_err(dv, rdv):.2e}")
        assert abs(float(loss) - float(rloss)) <= RTOL * abs(float(rloss)), n
        assert rel_err(dq, rdq) < 5e-3, n   # bf16 output rounding
        assert rel_err(dv, rdv) < 5e-3, n
        del ref
        torch.cuda.empty_cache()


def test_all_rows_masked_and_zero_targets() -> None:
    """Edge cases observed on the reference: a row whose every column is an accidental
    hit gives MINE = -inf and 0 for the mean losses; target 0 rows contribute nothing."""
    import xfmr_b200  # noqa: PLC0415

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(0)
    q = torch.nn.functional.normalize(torch.randn(4, 16, generator=gen), dim=-1).to(dev)
    v = torch.nn.functional.normalize(torch.randn(6, 16, generator=gen), dim=-1).to(dev)
    item_idx = torch.tensor([7, 7, 7, 7, 7, 7], device=dev)       # every column carries every row's positive id
    pos_idx = torch.zeros(4, 1, dtype=torch.int64, device=dev)
    target = torch.tensor([1.0, 2.0, 0.0, 3.0], device=dev)
    losses = xfmr_b200.fused_losses(q, v, target, item_idx=item_idx, pos_idx=pos_idx)
    assert torch.isinf(losses[4]) or torch.isnan(losses[4])
    for slot in (1, 5, 6):
        assert losses[slot].item() == 0.0
    assert losses[3].item() == pytest.approx(0.0, abs=1e-5)       # softmax over the positive alone
    # zero targets: the whole batch contributes nothing
    z = xfmr_b200.fused_losses(q, v, torch.zeros(4, device=dev), item_idx=torch.arange(1, 7, device=dev), pos_idx=pos_idx)
    assert torch.all(z == 0)


def test_graphed_step_matches_eager_and_takes_host_inputs() -> None:
    """``GraphedLossStep``: CUDA-graph replay with the double-buffered host feed gives the eager results, step after step."""
    import xfmr_b200  # noqa: PLC0415
    from xfmr_b200 import synthetic  # noqa: PLC0415

    dev = torch.device("cuda:0")
    batches = [synthetic.make_loss_inputs(300, 1000, 64, 6, n_catalog=800, seed=s) for s in range(4)]
    module = xfmr_b200.InfomationNoiseContrastiveEstimationLoss(sigma=3.0)
    keys = ("user_embed", "item_embed", "target", "item_idx", "pos_idx")
    stepper = xfmr_b200.GraphedLossStep(module, {k: batches[0][k].to(dev) for k in keys})
    results = []
    hosts = [{k: b[k].pin_memory() for k in keys} for b in batches]
    stepper.prefetch(hosts[0])
    for i in range(len(batches)):
        res = stepper.submit()
        got_dq, got_dv = res.d_user.clone(), res.d_item.clone()   # static buffers: copy before the next step
        if i + 1 < len(batches):
            stepper.prefetch(hosts[i + 1])
        results.append((res.loss_value(), got_dq, got_dv))
    for b, (loss, dq, dv) in zip(batches, results, strict=True):
        q = b["user_embed"].to(dev).requires_grad_(True)
        v = b["item_embed"].to(dev).requires_grad_(True)
        want = module(q, v, b["target"].to(dev), item_idx=b["item_idx"].to(dev), pos_idx=b["pos_idx"].to(dev))
        wq, wv = torch.autograd.grad(want, (q, v))
        assert loss == pytest.approx(float(want.detach()), rel=1e-6)
        assert torch.equal(dq, wq)
        assert torch.equal(dv, wv)
    # device inputs go through the same path
    res = stepper.submit({k: batches[1][k].to(dev) for k in keys})
    assert res.loss_value() == pytest.approx(results[1][0], rel=1e-6)
    # ... also when they have just been computed on the caller's stream (the copy stream must wait for the producer and
    # the allocator must keep the temporaries alive): a long-running producer right before the submit
    big = torch.randn(4096, 4096, device=dev)
    for _ in range(5):
        big = torch.tanh(big @ big * 1e-2)
    fresh = {k: (batches[2][k].to(dev) * 1 if batches[2][k].is_floating_point() else batches[2][k].to(dev) + 0) for k in keys}
    fresh["user_embed"] = fresh["user_embed"] + big[:300, :64] * 0.0
    res = stepper.submit(fresh)
    del fresh
    assert res.loss_value() == pytest.approx(results[2][0], rel=1e-6)


@pytest.mark.parametrize(("k", "dtype"), [(4, torch.float32), (32, torch.float32), (8, torch.bfloat16)])
def test_hard_mining_larger_shapes(k: int, dtype: torch.dtype) -> None:
    """``mining="hard"`` (``hard_mining``, losses.py:112-132) on shapes with several row blocks and column chunks."""
    from xfmr_b200 import synthetic  # noqa: PLC0415

    inp = synthetic.make_loss_inputs(700, 5000, 64, 6, n_catalog=3000, seed=40 + k, signed_targets=True, mean_extra_pos=2.0)
    bf16 = dtype == torch.bfloat16
    got = cuda_losses_and_grads(inp, num_negatives=k, sigma=3.0, margin=0.4, dtype=dtype, mining="hard")
    ref = oracle_losses_and_grads(inp, num_negatives=k, sigma=3.0, margin=0.4, round_bf16=bf16, mining="hard")
    assert_close(got, ref, rtol=2.0**-7 if bf16 else RTOL, label=f"hard k={k}")


def test_mining_keyword_is_validated() -> None:
    import xfmr_b200  # noqa: PLC0415

    with pytest.raises(ValueError, match="mining"):
        xfmr_b200.PairwiseHingeLoss(num_negatives=4, mining="easy")
