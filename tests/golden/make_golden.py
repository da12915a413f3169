"""Generates tests/golden/losses_*.npz from the UNMODIFIED reference (``/root/reference/xfmr_rec/losses.py``).

Run in the build container (the reference is not present on the GPU box):
    python tests/golden/make_golden.py
Every case stores the inputs, the configuration and, for each of the seven reference loss classes, the
loss value and its gradients w.r.t. ``user_embed`` / ``item_embed`` evaluated by the reference in float64
(and the float32 loss values for orientation).  Reference torch version is recorded in each file.
"""

from __future__ import annotations

import pathlib
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[2]))

import xfmr_rec.losses as ref  # noqa: E402

import xfmr_b200  # noqa: E402, F401  (only for the synthetic generator)
from xfmr_b200 import synthetic  # noqa: E402

OUT = pathlib.Path(__file__).resolve().parent
NAMES = [
    "AlignmentLoss",
    "ContrastiveLoss",
    "AlignmentContrastiveLoss",
    "InfomationNoiseContrastiveEstimationLoss",
    "MutualInformationNeuralEstimationLoss",
    "PairwiseHingeLoss",
    "PairwiseLogisticLoss",
]

CASES = [
    # name, B, N, d, P, n_catalog, K, sigma, margin, signed, normalize, scale
    ("dense_unit", 48, 112, 32, 5, 60, 0, 1.0, 1.0, False, True, 1.0),
    ("dense_signed_sigma", 48, 112, 32, 5, 60, 0, 2.5, 0.3, True, True, 1.0),
    ("dense_unnormalised", 40, 100, 64, 8, 80, 0, 1.0, 0.5, False, False, 0.35),
    ("mined_k4", 48, 112, 32, 5, 60, 4, 1.0, 1.0, False, True, 1.0),
    ("mined_k4_signed", 48, 112, 32, 5, 60, 4, 2.5, 0.3, True, True, 1.0),
    ("square_no_negs", 32, 32, 32, 3, 200, 0, 1.0, 1.0, False, True, 1.0),
    ("ragged_tile", 130, 300, 48, 6, 150, 0, 4.0, -0.5, True, True, 1.0),
    ("mined_k16_ragged", 130, 300, 48, 6, 150, 16, 4.0, 0.25, False, True, 1.0),
    # names starting with "hard": the reference's own (never-called) hard_mining method (losses.py:112-132) is bound in
    # place of semi_hard_mining on each module instance; the classes themselves stay unmodified
    ("hard_k4", 48, 112, 32, 5, 60, 4, 1.0, 1.0, False, True, 1.0),
    ("hard_k16_ragged_signed", 130, 300, 48, 6, 150, 16, 4.0, 0.25, True, True, 1.0),
]


def run_reference(inp: dict[str, torch.Tensor], k: int, sigma: float, margin: float, dtype: torch.dtype, hard: bool = False):
    out = {}
    for name in NAMES:
        module = getattr(ref, name)(num_negatives=k, sigma=sigma, margin=margin)
        if hard:
            module.semi_hard_mining = module.hard_mining
        q = inp["user_embed"].to(dtype).clone().requires_grad_(True)
        v = inp["item_embed"].to(dtype).clone().requires_grad_(True)
        loss = module(q, v, inp["target"].to(dtype), item_idx=inp["item_idx"], pos_idx=inp["pos_idx"])
        if torch.isfinite(loss):
            dq, dv = torch.autograd.grad(loss, (q, v), allow_unused=True)
        else:
            dq, dv = None, None
        dq = torch.zeros_like(q) if dq is None else dq
        dv = torch.zeros_like(v) if dv is None else dv
        out[name] = (loss.detach(), dq, dv)
    return out


def main(only: set[str] | None = None) -> None:
    for i, (name, b, n, d, p, ncat, k, sigma, margin, signed, normalize, scale) in enumerate(CASES):
        inp = synthetic.make_loss_inputs(
            b, n, d, p, n_catalog=ncat, seed=1000 + i, signed_targets=signed, normalize=normalize, scale=scale,
            mean_extra_pos=2.0,
        )
        if only is not None and name not in only:
            continue
        hard = name.startswith("hard")
        r64 = run_reference(inp, k, sigma, margin, torch.float64, hard)
        r32 = run_reference(inp, k, sigma, margin, torch.float32, hard)
        arrays = {
            "user_embed": inp["user_embed"].numpy(),
            "item_embed": inp["item_embed"].numpy(),
            "target": inp["target"].numpy(),
            "item_idx": inp["item_idx"].numpy(),
            "pos_idx": inp["pos_idx"].numpy(),
            "config": np.array([k, sigma, margin, 1.0 if hard else 0.0], dtype=np.float64),
            "torch_version": np.array(torch.__version__),
        }
        for lname in NAMES:
            arrays[f"{lname}.loss64"] = r64[lname][0].numpy()
            arrays[f"{lname}.loss32"] = r32[lname][0].numpy()
            arrays[f"{lname}.d_user"] = r64[lname][1].numpy().astype(np.float32)
            arrays[f"{lname}.d_item"] = r64[lname][2].numpy().astype(np.float32)
        np.savez_compressed(OUT / f"losses_{name}.npz", **arrays)
        print(name, {ln: float(r64[ln][0]) for ln in NAMES})


if __name__ == "__main__":
    main(set(sys.argv[1:]) or None)   # optional: case names to (re)generate; the committed files of the others stay
