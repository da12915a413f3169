"""Helpers shared by the test modules."""

from __future__ import annotations

import pathlib

import numpy as np
import torch

GOLDEN_DIR = pathlib.Path(__file__).resolve().parent / "golden"
LOSS_NAMES = (
    "AlignmentLoss",
    "ContrastiveLoss",
    "AlignmentContrastiveLoss",
    "InfomationNoiseContrastiveEstimationLoss",
    "MutualInformationNeuralEstimationLoss",
    "PairwiseHingeLoss",
    "PairwiseLogisticLoss",
)


def golden_cases() -> list[str]:
    return sorted(p.stem for p in GOLDEN_DIR.glob("losses_*.npz"))


def load_golden(name: str) -> dict:
    z = np.load(GOLDEN_DIR / f"{name}.npz")
    k, sigma, margin = z["config"][:3]
    mining = "hard" if len(z["config"]) > 3 and z["config"][3] else "semi_hard"  # noqa: PLR2004
    case = {
        "user_embed": torch.from_numpy(z["user_embed"]),
        "item_embed": torch.from_numpy(z["item_embed"]),
        "target": torch.from_numpy(z["target"]),
        "item_idx": torch.from_numpy(z["item_idx"]),
        "pos_idx": torch.from_numpy(z["pos_idx"]),
        "num_negatives": int(k),
        "sigma": float(sigma),
        "margin": float(margin),
        "mining": mining,
        "expected": {},
    }
    for n in LOSS_NAMES:
        case["expected"][n] = (
            float(z[f"{n}.loss64"]),
            torch.from_numpy(z[f"{n}.d_user"]),
            torch.from_numpy(z[f"{n}.d_item"]),
        )
    return case


def rel_err(got: torch.Tensor, ref: torch.Tensor) -> float:
    """Norm-wise relative error ||got - ref|| / ||ref|| (absolute when ref is ~0)."""
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    denom = ref.norm().item()
    diff = (got - ref).norm().item()
    return diff if denom < 1e-12 else diff / denom  # noqa: PLR2004


def bf16_round(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)
