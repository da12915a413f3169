"""xfmr_b200 — B200 (sm_100a) implementation of the query-by-item score path of ``xfmr_rec``.

Importable as ``xfmr_b200`` (shim module at the repository root; this directory's name has hyphens).
Public surface, mirroring the reference interfaces for this path:

* the seven loss modules of ``xfmr_rec/losses.py`` + ``fused_losses`` (one contraction for any subset),
* ``ItemProcessor`` / ``topk_search`` — exact top-k with the semantics of ``ItemProcessor.search``,
* ``uniformity_loss`` / ``DirectAULoss`` / ``MAWULoss`` — the uniformity family the reference's README cites,
* ``ItemProcessor.evaluate`` / ``retrieval_metrics`` — batched validation (search + the six ranking metrics),
* ``GraphedLossStep`` — a loss step as one CUDA graph with a pipelined host-input feed,
* ``hash_embedding_gather`` / ``HashEmbeddingBag`` — the hashed-embedding feeder,
* ``distributed`` — global negatives (training) and catalog sharding (retrieval).
"""

from . import distributed
from ._lib import LIB_PATH, XbError, launch_count
from .graphs import GraphedLossStep, StepResult
from .hashing import HashEmbeddingBag, hash_embedding_gather, hash_indices
from .losses import (
    ALL_LOSSES,
    LOSS_CLASSES,
    LOSS_SLOTS,
    AlignmentContrastiveLoss,
    AlignmentLoss,
    ContrastiveLoss,
    EmbeddingLoss,
    InfomationNoiseContrastiveEstimationLoss,
    MutualInformationNeuralEstimationLoss,
    PairwiseHingeLoss,
    PairwiseLogisticLoss,
    fused_losses,
)
from .retrieval import (
    METRIC_NAMES,
    TOP_K,
    ItemProcessor,
    arrow_to_catalog,
    build_pair_mask,
    retrieval_metrics,
    topk_filter,
    topk_merge,
    topk_search,
)
from .uniformity import DirectAULoss, MAWULoss, UniformityLoss, uniformity_loss

__all__ = [
    "ALL_LOSSES",
    "METRIC_NAMES",
    "LIB_PATH",
    "LOSS_CLASSES",
    "LOSS_SLOTS",
    "TOP_K",
    "AlignmentContrastiveLoss",
    "AlignmentLoss",
    "ContrastiveLoss",
    "DirectAULoss",
    "EmbeddingLoss",
    "GraphedLossStep",
    "HashEmbeddingBag",
    "InfomationNoiseContrastiveEstimationLoss",
    "ItemProcessor",
    "MAWULoss",
    "MutualInformationNeuralEstimationLoss",
    "PairwiseHingeLoss",
    "PairwiseLogisticLoss",
    "StepResult",
    "UniformityLoss",
    "XbError",
    "arrow_to_catalog",
    "build_pair_mask",
    "distributed",
    "fused_losses",
    "hash_embedding_gather",
    "hash_indices",
    "launch_count",
    "retrieval_metrics",
    "topk_filter",
    "topk_merge",
    "topk_search",
    "uniformity_loss",
]
