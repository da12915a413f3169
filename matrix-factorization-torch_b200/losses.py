"""Drop-in loss modules for ``xfmr_rec.losses`` backed by the fused sm_100a score-path kernels.

Same seven class names, constructor keywords and ``forward`` signature as the reference
(``xfmr_rec/losses.py:26-52, 249-359``): ``cls(num_negatives=, sigma=, margin=)`` and
``loss(user_embed, item_embed, target, *, item_idx=, pos_idx=) -> 0-d tensor`` that takes part in
autograd (gradients flow to ``user_embed`` and ``item_embed`` only — masks and mining carry none,
``losses.py:92, 112, 134``).  The caller's loop over all seven modules
(``xfmr_rec/lightning.py:137-146``) keeps working; :func:`fused_losses` evaluates any subset in ONE
contraction instead.

Two keyword-only extensions, both off by default: ``log_q`` (LogQ / sampling-bias correction, subtracts
``log_q[j]`` from every logit) and ``compute`` (``"bf16"`` to round fp32 inputs to bf16 for the tensor
cores; by default fp32 inputs use the split-bf16 contraction which is accurate to ~2^-16).

Everything below is plumbing around ``xb_loss_forward`` / ``xb_loss_backward``; no loss arithmetic
lives in Python and there is no CPU implementation.
"""

from __future__ import annotations

import abc
import ctypes

import torch

from . import _lib

# slots of the fused output vector (include/xfmr_b200.h)
LOSS_SLOTS = {
    "AlignmentLoss": 0,
    "ContrastiveLoss": 1,
    "AlignmentContrastiveLoss": 2,
    "InfomationNoiseContrastiveEstimationLoss": 3,
    "MutualInformationNeuralEstimationLoss": 4,
    "PairwiseHingeLoss": 5,
    "PairwiseLogisticLoss": 6,
}
ALL_LOSSES = (1 << _lib.XB_NUM_LOSSES) - 1


def _make_desc(
    batch: int,
    num_items: int,
    dim: int,
    num_pos: int,
    in_dtype: int,
    compute: int,
    num_negatives: int,
    loss_mask: int,
    sigma: float,
    margin: float,
    has_log_q: bool,
    mining: int = 0,
) -> _lib.LossDesc:
    return _lib.LossDesc(
        batch=batch,
        num_items=num_items,
        dim=dim,
        num_pos=num_pos,
        in_dtype=in_dtype,
        compute=compute,
        num_negatives=num_negatives,
        loss_mask=loss_mask,
        sigma=sigma,
        margin=margin,
        has_log_q=int(has_log_q),
        mining=mining,
    )


def _workspace_bytes(desc: _lib.LossDesc) -> int:
    nbytes = _lib.lib.xb_loss_workspace_bytes(ctypes.byref(desc))
    if nbytes == 0:
        raise _lib.XbError("xb_loss_workspace_bytes: " + _lib.lib.xb_last_error_string().decode())
    return int(nbytes)


@torch.library.custom_op("xfmr_b200::loss_fwd", mutates_args=())
def _loss_fwd(
    user_embed: torch.Tensor,
    item_embed: torch.Tensor,
    target: torch.Tensor,
    item_idx: torch.Tensor,
    pos_idx: torch.Tensor,
    log_q: torch.Tensor | None,
    num_negatives: int,
    sigma: float,
    margin: float,
    loss_mask: int,
    compute: int,
    mining: int,
) -> tuple[torch.Tensor, torch.Tensor]:
    device = _lib.require_cuda(user_embed, item_embed, target, item_idx, pos_idx, log_q)
    desc = _make_desc(
        user_embed.size(0),
        item_embed.size(0),
        user_embed.size(1),
        pos_idx.size(1),
        _lib.dtype_code(user_embed.dtype),
        compute,
        num_negatives,
        loss_mask,
        sigma,
        margin,
        log_q is not None,
        mining,
    )
    ws_bytes = _workspace_bytes(desc)
    with torch.cuda.device(device):
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        losses = torch.empty(_lib.XB_NUM_LOSSES, dtype=torch.float32, device=device)
        status = _lib.lib.xb_loss_forward(
            ctypes.byref(desc),
            user_embed.data_ptr(),
            item_embed.data_ptr(),
            target.data_ptr(),
            item_idx.data_ptr(),
            pos_idx.data_ptr() if pos_idx.numel() else None,
            _lib.ptr(log_q),
            losses.data_ptr(),
            workspace.data_ptr(),
            ws_bytes,
            _lib.stream_ptr(device),
        )
    _lib.check(status, "xb_loss_forward")
    return losses, workspace


@_loss_fwd.register_fake
def _(user_embed, item_embed, target, item_idx, pos_idx, log_q, num_negatives, sigma, margin, loss_mask, compute, mining):  # noqa: ANN001, ANN202, PLR0913
    desc = _make_desc(
        user_embed.size(0),
        item_embed.size(0),
        user_embed.size(1),
        pos_idx.size(1),
        _lib.dtype_code(user_embed.dtype),
        compute,
        num_negatives,
        loss_mask,
        sigma,
        margin,
        log_q is not None,
        mining,
    )
    return (
        user_embed.new_empty(_lib.XB_NUM_LOSSES, dtype=torch.float32),
        user_embed.new_empty(_workspace_bytes(desc), dtype=torch.uint8),
    )


@torch.library.custom_op("xfmr_b200::loss_bwd", mutates_args=("workspace",))
def _loss_bwd(
    workspace: torch.Tensor,
    d_losses: torch.Tensor,
    batch: int,
    num_items: int,
    dim: int,
    num_pos: int,
    bf16_io: bool,
    has_log_q: bool,
    num_negatives: int,
    sigma: float,
    margin: float,
    loss_mask: int,
    compute: int,
    mining: int,
) -> tuple[torch.Tensor, torch.Tensor]:
    device = _lib.require_cuda(workspace, d_losses)
    dtype = torch.bfloat16 if bf16_io else torch.float32
    desc = _make_desc(
        batch, num_items, dim, num_pos, _lib.dtype_code(dtype), compute, num_negatives, loss_mask, sigma, margin, has_log_q,
        mining,
    )
    with torch.cuda.device(device):
        d_user = torch.empty(batch, dim, dtype=dtype, device=device)
        d_item = torch.empty(num_items, dim, dtype=dtype, device=device)
        status = _lib.lib.xb_loss_backward(
            ctypes.byref(desc),
            d_losses.data_ptr(),
            d_user.data_ptr(),
            d_item.data_ptr(),
            workspace.data_ptr(),
            workspace.numel(),
            _lib.stream_ptr(device),
        )
    _lib.check(status, "xb_loss_backward")
    return d_user, d_item


@_loss_bwd.register_fake
def _(workspace, d_losses, batch, num_items, dim, num_pos, bf16_io, has_log_q, num_negatives, sigma, margin, loss_mask, compute, mining):  # noqa: ANN001, ANN202, PLR0913
    dtype = torch.bfloat16 if bf16_io else torch.float32
    return workspace.new_empty(batch, dim, dtype=dtype), workspace.new_empty(num_items, dim, dtype=dtype)


def _setup_context(ctx, inputs, output) -> None:  # noqa: ANN001
    user_embed, item_embed, _target, _item_idx, pos_idx, log_q, num_negatives, sigma, margin, loss_mask, compute, mining = inputs
    _losses, workspace = output
    ctx.mark_non_differentiable(workspace)
    ctx.set_materialize_grads(False)  # never allocate a zero "gradient" for the workspace bytes
    ctx.save_for_backward(workspace)
    ctx.meta = (
        user_embed.size(0),
        item_embed.size(0),
        user_embed.size(1),
        pos_idx.size(1),
        user_embed.dtype == torch.bfloat16,
        log_q is not None,
        num_negatives,
        sigma,
        margin,
        loss_mask,
        compute,
        mining,
    )


def _backward(ctx, d_losses, _d_workspace):  # noqa: ANN001, ANN202
    (workspace,) = ctx.saved_tensors
    if d_losses is None:
        return (None,) * 12
    d_user, d_item = _loss_bwd(workspace, d_losses.contiguous().float(), *ctx.meta)
    return d_user, d_item, None, None, None, None, None, None, None, None, None, None


_loss_fwd.register_autograd(_backward, setup_context=_setup_context)


def check_inputs(user_embed: torch.Tensor, item_embed: torch.Tensor, target: torch.Tensor) -> None:
    """Shape checks with the error behaviour of ``EmbeddingLoss.check_inputs`` (losses.py:54-79)."""
    if user_embed.dim() != 2 or item_embed.dim() != 2:  # noqa: PLR2004
        msg = f"embeddings must be 2-d: user_embed.dim()={user_embed.dim()}, item_embed.dim()={item_embed.dim()}"
        raise ValueError(msg)
    if user_embed.size(1) != item_embed.size(1):
        msg = f"embedding widths differ: user_embed {user_embed.size(1)} vs item_embed {item_embed.size(1)}"
        raise ValueError(msg)
    if not (user_embed.size(0) == target.size(0) and item_embed.size(0) >= target.size(0)):
        msg = (
            "row counts do not line up: "
            f"target {target.size(0)}, user_embed {user_embed.size(0)}, item_embed {item_embed.size(0)}"
        )
        raise ValueError(msg)


def fused_losses(
    user_embed: torch.Tensor,
    item_embed: torch.Tensor,
    target: torch.Tensor,
    *,
    item_idx: torch.Tensor,
    pos_idx: torch.Tensor | None,
    log_q: torch.Tensor | None = None,
    num_negatives: int = 0,
    sigma: float = 1.0,
    margin: float = 1.0,
    loss_mask: int = ALL_LOSSES,
    compute: str | None = None,
    mining: str = "semi_hard",
) -> torch.Tensor:
    """All selected losses in one contraction: returns a float32 vector of 7 slots (``LOSS_SLOTS``).

    Unselected slots are 0.  Differentiable w.r.t. ``user_embed`` and ``item_embed``.  ``mining`` picks which
    ``num_negatives`` columns survive: ``"semi_hard"`` (``semi_hard_mining``, losses.py:134-162 — what every
    reference loss calls) or ``"hard"`` (``hard_mining``, losses.py:112-132, defined but never called there).
    """
    check_inputs(user_embed, item_embed, target)
    device = _lib.require_cuda(user_embed, item_embed, target, item_idx, pos_idx, log_q)
    if user_embed.dtype not in (torch.float32, torch.bfloat16):
        msg = f"embeddings must be float32 or bfloat16, got {user_embed.dtype}"
        raise TypeError(msg)
    if item_embed.dtype != user_embed.dtype:
        item_embed = item_embed.to(user_embed.dtype)
    batch = user_embed.size(0)
    if pos_idx is None:
        pos_idx = torch.empty(batch, 0, dtype=torch.int64, device=device)
    if pos_idx.dim() != 2 or pos_idx.size(0) != batch:  # noqa: PLR2004
        msg = f"pos_idx must be [batch, num_pos], got {tuple(pos_idx.shape)}"
        raise ValueError(msg)
    if item_idx.dim() != 1 or item_idx.size(0) != item_embed.size(0):
        msg = f"item_idx must be [num_items], got {tuple(item_idx.shape)}"
        raise ValueError(msg)
    if log_q is not None:
        if log_q.shape != item_idx.shape:
            msg = f"log_q must be [num_items], got {tuple(log_q.shape)}"
            raise ValueError(msg)
        log_q = log_q.detach().to(torch.float32).contiguous()
    losses, _workspace = _loss_fwd(
        user_embed.contiguous(),
        item_embed.contiguous(),
        target.detach().to(torch.float32).contiguous(),
        item_idx.detach().to(torch.int64).contiguous(),
        pos_idx.detach().to(torch.int64).contiguous(),
        log_q,
        int(num_negatives),
        float(sigma),
        float(margin),
        int(loss_mask),
        _lib.compute_code(compute, user_embed.dtype),
        _lib.mining_code(mining),
    )
    return losses


class EmbeddingLoss(torch.nn.Module, abc.ABC):
    """Base of the seven drop-in modules (reference: ``xfmr_rec/losses.py:26-52``)."""

    def __init__(
        self,
        *,
        num_negatives: int = 0,
        sigma: float = 1.0,
        margin: float = 1.0,
        compute: str | None = None,
        mining: str = "semi_hard",
    ) -> None:
        super().__init__()
        self.num_negatives = num_negatives
        self.sigma = sigma
        self.margin = margin
        self.compute = compute
        _lib.mining_code(mining)  # validates
        self.mining = mining

    @property
    def slot(self) -> int:
        return LOSS_SLOTS[type(self).__name__]

    def check_inputs(self, user_embed: torch.Tensor, item_embed: torch.Tensor, target: torch.Tensor) -> None:
        check_inputs(user_embed, item_embed, target)

    def forward(
        self,
        user_embed: torch.Tensor,
        item_embed: torch.Tensor,
        target: torch.Tensor,
        *,
        item_idx: torch.Tensor,
        pos_idx: torch.Tensor,
        log_q: torch.Tensor | None = None,
    ) -> torch.Tensor:
        self.check_inputs(user_embed, item_embed, target)
        return self.loss(user_embed, item_embed, target, item_idx=item_idx, pos_idx=pos_idx, log_q=log_q)

    def loss(
        self,
        user_embed: torch.Tensor,
        item_embed: torch.Tensor,
        target: torch.Tensor,
        *,
        item_idx: torch.Tensor,
        pos_idx: torch.Tensor,
        log_q: torch.Tensor | None = None,
    ) -> torch.Tensor:
        losses = fused_losses(
            user_embed,
            item_embed,
            target,
            item_idx=item_idx,
            pos_idx=pos_idx,
            log_q=log_q,
            num_negatives=self.num_negatives,
            sigma=self.sigma,
            margin=self.margin,
            loss_mask=1 << self.slot,
            compute=self.compute,
            mining=self.mining,
        )
        return losses[self.slot]


class AlignmentLoss(EmbeddingLoss):
    """``sum_i D_ii * target_i * sigma`` (losses.py:164-170, 249-259)."""


class ContrastiveLoss(EmbeddingLoss):
    """Masked mean of ``relu(L + sign*margin)`` per row (losses.py:172-193, 262-274)."""


class AlignmentContrastiveLoss(EmbeddingLoss):
    """Alignment + Contrastive (losses.py:277-291)."""


class InfomationNoiseContrastiveEstimationLoss(EmbeddingLoss):  # (sic) the reference's spelling
    """Sampled softmax over the masked negatives and the positive (losses.py:195-223, 294-306)."""


class MutualInformationNeuralEstimationLoss(EmbeddingLoss):
    """``-L_ii + logsumexp`` over the masked negatives (losses.py:225-246, 309-321)."""


class PairwiseHingeLoss(EmbeddingLoss):
    """Masked mean of ``relu(L_ij - L_ii + margin)`` (losses.py:325-346, 357-359)."""


class PairwiseLogisticLoss(EmbeddingLoss):
    """Masked mean of ``softplus(L_ij - L_ii + margin)`` — BPR (losses.py:325-346, 352-354)."""


LOSS_CLASSES = [
    AlignmentLoss,
    ContrastiveLoss,
    AlignmentContrastiveLoss,
    InfomationNoiseContrastiveEstimationLoss,
    MutualInformationNeuralEstimationLoss,
    PairwiseHingeLoss,
    PairwiseLogisticLoss,
]
