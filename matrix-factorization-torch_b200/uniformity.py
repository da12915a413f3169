"""Uniformity, DirectAU and MAWU losses on the same score-path kernels.

The reference only *cites* these (``README.md:22-25``: DirectAU arXiv 2206.12811, MAWU arXiv 2308.06091);
``xfmr_rec/losses.py`` has the alignment half (``AlignmentLoss``, losses.py:249-259) and no uniformity term.
Parity of everything in this file is therefore **unpinned by the reference**; the oracle
(``oracle/losses_oracle.py:uniformity`` / ``directau`` / ``mawu``) restates the published definitions.

``uniformity(x, t) = log mean_{i != j} exp(-t |x_i - x_j|^2)`` is the Gram contraction ``X.X^T`` pushed through
the sweep kernel of the losses (``xb_uniformity_forward`` / ``_backward``): rows = columns = ``x``, only the
diagonal masked, an all-pairs logsumexp instead of a per-row one.  No ``n x n`` matrix reaches HBM and the
gradient costs one sweep (the forward already carries the unnormalised row-side gradient; the column side is
its mirror image).
"""

from __future__ import annotations

import ctypes

import torch

from . import _lib
from .losses import LOSS_SLOTS, EmbeddingLoss, fused_losses


def _make_desc(n: int, dim: int, in_dtype: int, compute: int, t: float) -> _lib.UniformityDesc:
    return _lib.UniformityDesc(n=n, dim=dim, in_dtype=in_dtype, compute=compute, t=t, reserved=0)


def _workspace_bytes(desc: _lib.UniformityDesc) -> int:
    nbytes = _lib.lib.xb_uniformity_workspace_bytes(ctypes.byref(desc))
    if nbytes == 0:
        raise _lib.XbError("xb_uniformity_workspace_bytes: " + _lib.lib.xb_last_error_string().decode())
    return int(nbytes)


@torch.library.custom_op("xfmr_b200::uniformity_fwd", mutates_args=())
def _uniformity_fwd(x: torch.Tensor, t: float, compute: int) -> tuple[torch.Tensor, torch.Tensor]:
    device = _lib.require_cuda(x)
    desc = _make_desc(x.size(0), x.size(1), _lib.dtype_code(x.dtype), compute, t)
    ws_bytes = _workspace_bytes(desc)
    with torch.cuda.device(device):
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
        loss = torch.empty(1, dtype=torch.float32, device=device)
        status = _lib.lib.xb_uniformity_forward(
            ctypes.byref(desc), x.data_ptr(), loss.data_ptr(), workspace.data_ptr(), ws_bytes, _lib.stream_ptr(device)
        )
    _lib.check(status, "xb_uniformity_forward")
    return loss, workspace


@_uniformity_fwd.register_fake
def _(x, t, compute):  # noqa: ANN001, ANN202
    desc = _make_desc(x.size(0), x.size(1), _lib.dtype_code(x.dtype), compute, t)
    return x.new_empty(1, dtype=torch.float32), x.new_empty(_workspace_bytes(desc), dtype=torch.uint8)


@torch.library.custom_op("xfmr_b200::uniformity_bwd", mutates_args=("workspace",))
def _uniformity_bwd(
    workspace: torch.Tensor, d_loss: torch.Tensor, n: int, dim: int, bf16_io: bool, t: float, compute: int
) -> torch.Tensor:
    device = _lib.require_cuda(workspace, d_loss)
    dtype = torch.bfloat16 if bf16_io else torch.float32
    desc = _make_desc(n, dim, _lib.dtype_code(dtype), compute, t)
    with torch.cuda.device(device):
        d_x = torch.empty(n, dim, dtype=dtype, device=device)
        status = _lib.lib.xb_uniformity_backward(
            ctypes.byref(desc), d_loss.data_ptr(), d_x.data_ptr(), workspace.data_ptr(), workspace.numel(),
            _lib.stream_ptr(device),
        )
    _lib.check(status, "xb_uniformity_backward")
    return d_x


@_uniformity_bwd.register_fake
def _(workspace, d_loss, n, dim, bf16_io, t, compute):  # noqa: ANN001, ANN202, PLR0913
    return workspace.new_empty(n, dim, dtype=torch.bfloat16 if bf16_io else torch.float32)


def _setup_context(ctx, inputs, output) -> None:  # noqa: ANN001
    x, t, compute = inputs
    _loss, workspace = output
    ctx.mark_non_differentiable(workspace)
    ctx.set_materialize_grads(False)
    ctx.save_for_backward(workspace)
    ctx.meta = (x.size(0), x.size(1), x.dtype == torch.bfloat16, t, compute)


def _backward(ctx, d_loss, _d_workspace):  # noqa: ANN001, ANN202
    (workspace,) = ctx.saved_tensors
    if d_loss is None:
        return None, None, None
    return _uniformity_bwd(workspace, d_loss.contiguous().float(), *ctx.meta), None, None


_uniformity_fwd.register_autograd(_backward, setup_context=_setup_context)


def uniformity_loss(x: torch.Tensor, t: float = 2.0, *, compute: str | None = None) -> torch.Tensor:
    """``log mean_{i != j} exp(-t |x_i - x_j|^2)`` over the rows of ``x [n, d]`` (0-d fp32, differentiable).

    Wang & Isola 2020 / DirectAU's ``uniformity`` with ``torch.pdist`` (ordered and unordered pairs give the
    same mean).  ``x`` is used as given — the model of the reference already emits unit-norm rows
    (``xfmr_rec/models.py:59``).
    """
    if x.dim() != 2:  # noqa: PLR2004
        msg = f"x must be [n, d], got {tuple(x.shape)}"
        raise ValueError(msg)
    if x.size(0) < 2:  # noqa: PLR2004
        msg = f"uniformity needs at least two rows, got {x.size(0)}"
        raise ValueError(msg)
    _lib.require_cuda(x)
    if x.dtype not in (torch.float32, torch.bfloat16):
        msg = f"embeddings must be float32 or bfloat16, got {x.dtype}"
        raise TypeError(msg)
    loss, _workspace = _uniformity_fwd(x.contiguous(), float(t), _lib.compute_code(compute, x.dtype))
    return loss[0]


class UniformityLoss(torch.nn.Module):
    """``uniformity_loss`` as a module (one embedding matrix in, 0-d loss out)."""

    def __init__(self, *, t: float = 2.0, compute: str | None = None) -> None:
        super().__init__()
        self.t = t
        self.compute = compute

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return uniformity_loss(x, self.t, compute=self.compute)


class DirectAULoss(EmbeddingLoss):
    """DirectAU (arXiv 2206.12811): ``mean_i |u_i - v_i|^2 + gamma * (U(users) + U(items)) / 2``.

    Call signature of the reference's loss modules (losses.py:39-52).  ``target`` weights the alignment pairs
    like ``AlignmentLoss`` does (losses.py:164-170) — with unit targets this is DirectAU's published loss;
    ``item_idx`` / ``pos_idx`` are accepted and unused (DirectAU has no negatives); the uniformity terms run
    over the batch users and the in-batch items ``item_embed[:B]``.
    """

    def __init__(self, *, gamma: float = 1.0, t: float = 2.0, compute: str | None = None, **kwargs: object) -> None:
        super().__init__(compute=compute, **kwargs)
        self.gamma = gamma
        self.t = t

    def alignment(self, user_embed, item_embed, target, *, item_idx, pos_idx):  # noqa: ANN001, ANN201
        # sum_i |u_i - v_i|^2 / 2 * target_i  (alignment slot of the fused kernel, sigma = 1)  ->  mean |.|^2
        losses = fused_losses(
            user_embed, item_embed, target, item_idx=item_idx, pos_idx=pos_idx, sigma=1.0,
            loss_mask=1 << LOSS_SLOTS["AlignmentLoss"], compute=self.compute,
        )
        return losses[LOSS_SLOTS["AlignmentLoss"]] * (2.0 / user_embed.size(0))

    def weighted_uniformity(self, user_embed, item_embed, gamma_user: float, gamma_item: float):  # noqa: ANN001, ANN201
        batch = user_embed.size(0)
        total = user_embed.new_zeros((), dtype=torch.float32)
        if gamma_user != 0.0:
            total = total + gamma_user * uniformity_loss(user_embed, self.t, compute=self.compute)
        if gamma_item != 0.0:
            total = total + gamma_item * uniformity_loss(item_embed[:batch], self.t, compute=self.compute)
        return total

    def loss(self, user_embed, item_embed, target, *, item_idx, pos_idx, log_q=None):  # noqa: ANN001, ANN201, ARG002
        align = self.alignment(user_embed, item_embed, target, item_idx=item_idx, pos_idx=pos_idx)
        return align + self.weighted_uniformity(user_embed, item_embed, self.gamma / 2, self.gamma / 2)


class MAWULoss(DirectAULoss):
    """MAWU (arXiv 2308.06091): margin-aware alignment + weighted uniformity.

    * weighted uniformity: ``gamma_user * U(users) + gamma_item * U(items)`` — separate weights for the two sides;
    * margin-aware alignment: ``mean_i (2 - 2 cos(theta_i + m_u[i] + m_v[i]))`` with ``theta_i`` the angle between
      ``u_i`` and ``v_i`` and per-row (learnable) margins passed as ``user_margin`` / ``item_margin``; with zero
      margins it is DirectAU's ``|u_i - v_i|^2`` for unit-norm rows.  This restates the paper from memory (no
      network here) — unpinned.  It is diagonal-only, O(B d) elementwise work, and is written with torch ops so
      the margins get gradients; the contraction-shaped part (the uniformities) is the kernel path above.
    """

    def __init__(
        self, *, gamma_user: float = 1.0, gamma_item: float = 1.0, t: float = 2.0, compute: str | None = None,
        **kwargs: object,
    ) -> None:
        super().__init__(gamma=gamma_user + gamma_item, t=t, compute=compute, **kwargs)
        self.gamma_user = gamma_user
        self.gamma_item = gamma_item

    def forward(self, user_embed, item_embed, target, *, item_idx, pos_idx, log_q=None,  # noqa: ANN001, ANN201, PLR0913
                user_margin=None, item_margin=None):  # noqa: ANN001
        self.check_inputs(user_embed, item_embed, target)
        _lib.require_cuda(user_embed, item_embed, target, user_margin, item_margin)   # no CPU path, whatever the weights
        if user_margin is None and item_margin is None:
            align = self.alignment(user_embed, item_embed, target, item_idx=item_idx, pos_idx=pos_idx)
        else:
            batch = user_embed.size(0)
            u = user_embed.float()
            v = item_embed[:batch].float()
            cos = (u * v).sum(-1) / (u.norm(dim=-1) * v.norm(dim=-1)).clamp_min(1e-12)
            theta = torch.acos(cos.clamp(-1 + 1e-6, 1 - 1e-6))
            if user_margin is not None:
                theta = theta + user_margin
            if item_margin is not None:
                theta = theta + item_margin
            align = ((2 - 2 * torch.cos(theta)) * target.float()).sum() / batch
        return align + self.weighted_uniformity(user_embed, item_embed, self.gamma_user, self.gamma_item)
