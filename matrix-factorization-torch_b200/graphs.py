"""CUDA-graph replay of a loss step with a pipelined host-input feed.

The C-ABI entry points allocate nothing and never synchronise (``include/xfmr_b200.h``), so a module's
forward + backward is capturable as one CUDA graph.  ``GraphedLossStep`` does the capture once for a fixed
input shape and then serves steps from either device tensors or pinned host tensors:

* host inputs are uploaded on a copy stream into one of three input slots, each with its own captured graph, so the
  upload of step ``i + 1`` runs under the graph replay of step ``i`` (what a training loop's prefetching data loader
  does for the reference's ``training_step``, ``xfmr_rec/lightning.py:189-192``) and no staging copy is needed;
* the step's loss is copied back to pinned host memory asynchronously; ``StepResult.loss_value()`` waits for
  that copy only, so a caller can read step ``i - 1`` while step ``i`` is in flight.

Everything here is stream/graph plumbing around the module call; no arithmetic.
"""

from __future__ import annotations

from typing import TYPE_CHECKING

import torch

if TYPE_CHECKING:
    from .losses import EmbeddingLoss

_INPUT_KEYS = ("user_embed", "item_embed", "target", "item_idx", "pos_idx")


class StepResult:
    """Handle of one submitted step: the slot's output tensors + the host copy of its loss."""

    def __init__(self, step: GraphedLossStep, slot: int, done: torch.cuda.Event) -> None:
        self._step = step
        self._slot = slot
        self._done = done

    def loss_value(self) -> float:
        """Host value of this step's loss (waits for the step's device -> host copy, not for later steps)."""
        self._done.synchronize()
        return float(self._step._loss_host[self._slot])  # noqa: SLF001

    @property
    def d_user(self) -> torch.Tensor:
        """Gradient buffer of the slot (overwritten when the slot comes round again, ``NUM_SLOTS`` steps later)."""
        return self._step._outputs[self._slot][1]  # noqa: SLF001

    @property
    def d_item(self) -> torch.Tensor:
        return self._step._outputs[self._slot][2]  # noqa: SLF001


class GraphedLossStep:
    """``module(user_embed, item_embed, target, item_idx=, pos_idx=)`` + backward as a CUDA graph per input slot.

    ``example`` fixes shapes and dtypes (a dict with the five input tensors on the GPU).  ``submit(inputs)``
    takes a dict of the same tensors on the GPU or in (pinned) host memory.  There are ``NUM_SLOTS`` input slots, each
    with its own captured graph reading the slot's tensors in place (no device-to-device staging copy): uploads land
    in the slot that was used ``NUM_SLOTS`` steps ago while the graphs of the slots in between run.  Results stay
    valid until their slot is reused.
    """

    NUM_SLOTS = 3

    def __init__(self, module: EmbeddingLoss, example: dict[str, torch.Tensor]) -> None:
        device = example["user_embed"].device
        if device.type != "cuda":
            msg = "GraphedLossStep needs CUDA tensors as the example inputs; there is no CPU path"
            raise RuntimeError(msg)
        self.module = module
        self.device = device
        n = self.NUM_SLOTS
        self._inputs = [{k: example[k].detach().clone() for k in _INPUT_KEYS} for _ in range(n)]
        for slot in self._inputs:
            slot["user_embed"].requires_grad_(True)
            slot["item_embed"].requires_grad_(True)
        self._uploaded = [torch.cuda.Event() for _ in range(n)]
        self._consumed: list[torch.cuda.Event | None] = [None] * n
        self._loss_host = torch.zeros(n, dtype=torch.float32).pin_memory()
        self._copy_stream = torch.cuda.Stream(device=device)
        self._next_slot = 0
        self._pending: int | None = None

        def step(s: dict[str, torch.Tensor]) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
            loss = module(s["user_embed"], s["item_embed"], s["target"], item_idx=s["item_idx"], pos_idx=s["pos_idx"])
            dq, dv = torch.autograd.grad(loss, (s["user_embed"], s["item_embed"]))
            return loss, dq, dv

        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                step(self._inputs[0])
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self._graphs: list[torch.cuda.CUDAGraph] = []
        self._outputs: list[tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = []
        pool = None
        for slot in range(n):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, pool=pool):
                self._outputs.append(step(self._inputs[slot]))
            pool = graph.pool()   # the graphs never run concurrently: they share one memory pool
            self._graphs.append(graph)
        torch.cuda.synchronize(device)

    def prefetch(self, inputs: dict[str, torch.Tensor]) -> None:
        """Start moving ``inputs`` into the next input slot on the copy stream (returns at once)."""
        if self._pending is not None:
            msg = "a prefetched step is already waiting: call submit() first"
            raise RuntimeError(msg)
        slot = self._next_slot
        consumed = self._consumed[slot]
        # Device-resident inputs were (or are being) written on the caller's stream: the copy stream must wait for that
        # work, and the caching allocator must not hand their memory out again while the copy is in flight.
        on_device = [inputs[k] for k in _INPUT_KEYS if inputs[k].is_cuda]
        if on_device:
            self._copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._copy_stream), torch.no_grad():
            if consumed is not None:
                self._copy_stream.wait_event(consumed)  # the step that last ran on this slot has finished
            for k in _INPUT_KEYS:
                self._inputs[slot][k].copy_(inputs[k], non_blocking=True)
            for t in on_device:
                t.record_stream(self._copy_stream)
            self._uploaded[slot].record(self._copy_stream)
        self._pending = slot
        self._next_slot = (slot + 1) % self.NUM_SLOTS

    def submit(self, inputs: dict[str, torch.Tensor] | None = None) -> StepResult:
        """Run one step on ``inputs`` (or on the inputs given to the last ``prefetch``); returns at once."""
        if self._pending is None:
            if inputs is None:
                msg = "nothing to run: give inputs or call prefetch() first"
                raise RuntimeError(msg)
            self.prefetch(inputs)
        slot = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._uploaded[slot])
        self._graphs[slot].replay()
        loss = self._outputs[slot][0]
        self._loss_host[slot : slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        done = torch.cuda.Event()
        done.record(cur)
        self._consumed[slot] = done   # the graph reads the slot's tensors throughout the step
        return StepResult(self, slot, done)
