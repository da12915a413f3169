"""Benchmark of the B200 score path.  ``python bench.py --gpus N --steps K --warmup W``.

Headline metric (BASELINE.json): **fused loss forward+backward samples/s** on config 2 — MovieLens-32M
shaped, batch 4,096 x 87,585 candidate items, d=128, bf16, P=32 — for the sampled-softmax loss
(``InfomationNoiseContrastiveEstimationLoss``); the other six losses, the exact top-100 retrieval rate
(the second half of BASELINE.json's metric) and the hash-gather bandwidth ride along as extra keys of the
same JSON line.

A "step" is one pass of the hot path over one synthetic batch: ``loss(user_embed, item_embed, target,
item_idx=, pos_idx=)`` + ``backward()`` through the drop-in module, i.e. through the C ABI.
``value`` times the steps with device-resident inputs; ``e2e`` times the same call with HOST inputs
(pinned), copying them to the device and reading the loss back inside the timed region.

N > 1 (torchrun, one rank per GPU): the loss path shards by users with no data-path collective at this
config (every rank scores its own 4,096 users), so ``value`` = N x 4,096 / max-over-ranks time, "weak"
scaling; the catalog-sharded retrieval extra uses NCCL for the top-k merge.

``--impl reference`` times the CPU port of the reference algorithm (``oracle/losses_oracle.py``; the
reference itself is Python and is not present on the GPU box) on the host cores of rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import time

import torch

ROOT = pathlib.Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "fused_loss_fwd_bwd_samples_per_s"
UNIT = "samples/s"
C2 = {"batch": 4096, "num_items": 87585, "dim": 128, "num_pos": 32}
HEADLINE_LOSS = "InfomationNoiseContrastiveEstimationLoss"
SIGMA, MARGIN = 5.0, 0.5


def peaks() -> dict:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        data = json.loads(path.read_text())
        return {"hbm_gbs": data["hbm_gbs"], "bf16_tflops": data["bf16_tflops"],
                "bf16_tflops_sustained": data.get("bf16_tflops_sustained", data["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic() -> float | None:
    """dram__bytes_read.sum + dram__bytes_write.sum per sweep launch, from the committed `ncu --set full` capture of
    this same command (profiles/ncu_traffic.json, written by profiles/ncu_summarize.py); None if it is absent."""
    path = ROOT / "profiles" / "ncu_traffic.json"
    if not path.exists():
        return None
    return float(json.loads(path.read_text())["sweep_dram_bytes_per_launch"])


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled every few milliseconds from
    a thread between ``mark_begin`` and ``mark_end`` (the timed region lasts ~0.1 s, too short for ``nvidia-smi -lms``
    to be sure of a sample), plus one sample taken by the caller itself while the GPU is still busy."""

    # nvmlClocksEventReasons bits
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, gpu_index: int) -> None:
        import threading  # noqa: PLC0415

        self.handle = None
        self.nvml = None
        self.samples: list[tuple[float, int]] = []
        self.max_mhz = None
        self._run = threading.Event()
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._poll, daemon=True)
        try:
            import pynvml  # noqa: PLC0415

            pynvml.nvmlInit()
            props = torch.cuda.get_device_properties(gpu_index)
            try:
                bus = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:  # noqa: BLE001
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.handle = None

    def sample(self) -> None:
        if self.handle is None:
            return
        try:
            mhz = float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            try:
                bits = int(self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            except Exception:  # noqa: BLE001
                bits = int(self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            self.samples.append((mhz, bits))
        except Exception:  # noqa: BLE001, S110
            pass

    def _poll(self) -> None:
        while not self._stop.is_set():
            if self._run.is_set():
                self.sample()
            time.sleep(0.004)

    def start(self) -> None:
        if self.handle is not None:
            self._thread.start()

    def mark_begin(self) -> None:
        self.samples.clear()
        self._run.set()

    def mark_end(self) -> None:
        """Call while the last timed step is still in flight (before the final synchronize)."""
        self.sample()
        self._run.clear()

    def stop(self) -> None:
        self._stop.set()
        if self._thread.is_alive():
            self._thread.join(timeout=1)

    def summary(self) -> dict:
        sm = [m for m, _ in self.samples if m > 0]
        reasons = sorted({name for _, bits in self.samples for bit, name in self.REASONS.items() if bits & bit})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(self.samples), "source": "nvml"}


def flush_l2(buf: torch.Tensor) -> None:
    buf.add_(1)  # 256 MiB read + write > 126 MB of L2


def timed_steps(step_fn, steps: int, warmup: int, flush_buf: torch.Tensor | None) -> list[float]:  # noqa: ANN001
    """Per-step device milliseconds (CUDA events on the current stream); L2 flushed between steps."""
    for _ in range(warmup):
        if flush_buf is not None:
            flush_l2(flush_buf)
        step_fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(steps):
        if flush_buf is not None:
            flush_l2(flush_buf)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        step_fn()
        e1.record()
        e1.synchronize()
        times.append(e0.elapsed_time(e1))
    return times


def make_c2(device: torch.device, seed: int, dtype: torch.dtype) -> dict:
    from xfmr_b200 import synthetic  # noqa: PLC0415

    inp = synthetic.make_loss_inputs(C2["batch"], C2["num_items"], C2["dim"], C2["num_pos"], n_catalog=C2["num_items"],
                                     seed=seed, device=device)
    inp["user_embed"] = inp["user_embed"].to(dtype)
    inp["item_embed"] = inp["item_embed"].to(dtype)
    return inp


def loss_step_fn(module, inp: dict):  # noqa: ANN001, ANN201
    q = inp["user_embed"].detach().requires_grad_(True)
    v = inp["item_embed"].detach().requires_grad_(True)

    def step() -> tuple:
        loss = module(q, v, inp["target"], item_idx=inp["item_idx"], pos_idx=inp["pos_idx"])
        dq, dv = torch.autograd.grad(loss, (q, v))
        return loss, dq, dv

    return step


def graphed(step_fn):  # noqa: ANN001, ANN201
    """Capture one step (module forward + backward) into a CUDA graph; returns the replay callable.
    The library allocates nothing and never synchronises, so the whole step is capturable."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step_fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        keep = step_fn()
    torch.cuda.synchronize()
    graph.keep = keep  # outputs live in the graph's pool
    return graph.replay


def cpu_reference_rate(steps: int, warmup: int, name: str = HEADLINE_LOSS) -> dict:
    """The reference's algorithm (CPU port) on the host cores: fwd + autograd.grad on a bounded sample of C2:
    full batch, the first N_s candidate columns (the B x N x P accidental-hit broadcast of losses.py:108 is
    11.5 GB at full N), extrapolated linearly in N to the full candidate count."""
    from oracle import losses_oracle  # noqa: PLC0415
    from xfmr_b200 import synthetic  # noqa: PLC0415

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_sample = 8192
    inp = synthetic.make_loss_inputs(C2["batch"], n_sample, C2["dim"], C2["num_pos"], n_catalog=C2["num_items"], seed=0)

    def run() -> float:
        t0 = time.perf_counter()
        losses_oracle.losses_and_grads(inp["user_embed"], inp["item_embed"], inp["target"], item_idx=inp["item_idx"],
                                       pos_idx=inp["pos_idx"], sigma=SIGMA, margin=MARGIN, names=(name,))
        return time.perf_counter() - t0

    for _ in range(max(warmup, 1)):
        run()
    times = [run() for _ in range(max(steps, 1))]
    best = min(times)
    scale = C2["num_items"] / n_sample
    return {
        "value": C2["batch"] / (best * scale),
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": (f"{name} fwd+bwd, B={C2['batch']} x N_s={n_sample} of {C2['num_items']} items, d={C2['dim']}, "
                   f"P={C2['num_pos']}, fp32, best of {len(times)}: {best * 1e3:.0f} ms; rate scaled by N_s/N"),
        "ms_sample": best * 1e3,
    }


def aten_gpu_reference_rate(device: torch.device, name: str = HEADLINE_LOSS) -> dict:
    """The same restatement of the reference's algorithm with CUDA tensors: what the reference's ATen-composed loss costs
    on this very GPU (SURVEY.md 8d "ATen-on-GPU bar").  Part of the baseline leg: the port is the thing measured here,
    never the product.  Bounded sample like the CPU leg (the B x N x P mask broadcast is 11.5 GB at full N)."""
    from oracle import losses_oracle  # noqa: PLC0415
    from xfmr_b200 import synthetic  # noqa: PLC0415

    n_sample = 8192
    inp = synthetic.make_loss_inputs(C2["batch"], n_sample, C2["dim"], C2["num_pos"], n_catalog=C2["num_items"], seed=0,
                                     device=device)

    def run() -> None:
        losses_oracle.losses_and_grads(inp["user_embed"], inp["item_embed"], inp["target"], item_idx=inp["item_idx"],
                                       pos_idx=inp["pos_idx"], sigma=SIGMA, margin=MARGIN, names=(name,))

    ms = statistics.median(timed_steps(run, 5, 2, None))
    scale = C2["num_items"] / n_sample
    return {"value": C2["batch"] / (ms * 1e-3 * scale), "unit": UNIT, "kind": "port on cuda (ATen kernels, fp32)",
            "sample": f"{name} fwd+bwd, B={C2['batch']} x N_s={n_sample} of {C2['num_items']} items, d={C2['dim']}, P={C2['num_pos']}: "
                      f"{ms:.2f} ms; rate scaled by N_s/N"}


def run_reference(args: argparse.Namespace) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 5)
    base = cpu_reference_rate(steps, min(args.warmup, 1))
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": base["value"],
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": min(args.warmup, 1),
        "ms_per_step": base["ms_sample"] * C2["num_items"] / 8192,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "C2 MovieLens-32M-shaped: 4096 x 87585, d=128, P=32, sampled-softmax (InfoNCE) fwd+bwd",
                   "note": "CPU port of xfmr_rec/losses.py (oracle/losses_oracle.py); the Python reference is absent on the GPU box"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))  # noqa: T201


def bench_retrieval(device: torch.device, world: int, rank: int, num_items_total: int, num_queries: int, k: int) -> dict:
    """Exact top-k over a row-sharded catalog (config 5: 65,536 queries x 10^8 items, d=128 bf16, k=100; strong
    scaling - the catalog is split across the ranks, every rank scores all queries against its shard and the per-shard
    top-k lists are all-gathered and merged)."""
    import torch.distributed as dist  # noqa: PLC0415

    import xfmr_b200  # noqa: PLC0415
    from xfmr_b200 import synthetic  # noqa: PLC0415

    shard = num_items_total // world
    items = synthetic.make_catalog(shard, 128, seed=100 + rank, device=device, dtype=torch.bfloat16)
    queries = synthetic.make_catalog(num_queries, 128, seed=7, device=device, dtype=torch.bfloat16)

    def search(qs: torch.Tensor, kk: int) -> tuple[torch.Tensor, torch.Tensor]:
        return xfmr_b200.topk_search(qs, items, kk, id_base=rank * shard)

    def step() -> None:
        if world > 1:
            xfmr_b200.distributed.sharded_topk(search, xfmr_b200.topk_merge, queries, k)
        else:
            search(queries, k)

    step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    e1.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    # the same search with 64 excluded ids per query: the 64 best items of every query are excluded, through
    # the sparse path (rank k + 64 per shard, drop the listed ids, keep k) - a dense Q x N mask would be 819 GB here
    n_excl = 64

    def search_excl(qs: torch.Tensor, kk: int) -> tuple[torch.Tensor, torch.Tensor]:
        s_, i_ = xfmr_b200.topk_search(qs, items, kk + n_excl, id_base=rank * shard)
        return xfmr_b200.topk_filter(s_, i_, excl, kk)

    def step_excl() -> tuple[torch.Tensor, torch.Tensor]:
        if world > 1:
            return xfmr_b200.distributed.sharded_topk(search_excl, xfmr_b200.topk_merge, queries, k)
        return search_excl(queries, k)

    if world > 1:
        _, excl = xfmr_b200.distributed.sharded_topk(search, xfmr_b200.topk_merge, queries, n_excl)
    else:
        _, excl = search(queries, n_excl)
    excl = excl.contiguous()
    step_excl()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, ids_excl = step_excl()
    e1.record()
    e1.synchronize()
    ms_excl = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms_excl, op=dist.ReduceOp.MAX)
    ms_excl = float(ms_excl)
    sample = slice(0, 512)
    leaked = bool((ids_excl[sample, :, None] == excl[sample, None, :]).any())
    assert not leaked, "an excluded id came back"
    flops = 2.0 * num_queries * shard * 128
    pk = peaks()
    del items
    return {"metric": "exact_top100_queries_per_s", "value": num_queries / (ms * 1e-3), "unit": "queries/s",
            "with_64_exclusions_per_query": {"value": num_queries / (ms_excl * 1e-3), "unit": "queries/s", "ms": ms_excl},
            "workload": f"{num_queries} queries x {num_items_total} items (sharded {world} ways), d=128 bf16, k={k}",
            "ms": ms, "scaling": "strong", "tflops_per_gpu": flops / (ms * 1e-3) / 1e12,
            "tensor_frac_of_sustained_peak": flops / (ms * 1e-3) / (pk["bf16_tflops_sustained"] * 1e12),
            "tensor_frac_of_burst_peak": flops / (ms * 1e-3) / (pk["bf16_tflops"] * 1e12)}


def bench_mns(device: torch.device, world: int, rank: int) -> dict:
    """Config 3: mixed negative sampling with global negatives - per rank 8,192 users, their 8,192 in-batch items and
    16,384 uniform negatives, d=256 bf16; items and negatives are all-gathered over NCCL (every rank scores its users
    against 24,576 x world candidates) and the item gradients are reduced back to their owners in the backward pass."""
    import torch.distributed as dist  # noqa: PLC0415

    import xfmr_b200  # noqa: PLC0415
    from xfmr_b200 import synthetic  # noqa: PLC0415

    b, u, d, p = 8192, 16384, 256, 32
    inp = synthetic.make_loss_inputs(b, b + u, d, p, n_catalog=200_000, seed=50 + rank)
    q = inp["user_embed"].to(device, torch.bfloat16)
    items = inp["item_embed"][:b].to(device, torch.bfloat16)
    negs = inp["item_embed"][b:].to(device, torch.bfloat16)
    target, pos_idx = inp["target"].to(device), inp["pos_idx"].to(device)
    item_idx, neg_idx = inp["item_idx"][:b].to(device), inp["item_idx"][b:].to(device)
    module = getattr(xfmr_b200, HEADLINE_LOSS)(sigma=SIGMA, margin=MARGIN)

    qq, ii, nn = q.requires_grad_(True), items.requires_grad_(True), negs.requires_grad_(True)
    idx_cat = torch.cat([item_idx, neg_idx])

    def step() -> tuple:
        if world > 1:
            loss = xfmr_b200.distributed.global_negatives_losses(module, qq, ii, nn, target, item_idx=item_idx,
                                                                 neg_idx=neg_idx, pos_idx=pos_idx)
        else:
            loss = module(qq, torch.cat([ii, nn]), target, item_idx=idx_cat, pos_idx=pos_idx)
        return (loss, *torch.autograd.grad(loss, (qq, ii, nn)))

    launch = "eager (NCCL collectives inside the step)"
    if world == 1:
        step = graphed(step)      # one rank: no collective in the step, replay it like the headline
        launch = "CUDA-graph replay"
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    e1.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    flops = 6.0 * b * (b + u) * world * d   # per rank: its users against the global candidate set
    pk = peaks()
    return {"metric": "mns_global_negatives_samples_per_s", "value": world * b / (ms * 1e-3), "unit": "samples/s",
            "workload": f"C3: per rank {b} users x ({b} in-batch + {u} uniform) x {world} ranks candidates, d={d} bf16, "
                        f"sampled-softmax fwd+bwd, NCCL all-gather of items/negatives + gradient reduction",
            "ms_per_step": ms, "launch": launch, "tflops_per_gpu": flops / (ms * 1e-3) / 1e12,
            "tensor_frac_of_sustained_peak": flops / (ms * 1e-3) / (pk["bf16_tflops_sustained"] * 1e12)}


def bench_evaluate(device: torch.device) -> dict:
    """Batched validation at MovieLens-1M shape: every user at once — exact top-20 with the user's
    history excluded + the six ranking metrics — where the reference runs one user per step (lightning.py:149-206)."""
    import xfmr_b200  # noqa: PLC0415
    from xfmr_b200 import synthetic  # noqa: PLC0415

    users, items, dim, k, hist, tgt = 6040, 3706, 64, 20, 165, 20
    gen = torch.Generator(device=device).manual_seed(5)
    catalog = synthetic.make_catalog(items, dim, seed=3, device=device)
    queries = synthetic.make_catalog(users, dim, seed=4, device=device)
    history = torch.randint(1, items + 1, (users, hist), generator=gen, device=device)
    target_ids = torch.randint(1, items + 1, (users, tgt), generator=gen, device=device)
    target_vals = torch.randint(1, 6, (users, tgt), generator=gen, device=device).float()
    index = xfmr_b200.ItemProcessor().get_index(catalog, torch.arange(1, items + 1, device=device))

    def step() -> None:
        index.evaluate(queries, target_ids, target_vals, history, top_k=k)

    t = timed_steps(step, 5, 3, None)
    ms = statistics.median(t)
    return {"metric": "batched_validation_users_per_s", "value": users / (ms * 1e-3), "unit": "users/s", "ms": ms,
            "workload": f"C1-shaped: {users} users x {items} items, d={dim} fp32 (exact ids), top-{k}, {hist} excluded history "
                        f"ids per user (dense mask), {tgt} graded targets per user, 6 metrics"}


def bench_gather(device: torch.device) -> dict:
    import xfmr_b200  # noqa: PLC0415

    n, d, log2 = 1 << 22, 128, 22
    gen = torch.Generator(device=device).manual_seed(3)
    table = (torch.randn(1 << log2, d, device=device, generator=gen) * 0.02).to(torch.bfloat16)
    ids = torch.randint(0, 2**62, (n,), device=device, generator=gen)
    fn = lambda: xfmr_b200.hash_embedding_gather(table, ids, 2)  # noqa: E731
    times = timed_steps(fn, 10, 3, None)
    ms = statistics.median(times)
    algo_bytes = n * (8 + 2 * d * 2 + d * 2)
    pk = peaks()
    return {"metric": "hash_gather_GBps", "value": algo_bytes / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms": ms,
            "workload": "C4: 4Mi ids, k=2 hashes, 2^22 x 128 bf16 table", "frac_of_hbm_peak": algo_bytes / (ms * 1e-3) / 1e9 / pk["hbm_gbs"]}


def main() -> None:  # noqa: PLR0915
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="time eager module calls instead of CUDA-graph replay")
    ap.add_argument("--no-extras", action="store_true", help="headline line only (skip per-loss / retrieval / gather / cpu baseline)")
    ap.add_argument("--retrieval-items", type=int, default=100_000_000,
                    help="catalog rows of the retrieval extra, summed over all ranks (config 5: 100,000,000 = 25.6 GB bf16)")
    ap.add_argument("--retrieval-queries", type=int, default=65_536)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist  # noqa: PLC0415

    import xfmr_b200  # noqa: PLC0415
    from xfmr_b200 import _lib  # noqa: PLC0415

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries the one JSON line and nothing else: native libraries write there too (NCCL prints its version
    # banner with printf at NCCL_DEBUG=VERSION and WARN), so file descriptor 1 points at stderr until the line is ready
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    warmup = max(args.warmup, 3)
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    inp = make_c2(device, rank, torch.bfloat16)
    module = getattr(xfmr_b200, HEADLINE_LOSS)(sigma=SIGMA, margin=MARGIN)
    eager_step = loss_step_fn(module, inp)
    flush_buf = torch.zeros(64 << 20, dtype=torch.float32, device=device)
    eager_ms = statistics.mean(timed_steps(eager_step, min(args.steps, 50), warmup, flush_buf))
    step = eager_step if args.no_graph else graphed(eager_step)

    # ---- device-resident timing (value) + live sweep-kernel timing (roofline)
    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(warmup):
        flush_l2(flush_buf)
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks.mark_begin()
    times = timed_steps(step, args.steps, 0, flush_buf)
    torch.cuda.synchronize()
    clocks.mark_end()
    clocks.stop()
    # live duration of the dominant kernel (the three sweep launches of a step): CUDA events recorded by the
    # library on the launch stream, over the same number of eager steps (events cannot be read from a replay)
    _lib.launch_count(reset=True)
    _lib.sweep_timing(True)
    timed_steps(eager_step, args.steps, 0, flush_buf)
    torch.cuda.synchronize()
    sweep_ms_total, sweep_count = _lib.sweep_timing_read()
    _lib.sweep_timing(False)
    launches = _lib.launch_count()
    total_ms = torch.tensor([sum(times)], device=device, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)
    ms_per_step = total_ms / args.steps
    value = world * C2["batch"] / (ms_per_step * 1e-3)

    # ---- end-to-end through the public API with host (pinned) inputs
    host = {k: v.cpu().pin_memory() for k, v in inp.items() if k != "log_q"}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    # Input pipeline of the timed loop: a copy stream uploads the inputs of step i+1 (pinned host -> device) while the
    # compute stream runs step i; every step still pays for its own upload and its own loss read-back.
    copy_stream = torch.cuda.Stream(device=device)

    def upload() -> tuple[dict, torch.cuda.Event]:
        with torch.cuda.stream(copy_stream):
            dev = {k: v.to(device, non_blocking=True) for k, v in host.items()}
            done = torch.cuda.Event()
            done.record(copy_stream)
        return dev, done

    def e2e_run(steps: int) -> float:
        last = 0.0
        nxt = upload()
        for i in range(steps):
            dev, done = nxt
            if i + 1 < steps:
                nxt = upload()
            cur = torch.cuda.current_stream(device)
            cur.wait_event(done)
            for t in dev.values():
                t.record_stream(cur)
            q = dev["user_embed"].requires_grad_(True)
            v = dev["item_embed"].requires_grad_(True)
            loss = module(q, v, dev["target"], item_idx=dev["item_idx"], pos_idx=dev["pos_idx"])
            loss.backward()
            last = float(loss.detach())  # device -> host read of the step's result (synchronises the compute stream)
        return last

    # Graph replay with a pipelined feed (xfmr_b200.GraphedLossStep): the upload of step i+1 runs on a copy stream under
    # the replay of step i, and the host reads the loss of step i-1 while step i is in flight.  Every step still pays
    # for its own upload from pinned memory and its own loss read-back, all inside the timed region.
    stepper = None if args.no_graph else xfmr_b200.GraphedLossStep(module, inp)

    def e2e_run_graph(steps: int) -> float:
        last = 0.0
        prev = None
        stepper.prefetch(host)
        for i in range(steps):
            res = stepper.submit()
            if i + 1 < steps:
                stepper.prefetch(host)
            if prev is not None:
                last = prev.loss_value()
            prev = res
        if prev is not None:
            last = prev.loss_value()
        return last

    e2e_eager_value = None
    if stepper is not None:
        eager_loss = e2e_run(3)
        graph_loss = e2e_run_graph(3)
        assert abs(eager_loss - graph_loss) <= 1e-5 * abs(eager_loss), (eager_loss, graph_loss)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_run(20)   # the same loop with eager module calls and a blocking loss read per step, reported beside it
        torch.cuda.synchronize()
        e2e_eager_value = world * C2["batch"] * 20 / (time.perf_counter() - t0)
        e2e_run = e2e_run_graph
    # what the host -> device path of this box delivers for exactly these buffers (explains e2e: it is copy-bound)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        upload()
    torch.cuda.synchronize()
    h2d_gbps = 10 * h2d_bytes / (time.perf_counter() - t0) / 1e9
    e2e_run(3)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_steps = max(5, min(args.steps, 50))
    e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * C2["batch"] * e2e_steps / float(e2e_s)

    pk = peaks()
    algo_flops = 6.0 * C2["batch"] * C2["num_items"] * C2["dim"]  # fwd 2BNd + dQ 2BNd + dI 2BNd
    sweep_ms_per_step = sweep_ms_total / args.steps
    achieved = algo_flops / (sweep_ms_per_step * 1e-3) / 1e12
    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16",
        "data": "synthetic",
        "config": {
            "workload": "C2 MovieLens-32M-shaped: batch 4096 x 87585 items, d=128, P=32, bf16, sampled-softmax "
                        "(InfomationNoiseContrastiveEstimationLoss) fwd+bwd through the drop-in module",
            "sigma": SIGMA, "margin": MARGIN, "num_negatives": 0,
            "l2": "flushed between steps (256 MiB read+write)",
            "launch": "eager" if args.no_graph else "CUDA-graph replay of the module's forward+backward (eager ms in eager_ms_per_step)",
            "parallelism": f"dp{world} (users sharded, no data-path collective)",
        },
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "eager_blocking_value": e2e_eager_value, "h2d_only_GBps": h2d_gbps,
                "pipeline": ("eager module calls; the pinned-host upload of step i+1 runs on a copy stream under step i" if args.no_graph else
                             "xfmr_b200.GraphedLossStep: CUDA-graph replay (one graph per input slot, no staging copy); the pinned-host "
                             "upload of step i+1 runs on a copy stream under step i and the loss of step i-1 is read on the "
                             "host while step i runs")},
        "gpu_launches": launches,
        "eager_ms_per_step": eager_ms,
        "roofline": {
            "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": achieved / pk["bf16_tflops_sustained"], "traffic": ncu_traffic(),
            "kernel": "xb::sweep_kernel (2 working launches per step: merged forward + dQ sweep, dI sweep; plus 2 conditional "
                      "fallback launches that exit at once)",
            "algorithmic_flops_per_step": algo_flops, "sweep_ms_per_step": sweep_ms_per_step,
            "sweep_launches_per_step": sweep_count / args.steps, "peak_source": pk["source"] + " sustained bf16",
            "sweep_share_of_step": sweep_ms_per_step / ms_per_step,
        },
    }

    if not args.no_extras:
        per_loss = {}
        for name in xfmr_b200.LOSS_SLOTS:
            m = getattr(xfmr_b200, name)(sigma=SIGMA, margin=MARGIN)
            t = timed_steps(graphed(loss_step_fn(m, inp)), 8, 3, flush_buf)
            per_loss[name] = {"ms_per_step": statistics.median(t), "samples_per_s": C2["batch"] / (statistics.median(t) * 1e-3)}
        mined = xfmr_b200.PairwiseHingeLoss(num_negatives=4, sigma=SIGMA, margin=MARGIN)  # the reference's training default
        t = timed_steps(graphed(loss_step_fn(mined, inp)), 8, 3, flush_buf)
        per_loss["PairwiseHingeLoss[num_negatives=4]"] = {"ms_per_step": statistics.median(t),
                                                           "samples_per_s": C2["batch"] / (statistics.median(t) * 1e-3)}

        hard = xfmr_b200.PairwiseHingeLoss(num_negatives=4, sigma=SIGMA, margin=MARGIN, mining="hard")
        t = timed_steps(graphed(loss_step_fn(hard, inp)), 8, 3, flush_buf)
        per_loss["PairwiseHingeLoss[num_negatives=4, mining=hard]"] = {"ms_per_step": statistics.median(t),
                                                                        "samples_per_s": C2["batch"] / (statistics.median(t) * 1e-3)}
        directau = xfmr_b200.DirectAULoss(gamma=1.0, t=2.0)
        t = timed_steps(graphed(loss_step_fn(directau, inp)), 8, 3, flush_buf)
        per_loss["DirectAULoss"] = {"ms_per_step": statistics.median(t), "samples_per_s": C2["batch"] / (statistics.median(t) * 1e-3)}
        for n_rows in (C2["batch"], C2["num_items"]):
            x = inp["item_embed"][:n_rows].detach().requires_grad_(True)

            def uni_step(x=x) -> tuple:  # noqa: ANN001
                loss = xfmr_b200.uniformity_loss(x, 2.0)
                return loss, torch.autograd.grad(loss, x)[0]

            t = timed_steps(graphed(uni_step), 8, 3, flush_buf)
            ms = statistics.median(t)
            # one sweep of 2 tile-MMAs per tile pair does forward and backward: 4 n^2 d algorithmic flops
            per_loss[f"uniformity_loss[n={n_rows}]"] = {"ms_per_step": ms, "tflops": 4.0 * n_rows * n_rows * C2["dim"] / (ms * 1e-3) / 1e12}

        def fused_fwd() -> None:
            xfmr_b200.fused_losses(inp["user_embed"], inp["item_embed"], inp["target"], item_idx=inp["item_idx"],
                                   pos_idx=inp["pos_idx"], sigma=SIGMA, margin=MARGIN)

        t = timed_steps(fused_fwd, 8, 3, flush_buf)
        per_loss["all_seven_forward_one_call"] = {"ms_per_step": statistics.median(t)}
        line["per_loss"] = per_loss
        del inp, host
        torch.cuda.empty_cache()
        line["mns"] = bench_mns(device, world, rank)
        torch.cuda.empty_cache()
        line["retrieval"] = bench_retrieval(device, world, rank, args.retrieval_items, args.retrieval_queries, 100)
        if rank == 0:
            line["evaluate"] = bench_evaluate(device)
            line["gather"] = bench_gather(device)
            base = cpu_reference_rate(2, 1)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["aten_gpu_baseline"] = aten_gpu_reference_rate(device)
    elif rank == 0:
        base = cpu_reference_rate(1, 1)
        line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
