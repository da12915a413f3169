"""Benchmark of the B200 score path.  ``python bench.py --gpus N --steps K --warmup W [--impl reference]``.

Headline (BASELINE.json: "top-100 retrieval queries/s at 1/2/4/8 B200; fused loss fwd+bwd samples/s"):
**exact top-100 retrieval queries/s on config 5** - 65,536 queries x 100,000,000 items, d=128 bf16 - with the
catalog row-sharded over the ranks and a query-sharded NCCL merge of the per-shard lists (`xfmr_b200.distributed.
RetrievalGrid` / `sharded_topk`; `--retrieval-shards R` selects R shards x N / R query groups instead): STRONG scaling, the one multi-GPU path of this repository that exchanges data.  A "step" is one search
of all 65,536 queries.  `value` times the steps with the queries resident on the device; `e2e` takes the queries from
pinned host memory and brings scores + ids back to the host inside the timed region, through the same public call.

The second half of BASELINE.json's metric rides in the same JSON line under `"loss"`: config 2 (batch 4,096 x 87,585
items, d=128, P=32, bf16), `InfomationNoiseContrastiveEstimationLoss` forward + backward through the drop-in module, with
its own `roofline` (sweep kernels against the measured BURST bf16 peak: the timed region is a 10 ms burst), `e2e` and
per-loss table; further extras: config 3 (`mns`), config 1 (`c1`, microseconds), config 4 (`gather`), batched validation.

`--impl reference` times the reference's side on the host cores of rank 0 (never the product: it does not import the
package, so `libxfmr_b200.so` is not mapped): for the headline the exact brute-force restatement of `ItemProcessor.search`
(the reference's own search is LanceDB's approximate index, absent here) on a bounded sample of config 5 per step, and
under `"loss"` the UNMODIFIED reference loss class from `oracle/_ref` (copied by `make -C oracle _ref` at build time) at
the full config 2.
"""

from __future__ import annotations

import argparse
import importlib.util
import json
import os
import pathlib
import statistics
import sys
import time

import torch

ROOT = pathlib.Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "exact_top100_retrieval_queries_per_s"
UNIT = "queries/s"
C5 = {"num_queries": 65_536, "num_items": 100_000_000, "dim": 128, "k": 100}
LOSS_METRIC = "fused_loss_fwd_bwd_samples_per_s"
C2 = {"batch": 4096, "num_items": 87585, "dim": 128, "num_pos": 32}
C1 = {"batch": 1024, "num_items": 3706, "dim": 64, "num_pos": 32, "k": 10}
HEADLINE_LOSS = "InfomationNoiseContrastiveEstimationLoss"
SIGMA, MARGIN = 5.0, 0.5


def workload_config(num_items: int, num_queries: int) -> dict:
    """The `config` object: identical in both arms (the driver compares them)."""
    return {
        "workload": f"C5 exact top-{C5['k']} retrieval: {num_queries} queries x {num_items} items, d={C5['dim']} bf16, "
                    "catalog row-sharded over the ranks, NCCL merge of the per-shard lists",
        "l2": "inputs larger than L2 (every step streams the whole catalog shard, 25.6 GB in total)",
    }


def load_synthetic():  # noqa: ANN201
    """`synthetic.py` straight from its file: the reference arm must not import the package (that would map the product)."""
    spec = importlib.util.spec_from_file_location("xb_synthetic", ROOT / "matrix-factorization-torch_b200" / "synthetic.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def peaks() -> dict:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        data = json.loads(path.read_text())
        return {"hbm_gbs": data["hbm_gbs"], "bf16_tflops": data["bf16_tflops"],
                "bf16_tflops_sustained": data.get("bf16_tflops_sustained", data["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(key: str) -> float | None:
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the named kernel, from the committed `ncu` capture
    (profiles/ncu_traffic.json); None if it has not been captured."""
    path = ROOT / "profiles" / "ncu_traffic.json"
    if not path.exists():
        return None
    val = json.loads(path.read_text()).get(key)
    return None if val is None else float(val)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled every few milliseconds from
    a thread between ``mark_begin`` and ``mark_end``, plus one sample taken by the caller while the GPU is still busy."""

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
    """Per-step device milliseconds (CUDA events on the current stream); L2 flushed between steps if a buffer is given."""
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
    synthetic = load_synthetic()
    inp = synthetic.make_loss_inputs(C2["batch"], C2["num_items"], C2["dim"], C2["num_pos"], n_catalog=C2["num_items"],
                                     seed=seed, device=device)
    inp["user_embed"] = inp["user_embed"].to(dtype)
    inp["item_embed"] = inp["item_embed"].to(dtype)
    return inp


def loss_step_fn(module, inp: dict):  # noqa: ANN001, ANN201
    q = inp["user_embed"].detach().requires_grad_(True)
    v = inp["item_embed"].detach().requires_grad_(True)
    extra = {"log_q": inp["log_q_arg"]} if "log_q_arg" in inp else {}

    def step() -> tuple:
        loss = module(q, v, inp["target"], item_idx=inp["item_idx"], pos_idx=inp["pos_idx"], **extra)
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


# ====================================================================================================== reference arm
def cpu_topk_sample(num_queries: int, num_items: int, steps: int, warmup: int) -> dict:
    """The exact brute-force restatement of ``ItemProcessor.search`` (xfmr_rec/data/lightning.py:237-259: cosine scores of
    unit-norm embeddings = inner products, top-k by score) on the host cores: ``(Q_s @ I_s^T).topk(k)`` in
    fp32 with every thread torch has.  One step = ``num_queries`` queries against ``num_items`` items; MEASURED times only."""
    synthetic = load_synthetic()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    items = synthetic.make_catalog(num_items, C5["dim"], seed=100)
    queries = synthetic.make_catalog(num_queries, C5["dim"], seed=7)

    def run() -> float:
        t0 = time.perf_counter()
        scores = queries @ items.t()
        scores.topk(C5["k"], dim=1)
        return time.perf_counter() - t0

    for _ in range(warmup):
        run()
    times = [run() for _ in range(steps)]
    return {"cores": cores, "times_s": times, "num_queries": num_queries, "num_items": num_items}


def reference_loss_module():  # noqa: ANN201
    """(module class, kind): the unmodified reference class from oracle/_ref when the recipe has been run, else the port."""
    ref_dir = ROOT / "oracle" / "_ref"
    if (ref_dir / "xfmr_rec" / "losses.py").exists():
        sys.path.insert(0, str(ref_dir))
        import xfmr_rec.losses as ref  # noqa: PLC0415

        return getattr(ref, HEADLINE_LOSS), "reference"
    return None, "port"


def cpu_loss_full_c2(steps: int, warmup: int) -> dict:
    """The reference's loss (fwd + ``autograd.grad``) at the FULL config 2 on the host cores.  The ``B x N x P`` accidental-hit
    broadcast of losses.py:108 is 11.5 GB of booleans at P=32, which the host has room for."""
    synthetic = load_synthetic()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inp = synthetic.make_loss_inputs(C2["batch"], C2["num_items"], C2["dim"], C2["num_pos"], n_catalog=C2["num_items"], seed=0)
    cls, kind = reference_loss_module()
    if cls is not None:
        module = cls(sigma=SIGMA, margin=MARGIN)

        def run() -> float:
            q = inp["user_embed"].clone().requires_grad_(True)
            v = inp["item_embed"].clone().requires_grad_(True)
            t0 = time.perf_counter()
            loss = module(q, v, inp["target"], item_idx=inp["item_idx"], pos_idx=inp["pos_idx"])
            torch.autograd.grad(loss, (q, v))
            return time.perf_counter() - t0
    else:
        from oracle import losses_oracle  # noqa: PLC0415

        def run() -> float:
            t0 = time.perf_counter()
            losses_oracle.losses_and_grads(inp["user_embed"], inp["item_embed"], inp["target"], item_idx=inp["item_idx"],
                                           pos_idx=inp["pos_idx"], sigma=SIGMA, margin=MARGIN, names=(HEADLINE_LOSS,))
            return time.perf_counter() - t0

    for _ in range(warmup):
        run()
    times = [run() for _ in range(steps)]
    return {"cores": cores, "kind": kind, "times_s": times}


def run_reference(args: argparse.Namespace) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    # One step = a bounded sample of config 5: all host threads on 512 queries x 1,000,000 items (1.3e11 flop + top-100)
    q_s, n_s = 512, 1_000_000
    res = cpu_topk_sample(q_s, n_s, steps, warmup)
    ms = statistics.mean(res["times_s"]) * 1e3
    # queries/s at the full catalog: the brute-force cost is linear in the number of items
    value = q_s / (ms * 1e-3) * (n_s / args.retrieval_items)
    sample = (f"exact brute-force top-{C5['k']} (torch fp32 matmul + topk, {res['cores']} threads) on {q_s} queries x {n_s} items per step: "
              f"{ms:.1f} ms measured; value = {q_s} / t x ({n_s} / {args.retrieval_items} items), i.e. the full-catalog rate")
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": steps,
        "warmup": warmup,
        "ms_per_step": ms,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(args.retrieval_items, args.retrieval_queries),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference's own search is LanceDB IVF_HNSW_PQ (approximate, third-party, not installable here); this arm "
                "times the exact restatement of its semantics; ms_per_step is the measured time of one sample step",
    }
    if not args.no_extras:
        loss_steps = min(steps, 3)
        lres = cpu_loss_full_c2(loss_steps, min(warmup, 1))
        lms = statistics.mean(lres["times_s"]) * 1e3
        line["loss"] = {
            "metric": LOSS_METRIC, "value": C2["batch"] / (lms * 1e-3), "unit": "samples/s", "ms_per_step": lms,
            "steps": loss_steps, "kind": lres["kind"], "cores": lres["cores"],
            "config": "C2 at full size: 4096 x 87585, d=128, P=32, fp32, " + HEADLINE_LOSS + " fwd + autograd.grad "
                      + ("(unmodified xfmr_rec/losses.py from oracle/_ref)" if lres["kind"] == "reference" else "(oracle port)"),
        }
    print(json.dumps(line))  # noqa: T201


# ====================================================================================================== retrieval (headline)
def bench_retrieval(device: torch.device, world: int, rank: int, args: argparse.Namespace, clocks: ClockSampler) -> dict:
    """Config 5, strong scaling: rank r owns items [r * N/G, (r+1) * N/G), every rank scores all queries against its
    shard, `sharded_topk` merges (query-sharded all-to-all + merge + all-gather)."""
    import torch.distributed as dist  # noqa: PLC0415

    import xfmr_b200  # noqa: PLC0415
    from xfmr_b200 import _lib  # noqa: PLC0415

    synthetic = load_synthetic()
    num_items_total, num_queries, k, d = args.retrieval_items, args.retrieval_queries, C5["k"], C5["dim"]
    # layout: R catalog shards x (world / R) query groups; rank r owns shard r % R and query group r // R.  Default R = world:
    # plain row sharding, the fastest layout measured at config 5 (profiles/r02_layouts_8gpu.txt)
    n_shards = args.retrieval_shards or world
    grid = xfmr_b200.distributed.RetrievalGrid(n_shards) if world > 1 else None
    shard_idx = grid.shard if grid else 0
    shard = num_items_total // n_shards
    items = synthetic.make_catalog(shard, d, seed=100 + shard_idx, device=device, dtype=torch.bfloat16)
    queries = synthetic.make_catalog(num_queries, d, seed=7, device=device, dtype=torch.bfloat16)
    queries_host = queries.cpu().pin_memory()
    queries_per_rank_group = num_queries // (grid.query_groups if grid else 1)

    def search(qs: torch.Tensor, kk: int) -> tuple[torch.Tensor, torch.Tensor]:
        return xfmr_b200.topk_search(qs, items, kk, id_base=shard_idx * shard)

    def step(qs: torch.Tensor = queries) -> tuple[torch.Tensor, torch.Tensor]:
        return step_n(k, qs)

    def step_n(kk: int, qs: torch.Tensor = queries) -> tuple[torch.Tensor, torch.Tensor]:
        if grid:
            return grid.search(search, xfmr_b200.topk_merge, qs, kk)
        return search(qs, kk)

    def barrier() -> None:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    steps, warmup = args.steps, max(args.warmup, 3)
    for _ in range(warmup):
        step()
    barrier()
    # ---- device-resident steps (value): CUDA events per step on the launch stream, max over ranks of the total
    _lib.launch_count(reset=True)
    _lib.sweep_timing(True)
    clocks.mark_begin()
    times = timed_steps(step, steps, 0, None)
    clocks.mark_end()
    torch.cuda.synchronize()
    sweep_ms_total, sweep_count = _lib.sweep_timing_read()
    _lib.sweep_timing(False)
    launches = _lib.launch_count()
    total = torch.tensor([sum(times)], device=device, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    ms_per_step = float(total) / steps
    # ---- end to end: queries from pinned host memory, scores + ids back to the host, every step
    out_scores = torch.empty(num_queries, k, dtype=torch.float32).pin_memory()
    out_ids = torch.empty(num_queries, k, dtype=torch.int64).pin_memory()

    def e2e_step() -> None:
        qd = queries_host.to(device, non_blocking=True)
        s_, i_ = step(qd)
        out_scores.copy_(s_, non_blocking=True)
        out_ids.copy_(i_, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    e2e_step()
    barrier()
    e2e_steps = max(3, min(steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_s) / e2e_steps * 1e3
    # ---- recall of the bf16 search at full scale: 64 random queries, brute force over every shard on the same values
    scores, ids = step()
    gen = torch.Generator().manual_seed(11)
    sample = torch.randperm(num_queries, generator=gen)[:64].to(device)
    qf = queries[sample].float()
    best_s = torch.full((64, k), float("-inf"), device=device)
    best_i = torch.full((64, k), -1, dtype=torch.int64, device=device)
    chunk = 4_000_000
    for lo in range(0, shard, chunk):
        sc = qf @ items[lo:lo + chunk].float().t()
        top_s, top_i = sc.topk(min(k, sc.size(1)), dim=1)
        cat_s = torch.cat([best_s, top_s], dim=1)
        cat_i = torch.cat([best_i, top_i + (shard_idx * shard + lo)], dim=1)
        best_s, sel = cat_s.topk(k, dim=1)
        best_i = cat_i.gather(1, sel)
        del sc
    if grid:   # one copy of every shard: the ranks of this rank's catalog group
        all_s = [torch.empty_like(best_s) for _ in range(grid.shards)]
        all_i = [torch.empty_like(best_i) for _ in range(grid.shards)]
        dist.all_gather(all_s, best_s, group=grid.catalog_group)
        dist.all_gather(all_i, best_i, group=grid.catalog_group)
        cat_s, cat_i = torch.cat(all_s, dim=1), torch.cat(all_i, dim=1)
        best_s, sel = cat_s.topk(k, dim=1)
        best_i = cat_i.gather(1, sel)
    got = ids[sample]
    kth = best_s[:, -1:]
    must = best_s > kth                       # strictly above the k-th exact score: must be in the returned list
    hit = (best_i.unsqueeze(2) == got.unsqueeze(1)).any(dim=2)
    recall = float((hit & must).sum()) / max(float(must.sum()), 1.0)
    assert recall >= 0.999, f"recall@{k} = {recall} at {num_items_total} items"  # noqa: S101
    # ---- the same search with 64 excluded ids per query through the sparse path (a dense Q x N mask would be 819 GB)
    n_excl = 64

    def search_excl(qs: torch.Tensor, kk: int) -> tuple[torch.Tensor, torch.Tensor]:
        s_, i_ = xfmr_b200.topk_search(qs, items, kk + n_excl, id_base=shard_idx * shard)
        return xfmr_b200.topk_filter(s_, i_, excl_mine, kk)

    def step_excl() -> tuple[torch.Tensor, torch.Tensor]:
        if grid:
            return grid.search(search_excl, xfmr_b200.topk_merge, queries, k)
        return search_excl(queries, k)

    _, excl = step_n(n_excl)
    excl = excl.contiguous()
    excl_mine = excl
    if grid and grid.query_groups > 1:   # the exclusion lists of the queries this rank's group searches (padded like them)
        excl_mine = excl[grid.query_slice(num_queries)]
        short = grid.rows_per_group(num_queries) - excl_mine.size(0)
        if short:
            excl_mine = torch.cat([excl_mine, excl_mine.new_full((short, n_excl), -1)])
        excl_mine = excl_mine.contiguous()
    step_excl()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, ids_excl = step_excl()
    e1.record()
    e1.synchronize()
    ms_excl = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms_excl, op=dist.ReduceOp.MAX)
    ms_excl = float(ms_excl)
    leaked = bool((ids_excl[:512, :, None] == excl[:512, None, :]).any())
    assert not leaked, "an excluded id came back"  # noqa: S101
    del items
    torch.cuda.empty_cache()
    pk = peaks()
    flops_per_rank = 2.0 * queries_per_rank_group * shard * d
    sweep_ms = sweep_ms_total / max(sweep_count, 1)
    achieved = flops_per_rank / (sweep_ms * 1e-3) / 1e12
    return {
        "value": num_queries / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "gpu_launches": launches,
        "e2e": {"value": num_queries / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": world * queries_host.numel() * 2,
                "d2h_bytes_per_step": out_scores.numel() * 4 + out_ids.numel() * 8, "ms_per_step": e2e_ms, "steps": e2e_steps,
                "pipeline": "every rank uploads the queries from pinned memory, searches its shard, merges; scores + ids are "
                            "read back into pinned memory; host waits for the step before starting the next"},
        "roofline": {
            "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": achieved / pk["bf16_tflops_sustained"], "traffic": (None if ncu_traffic("topk_sweep_dram_bytes_per_item") is None
                        else ncu_traffic("topk_sweep_dram_bytes_per_item") * shard),
            "traffic_note": "ncu capture on a 12.5M-item shard scaled by the items per rank; the catalog does not fit the 126 MB "
                            "L2, each 64 MB chunk is re-read once per wave of query-tile pairs (DESIGN.md 3.5)",
            "kernel": "xb::rt_kernel (one launch per search and rank)",
            "algorithmic_flops_per_launch": flops_per_rank, "launch_ms": sweep_ms, "launches_timed": sweep_count,
            "peak_source": pk["source"] + " sustained bf16 (the launch lasts seconds under the power cap)",
            "frac_of_burst_peak": achieved / pk["bf16_tflops"],
            "sweep_share_of_step": sweep_ms / ms_per_step,
        },
        "recall_at_k_sampled": {"value": recall, "queries": 64, "against": "fp32 brute force over all shards on the same bf16 values"},
        "with_64_exclusions_per_query": {"value": num_queries / (ms_excl * 1e-3), "unit": UNIT, "ms": ms_excl},
        "layout": {"catalog_shards": n_shards, "query_groups": grid.query_groups if grid else 1, "items_per_rank": shard,
                   "queries_per_rank": queries_per_rank_group},
    }


# ====================================================================================================== loss (config 2)
def bench_loss(device: torch.device, world: int, rank: int, args: argparse.Namespace) -> dict:
    import xfmr_b200  # noqa: PLC0415
    from xfmr_b200 import _lib  # noqa: PLC0415

    steps, warmup = args.steps, max(args.warmup, 3)
    inp = make_c2(device, rank, torch.bfloat16)
    module = getattr(xfmr_b200, HEADLINE_LOSS)(sigma=SIGMA, margin=MARGIN)
    eager_step = loss_step_fn(module, inp)
    flush_buf = torch.zeros(64 << 20, dtype=torch.float32, device=device)
    eager_ms = statistics.mean(timed_steps(eager_step, min(steps, 50), warmup, flush_buf))
    step = graphed(eager_step)
    clocks = ClockSampler(device.index or 0)
    clocks.start()
    for _ in range(warmup):
        flush_l2(flush_buf)
        step()
    torch.cuda.synchronize()
    clocks.mark_begin()
    times = timed_steps(step, steps, 0, flush_buf)
    torch.cuda.synchronize()
    clocks.mark_end()
    clocks.stop()
    # live duration of the dominant kernels (the sweep launches of a step): CUDA events recorded by the library on the
    # launch stream, over the same number of eager steps (events cannot be read from a graph replay)
    _lib.launch_count(reset=True)
    _lib.sweep_timing(True)
    timed_steps(eager_step, steps, 0, flush_buf)
    torch.cuda.synchronize()
    sweep_ms_total, sweep_count = _lib.sweep_timing_read()
    _lib.sweep_timing(False)
    launches = _lib.launch_count()
    ms_per_step = sum(times) / steps
    # ---- end to end through GraphedLossStep (pinned host inputs every step, loss read back every step)
    host = {k: v.cpu().pin_memory() for k, v in inp.items() if k != "log_q"}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
    stepper = xfmr_b200.GraphedLossStep(module, inp)

    def e2e_run(n: int) -> float:
        last, prev = 0.0, None
        stepper.prefetch(host)
        for i in range(n):
            res = stepper.submit()
            if i + 1 < n:
                stepper.prefetch(host)
            if prev is not None:
                last = prev.loss_value()
            prev = res
        if prev is not None:
            last = prev.loss_value()
        return last

    e2e_run(3)
    torch.cuda.synchronize()
    e2e_steps = max(5, min(steps, 50))
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize()
    e2e_value = C2["batch"] * e2e_steps / (time.perf_counter() - t0)
    # "ids / targets from the host, embeddings already on the device" (what a trainer sees: the towers emit them on the GPU)
    small = {k: host[k] for k in ("target", "item_idx", "pos_idx")}
    small_bytes = sum(v.numel() * v.element_size() for v in small.values())
    dev_embed = {"user_embed": inp["user_embed"], "item_embed": inp["item_embed"]}

    def e2e_small(n: int) -> None:
        prev = None
        for _ in range(n):
            res = stepper.submit({**dev_embed, **small})
            if prev is not None:
                prev.loss_value()
            prev = res
        prev.loss_value()

    e2e_small(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_small(e2e_steps)
    torch.cuda.synchronize()
    e2e_small_value = C2["batch"] * e2e_steps / (time.perf_counter() - t0)

    pk = peaks()
    algo_flops = 6.0 * C2["batch"] * C2["num_items"] * C2["dim"]  # fwd 2BNd + dQ 2BNd + dI 2BNd
    sweep_ms_per_step = sweep_ms_total / steps
    achieved = algo_flops / (sweep_ms_per_step * 1e-3) / 1e12
    out = {
        "metric": LOSS_METRIC, "value": C2["batch"] / (ms_per_step * 1e-3), "unit": "samples/s", "ms_per_step": ms_per_step,
        "steps": steps, "warmup": warmup, "dtype": "bf16",
        "config": {"workload": "C2 MovieLens-32M-shaped: batch 4096 x 87585 items, d=128, P=32, bf16, sampled-softmax "
                               f"({HEADLINE_LOSS}) fwd+bwd through the drop-in module; one rank (users shard with no data-path collective)",
                   "sigma": SIGMA, "margin": MARGIN, "num_negatives": 0, "l2": "flushed between steps (256 MiB read+write)",
                   "launch": "CUDA-graph replay of the module's forward+backward (eager ms in eager_ms_per_step)"},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                "embeddings_on_device": {"value": e2e_small_value, "unit": "samples/s", "h2d_bytes_per_step": small_bytes,
                                         "note": "target / item_idx / pos_idx from pinned host memory, embeddings resident (the towers "
                                                 "produce them on the device in a trainer)"},
                "pipeline": "xfmr_b200.GraphedLossStep: CUDA-graph replay per input slot; the pinned-host upload of step i+1 runs on a "
                            "copy stream under step i, the loss of step i-1 is read on the host while step i runs"},
        "gpu_launches": launches, "eager_ms_per_step": eager_ms,
        "roofline": {
            "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
            "frac": achieved / pk["bf16_tflops"], "traffic": ncu_traffic("loss_sweep_dram_bytes_per_launch"),
            "kernel": "xb::wg_kernel (2 working launches per step: merged forward + dQ sweep, dI sweep) + 2 conditional fallback "
                      "launches of xb::sweep_kernel that exit at once",
            "algorithmic_flops_per_step": algo_flops, "sweep_ms_per_step": sweep_ms_per_step,
            "sweep_launches_per_step": sweep_count / steps,
            "peak_source": pk["source"] + " BURST bf16 (20 steps of ~0.4 ms: a 10 ms burst at boost clocks)",
            "frac_of_sustained_peak": achieved / pk["bf16_tflops_sustained"],
            "sweep_share_of_step": sweep_ms_per_step / ms_per_step,
        },
    }
    if not args.no_extras:
        per_loss = {}

        def time_module(m, key: str, inputs: dict = inp) -> None:  # noqa: ANN001
            t = statistics.median(timed_steps(graphed(loss_step_fn(m, inputs)), 8, 3, flush_buf))
            per_loss[key] = {"ms_per_step": t, "samples_per_s": inputs["user_embed"].size(0) / (t * 1e-3)}

        for name in xfmr_b200.LOSS_SLOTS:
            time_module(getattr(xfmr_b200, name)(sigma=SIGMA, margin=MARGIN), name)
        time_module(xfmr_b200.PairwiseHingeLoss(num_negatives=4, sigma=SIGMA, margin=MARGIN), "PairwiseHingeLoss[num_negatives=4]")
        time_module(xfmr_b200.PairwiseHingeLoss(num_negatives=4, sigma=SIGMA, margin=MARGIN, mining="hard"),
                    "PairwiseHingeLoss[num_negatives=4, mining=hard]")
        time_module(xfmr_b200.DirectAULoss(gamma=1.0, t=2.0), "DirectAULoss")
        for n_rows in (C2["batch"], C2["num_items"]):
            x = inp["item_embed"][:n_rows].detach().requires_grad_(True)

            def uni_step(x=x) -> tuple:  # noqa: ANN001
                loss = xfmr_b200.uniformity_loss(x, 2.0)
                return loss, torch.autograd.grad(loss, x)[0]

            ms = statistics.median(timed_steps(graphed(uni_step), 8, 3, flush_buf))
            # one sweep of 2 tile-MMAs per tile pair does forward and backward: 4 n^2 d algorithmic flops
            per_loss[f"uniformity_loss[n={n_rows}]"] = {"ms_per_step": ms, "tflops": 4.0 * n_rows * n_rows * C2["dim"] / (ms * 1e-3) / 1e12}

        def fused_fwd() -> None:
            xfmr_b200.fused_losses(inp["user_embed"], inp["item_embed"], inp["target"], item_idx=inp["item_idx"],
                                   pos_idx=inp["pos_idx"], sigma=SIGMA, margin=MARGIN)

        per_loss["all_seven_forward_one_call"] = {"ms_per_step": statistics.median(timed_steps(fused_fwd, 8, 3, flush_buf))}
        out["per_loss"] = per_loss
    return out


def bench_c1(device: torch.device) -> dict:
    """BASELINE config 1 (the reference's own CPU-runnable case): 1,024 queries x 3,706 items, d=64 fp32, sampled softmax
    with LogQ (forward + backward) and exact top-10 retrieval.  Launch / latency bound: reported in microseconds."""
    import xfmr_b200  # noqa: PLC0415

    synthetic = load_synthetic()
    inp = synthetic.make_loss_inputs(C1["batch"], C1["num_items"], C1["dim"], C1["num_pos"], n_catalog=C1["num_items"], seed=0,
                                     device=device)
    inp["log_q_arg"] = inp["log_q"]
    module = getattr(xfmr_b200, HEADLINE_LOSS)(sigma=1.0, margin=1.0)
    eager = loss_step_fn(module, inp)
    t_graph = statistics.median(timed_steps(graphed(eager), 20, 5, None))
    t_eager = statistics.median(timed_steps(eager, 20, 5, None))
    queries, items = inp["user_embed"], inp["item_embed"]
    excl = torch.randint(1, C1["num_items"] + 1, (C1["batch"], 20), device=device)
    index = xfmr_b200.ItemProcessor().get_index(items, torch.arange(1, C1["num_items"] + 1, device=device))

    def search() -> None:
        index.search_batch(queries, excl, top_k=C1["k"])

    t_search = statistics.median(timed_steps(search, 20, 5, None))
    return {"workload": "C1: 1024 x 3706, d=64 fp32 (split-bf16 contraction, exact ids), P=32; sampled softmax with LogQ fwd+bwd; exact "
                        "top-10 with 20 excluded ids per query",
            "loss_fwd_bwd_us": {"graph_replay": t_graph * 1e3, "eager": t_eager * 1e3},
            "top10_search_us": t_search * 1e3, "note": "launch / latency bound (1.5 GFLOP = ~1 us of tensor time): microseconds, not a roofline fraction"}


def bench_mns(device: torch.device, world: int, rank: int) -> dict:
    """Config 3: mixed negative sampling with global negatives - per rank 8,192 users, their 8,192 in-batch items and
    16,384 uniform negatives, d=256 bf16; items and negatives are all-gathered over NCCL (every rank scores its users
    against 24,576 x world candidates) and the item gradients are reduced back to their owners in the backward pass."""
    import torch.distributed as dist  # noqa: PLC0415

    import xfmr_b200  # noqa: PLC0415

    synthetic = load_synthetic()
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
        step = graphed(step)      # one rank: no collective in the step, replay it like the loss line
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
            "tensor_frac_of_sustained_peak": flops / (ms * 1e-3) / (pk["bf16_tflops_sustained"] * 1e12),
            "tensor_frac_of_burst_peak": flops / (ms * 1e-3) / (pk["bf16_tflops"] * 1e12)}


def bench_evaluate(device: torch.device) -> dict:
    """Batched validation at MovieLens-1M shape: every user at once - exact top-20 with the user's
    history excluded + the six ranking metrics - where the reference runs one user per step (lightning.py:149-206)."""
    import xfmr_b200  # noqa: PLC0415

    synthetic = load_synthetic()
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

    ms = statistics.median(timed_steps(step, 5, 3, None))
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
    ms = statistics.median(timed_steps(fn, 10, 3, None))
    algo_bytes = n * (8 + 2 * d * 2 + d * 2)
    pk = peaks()
    return {"metric": "hash_gather_GBps", "value": algo_bytes / (ms * 1e-3) / 1e9, "unit": "GB/s", "ms": ms,
            "workload": "C4: 4Mi ids, k=2 hashes, 2^22 x 128 bf16 table", "frac_of_hbm_peak": algo_bytes / (ms * 1e-3) / 1e9 / pk["hbm_gbs"],
            "traffic": ncu_traffic("hash_gather_dram_bytes_per_launch"), "algorithmic_bytes": algo_bytes}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="headline + loss lines only (skip per-loss / C1 / C3 / C4 / cpu baselines)")
    ap.add_argument("--only-loss", action="store_true", help="development: the config-2 loss section alone (prints its object)")
    ap.add_argument("--retrieval-items", type=int, default=C5["num_items"],
                    help="catalog rows summed over all ranks (config 5: 100,000,000 = 25.6 GB bf16)")
    ap.add_argument("--retrieval-queries", type=int, default=C5["num_queries"])
    ap.add_argument("--retrieval-shards", type=int, default=0,
                    help="catalog shards R (a divisor of N); the N / R query groups each search a slice of the queries. "
                         "0 = N: plain row sharding")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist  # noqa: PLC0415

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries the one JSON line and nothing else: native libraries write there too (NCCL prints its version
    # banner with printf at NCCL_DEBUG=VERSION and WARN), so file descriptor 1 points at stderr until the line is ready
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"  # noqa: S101

    if args.only_loss:
        line = bench_loss(device, world, rank, args)
    else:
        clocks = ClockSampler(local_rank)
        clocks.start()
        r = bench_retrieval(device, world, rank, args, clocks)
        clocks.stop()
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(args.retrieval_items, args.retrieval_queries),
            "clocks": clocks.summary(), "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "roofline": r["roofline"],
            "recall_at_k_sampled": r["recall_at_k_sampled"], "with_64_exclusions_per_query": r["with_64_exclusions_per_query"],
            "parallelism": (f"catalog row-sharded {r['layout']['catalog_shards']} ways x {r['layout']['query_groups']} query groups "
                            f"({r['layout']['items_per_rank']} items and {r['layout']['queries_per_rank']} queries per rank); per-shard "
                            "top-k lists exchanged with all_to_all inside a query group, merged per query slice, all-gathered over all ranks"),
            "layout": r["layout"],
        }
        if rank == 0 and world == 1:
            line["loss"] = bench_loss(device, world, rank, args)
        if not args.no_extras:
            torch.cuda.empty_cache()
            line["mns"] = bench_mns(device, world, rank)
            if rank == 0 and world == 1:
                line["c1"] = bench_c1(device)
                line["evaluate"] = bench_evaluate(device)
                line["gather"] = bench_gather(device)
        if rank == 0 and world == 1:
            # the reference's side on this box's host cores, bounded samples (10-30 s of CPU work in total)
            res = cpu_topk_sample(256, 1_000_000, 3, 1)
            ms = min(res["times_s"]) * 1e3
            line["cpu_baseline"] = {
                "value": 256 / (ms * 1e-3) * (1_000_000 / args.retrieval_items), "unit": UNIT, "cores": res["cores"], "kind": "port",
                "sample": f"exact brute-force top-{C5['k']} (torch fp32 matmul + topk) on 256 queries x 1000000 items, best of 3: {ms:.0f} ms; "
                          f"scaled by 1000000 / {args.retrieval_items} items to the full-catalog rate"}
            if not args.no_extras:
                lres = cpu_loss_full_c2(1, 1)
                lms = min(lres["times_s"]) * 1e3
                line["loss"]["cpu_baseline"] = {
                    "value": C2["batch"] / (lms * 1e-3), "unit": "samples/s", "cores": lres["cores"], "kind": lres["kind"],
                    "sample": f"{HEADLINE_LOSS} fwd + autograd.grad at the FULL config 2 (4096 x 87585, d=128, P=32, fp32), one run after a warm-up: {lms:.0f} ms"}

    if rank == 0:
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
