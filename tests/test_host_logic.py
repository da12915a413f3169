"""Host-side behaviour that needs no GPU: argument checking with the reference's error behaviour
(xfmr_rec/losses.py:54-79), loud failure instead of a CPU fallback, the class surface, fake-tensor shapes."""

from __future__ import annotations

import pytest
import torch

import xfmr_b200
from xfmr_b200 import synthetic


def test_same_seven_class_names_and_constructor_keywords() -> None:
    names = [
        "AlignmentLoss", "ContrastiveLoss", "AlignmentContrastiveLoss", "InfomationNoiseContrastiveEstimationLoss",
        "MutualInformationNeuralEstimationLoss", "PairwiseHingeLoss", "PairwiseLogisticLoss",
    ]
    for n in names:
        module = getattr(xfmr_b200, n)(num_negatives=4, sigma=2.0, margin=0.5)   # lightning.py:279-286
        assert isinstance(module, torch.nn.Module)
        assert (module.num_negatives, module.sigma, module.margin) == (4, 2.0, 0.5)
        assert type(module).__name__ == n                                       # selected by class name, lightning.py:139
    torch.nn.ModuleList([c() for c in xfmr_b200.LOSS_CLASSES])                  # lightning.py:270-278


@pytest.mark.parametrize(("q", "v", "t"), [
    (torch.zeros(4, 8, 2), torch.zeros(6, 8), torch.zeros(4)),     # rank != 2
    (torch.zeros(4, 8), torch.zeros(8), torch.zeros(4)),
    (torch.zeros(4, 8), torch.zeros(6, 9), torch.zeros(4)),         # width mismatch
    (torch.zeros(4, 8), torch.zeros(6, 8), torch.zeros(5)),         # target rows != user rows
    (torch.zeros(4, 8), torch.zeros(3, 8), torch.zeros(4)),         # fewer items than users
])
def test_shape_errors_are_value_errors(q: torch.Tensor, v: torch.Tensor, t: torch.Tensor) -> None:
    loss = xfmr_b200.PairwiseHingeLoss()
    with pytest.raises(ValueError):  # noqa: PT011
        loss(q, v, t, item_idx=torch.arange(v.size(0)), pos_idx=torch.zeros(4, 1, dtype=torch.int64))


def test_cpu_tensors_fail_loudly() -> None:
    inp = synthetic.make_loss_inputs(8, 16, 8, 2, seed=0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        xfmr_b200.InfomationNoiseContrastiveEstimationLoss()(inp["user_embed"], inp["item_embed"], inp["target"],
                                                              item_idx=inp["item_idx"], pos_idx=inp["pos_idx"])
    with pytest.raises(RuntimeError, match="no CPU path"):
        xfmr_b200.topk_search(inp["user_embed"], inp["item_embed"], 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        xfmr_b200.hash_indices(torch.arange(4), 2, 10)
    with pytest.raises(RuntimeError, match="no CPU path"):
        xfmr_b200.hash_embedding_gather(torch.zeros(16, 8, dtype=torch.bfloat16), torch.arange(4), 2)


def test_fake_tensor_shapes() -> None:
    """register_fake: shape inference without touching the GPU (torch.compile / export tracing)."""
    from torch._subclasses.fake_tensor import FakeTensorMode  # noqa: PLC0415

    with FakeTensorMode():
        q = torch.empty(64, 32, device="cuda")
        v = torch.empty(200, 32, device="cuda")
        t = torch.empty(64, device="cuda")
        idx = torch.empty(200, dtype=torch.int64, device="cuda")
        pos = torch.empty(64, 4, dtype=torch.int64, device="cuda")
        losses, ws = torch.ops.xfmr_b200.loss_fwd(q, v, t, idx, pos, None, 0, 1.0, 1.0, 127, 1, 0)
        assert losses.shape == (7,) and losses.dtype == torch.float32
        assert ws.dtype == torch.uint8 and ws.numel() > 0
        dq, dv = torch.ops.xfmr_b200.loss_bwd(ws, losses, 64, 200, 32, 4, False, False, 0, 1.0, 1.0, 127, 1, 0)
        assert dq.shape == (64, 32) and dv.shape == (200, 32)
        loss, ws_u = torch.ops.xfmr_b200.uniformity_fwd(q, 2.0, 1)
        assert loss.shape == (1,) and ws_u.numel() > ws.numel() // 4
        assert torch.ops.xfmr_b200.uniformity_bwd(ws_u, loss, 64, 32, False, 2.0, 1).shape == (64, 32)


def test_unsupported_shapes_are_reported_by_the_library() -> None:
    import ctypes  # noqa: PLC0415

    from xfmr_b200 import _lib  # noqa: PLC0415

    wide = _lib.TopkDesc(num_queries=4, num_items=100, dim=300, k=10, in_dtype=0, compute=0, has_exclusions=0, reserved=0, id_base=0)
    assert _lib.lib.xb_topk_workspace_bytes(ctypes.byref(wide)) == 0
    assert b"dim" in _lib.lib.xb_last_error_string()
    big_k = _lib.TopkDesc(num_queries=4, num_items=100, dim=64, k=1000, in_dtype=0, compute=0, has_exclusions=0, reserved=0, id_base=0)
    assert _lib.lib.xb_topk_workspace_bytes(ctypes.byref(big_k)) == 0
    mined = _lib.LossDesc(batch=8, num_items=1000, dim=64, num_pos=0, in_dtype=0, compute=1, num_negatives=100,
                          loss_mask=8, sigma=1.0, margin=1.0, has_log_q=0, mining=0)
    assert _lib.lib.xb_loss_workspace_bytes(ctypes.byref(mined)) == 0


def test_synthetic_inputs_follow_the_reference_layout() -> None:
    inp = synthetic.make_loss_inputs(64, 200, 16, 8, n_catalog=500, seed=1)
    assert inp["item_idx"].min() >= 1                       # 0 is reserved for padding (data/prepare.py:85)
    assert (inp["pos_idx"][:, 0] == inp["item_idx"][:64]).all()
    assert inp["pos_idx"].dtype == torch.int64 and (inp["pos_idx"] >= 0).all()
    assert torch.allclose(inp["user_embed"].norm(dim=-1), torch.ones(64), atol=1e-5)   # models.py:59
    assert set(inp["target"].tolist()) <= {1.0, 2.0, 3.0, 4.0, 5.0}                    # ratings, params.py:8
    assert inp["item_idx"][64:].unique().numel() == 136                                  # negatives without replacement


def test_arrow_catalog_adapter_reads_the_reference_table_layout(tmp_path) -> None:  # noqa: ANN001
    """``movie_rn, movie_id, movie_text, embedding: fixed_size_list<float32, d>`` (data/lightning.py:208-219)."""
    import numpy as np  # noqa: PLC0415
    import pyarrow as pa  # noqa: PLC0415
    import pyarrow.parquet as pq  # noqa: PLC0415

    rng = np.random.default_rng(0)
    emb = rng.standard_normal((37, 8)).astype(np.float32)
    ids = rng.permutation(1000)[:37].astype(np.int64) + 1
    texts = [f"movie {i}" for i in ids]
    table = pa.table({
        "movie_rn": pa.array(np.arange(1, 38), type=pa.int64()),
        "movie_id": pa.array(ids),
        "movie_text": pa.array(texts),
        "embedding": pa.FixedSizeListArray.from_arrays(pa.array(emb.reshape(-1)), 8),
    })
    got_emb, got_ids, got_texts = xfmr_b200.arrow_to_catalog(table, id_col="movie_id", text_col="movie_text")
    assert np.array_equal(got_emb, emb) and np.array_equal(got_ids, ids) and got_texts == texts
    # through parquet (chunked columns) and with a variable-length list column
    pq.write_table(table, tmp_path / "items.parquet", row_group_size=10)
    back = pq.read_table(tmp_path / "items.parquet")
    assert np.array_equal(xfmr_b200.arrow_to_catalog(back, id_col="movie_id", text_col=None)[0], emb)
    as_list = table.set_column(3, "embedding", pa.array(emb.tolist(), type=pa.list_(pa.float32())))
    assert np.array_equal(xfmr_b200.arrow_to_catalog(as_list, id_col="movie_id", text_col="movie_text")[0], emb)
    ragged = table.set_column(3, "embedding", pa.array([[1.0]] * 36 + [[1.0, 2.0]], type=pa.list_(pa.float32())))
    with pytest.raises(ValueError, match="ragged"):
        xfmr_b200.arrow_to_catalog(ragged, id_col="movie_id", text_col=None)
    # the index itself lives on the GPU: no CPU index
    with pytest.raises((RuntimeError, AssertionError)):
        xfmr_b200.ItemProcessor().get_index_from_arrow(table, device="cpu").search_batch(torch.zeros(1, 8), None, 5)


def test_reference_arm_prints_one_contract_line() -> None:
    """``bench.py --impl reference``: the reference's side timed on the host cores, one JSON line with the contract keys,
    the requested steps honoured, measured (not extrapolated) milliseconds, and the product library never mapped."""
    import json  # noqa: PLC0415
    import pathlib  # noqa: PLC0415
    import subprocess  # noqa: PLC0415
    import sys  # noqa: PLC0415

    root = pathlib.Path(__file__).resolve().parents[1]
    probe = (
        "import runpy, sys; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '2', '--warmup', '1', '--no-extras'];"
        "runpy.run_path(r'%s', run_name='__main__');"
        "maps = open('/proc/self/maps').read(); sys.stderr.write('MAPPED_PRODUCT=%%d\\n' %% ('libxfmr_b200' in maps))"
    ) % str(root / "bench.py")
    out = subprocess.run([sys.executable, "-c", probe], capture_output=True, text=True, check=True, timeout=900)  # noqa: S603
    assert "MAPPED_PRODUCT=0" in out.stderr, "the reference arm must not load libxfmr_b200.so"
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "exact_top100_retrieval_queries_per_s" and line["unit"] == "queries/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1
    assert line["steps"] == 2 and line["warmup"] == 1 and line["ms_per_step"] > 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["vs_baseline"] is None
    # both arms print the same config object
    import importlib.util  # noqa: PLC0415

    spec = importlib.util.spec_from_file_location("xb_bench", root / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert line["config"] == bench.workload_config(bench.C5["num_items"], bench.C5["num_queries"])


def test_exclusion_path_selection_and_padding_helpers() -> None:
    """Host logic of ``ItemProcessor.search_batch``: dense bit mask while it is small, sparse post-filter otherwise,
    a clear error when neither fits (no CUDA call is reached on these paths)."""
    from xfmr_b200 import retrieval  # noqa: PLC0415

    padded = retrieval.pad_id_lists([[3, 4], [], [5]])
    assert padded.shape == (3, 2) and padded[1].tolist() == [retrieval._PAD_ID] * 2 and padded[2, 0] == 5  # noqa: SLF001
    assert retrieval.pad_id_lists([[], []]).shape == (2, 1)                      # at least one column
    vals = retrieval.pad_id_lists([[1.5], []], pad=0, dtype=torch.float32)
    assert vals.dtype == torch.float32 and vals.tolist() == [[1.5], [0.0]]

    index = xfmr_b200.ItemProcessor().get_index(torch.zeros(1000, 8), device="cpu")
    assert index._exclusions(None, 4, 20) == (None, None)                        # noqa: SLF001
    assert index._exclusions([[], [], [], []], 4, 20) == (None, None)            # noqa: SLF001
    index.DENSE_MASK_BYTES = 0                                                   # force the sparse decision
    mask, sparse = index._exclusions([[1, 2, 3], [4], [], [5, 6]], 4, 20)        # noqa: SLF001
    assert mask is None and sparse.shape == (4, 3) and sparse[2].tolist() == [retrieval._PAD_ID] * 3  # noqa: SLF001
    with pytest.raises(ValueError, match="one exclusion list per query"):
        index._exclusions([[1]], 4, 20)                                          # noqa: SLF001
    with pytest.raises(ValueError, match="exclusion lists of 250 ids"):
        index._exclusions(torch.zeros(4, 250, dtype=torch.int64), 4, 20)         # noqa: SLF001
    # metrics / filter / graph wrappers refuse CPU tensors like everything else
    with pytest.raises(RuntimeError, match="no CPU path"):
        xfmr_b200.retrieval_metrics(torch.zeros(2, 3, dtype=torch.int64), torch.zeros(2, 1, dtype=torch.int64), torch.zeros(2, 1))
    with pytest.raises(RuntimeError, match="no CPU path"):
        xfmr_b200.topk_filter(torch.zeros(2, 5), torch.zeros(2, 5, dtype=torch.int64), torch.zeros(2, 1, dtype=torch.int64), 3)
    inp = synthetic.make_loss_inputs(8, 16, 8, 2, seed=0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        xfmr_b200.GraphedLossStep(xfmr_b200.PairwiseHingeLoss(), inp)
    with pytest.raises(RuntimeError, match="no CPU path"):
        xfmr_b200.DirectAULoss()(inp["user_embed"], inp["item_embed"], inp["target"], item_idx=inp["item_idx"], pos_idx=inp["pos_idx"])
    with pytest.raises(RuntimeError, match="no CPU path"):   # even when only the torch-op margin term would be left
        xfmr_b200.MAWULoss(gamma_user=0.0, gamma_item=0.0)(
            inp["user_embed"], inp["item_embed"], inp["target"], item_idx=inp["item_idx"], pos_idx=inp["pos_idx"],
            user_margin=torch.zeros(8), item_margin=torch.zeros(8))


def test_search_frame_assembly_with_a_stubbed_kernel(monkeypatch) -> None:  # noqa: ANN001
    """``ItemProcessor.search`` (data/lightning.py:237-259): the reference's result columns (every table column + score),
    empty slots dropped, rows looked up by id; then the reference's own call site, ``recommend``
    (xfmr_rec/lightning.py:93-95: ``.search(...).drop(columns="embedding")``), and the BentoML ``ItemCandidate`` fields
    (bentoml/service.py:52-55, :129-131: ``movie_id, movie_text, score``) replayed on the frame.  The kernel call is
    stubbed: only the host-side assembly runs here."""
    emb = torch.arange(32, dtype=torch.float32).reshape(4, 8) + 1.0
    index = xfmr_b200.ItemProcessor(id_col="movie_id", text_col="movie_text").get_index(
        emb, torch.tensor([40, 10, 30, 20]), ["forty", "ten", "thirty", "twenty"], item_idx=torch.tensor([1, 2, 3, 4]), device="cpu")
    assert torch.allclose(index.embeddings.norm(dim=-1), torch.ones(4), atol=1e-6)          # metric="cosine": unit rows
    calls = []

    def fake_search_batch(embedding, exclude, top_k):  # noqa: ANN001, ANN202
        calls.append((tuple(embedding.shape), exclude, top_k))
        return torch.tensor([[0.9, 0.5, float("-inf")]]), torch.tensor([[30, 40, -1]])

    monkeypatch.setattr(index, "search_batch", fake_search_batch)
    frame = index.search(torch.zeros(8).numpy(), exclude_item_ids=[10], top_k=3)
    assert calls == [((1, 8), [[10]], 3)]
    assert list(frame.columns) == ["movie_rn", "movie_id", "movie_text", "embedding", "score"]
    assert frame["movie_id"].tolist() == [30, 40] and frame["movie_text"].tolist() == ["thirty", "forty"]
    assert frame["movie_rn"].tolist() == [3, 1]
    assert frame["score"].tolist() == pytest.approx([0.9, 0.5])
    assert torch.allclose(torch.tensor(frame["embedding"].iloc[0]), index.embeddings[2])
    # the reference's caller and the serving schema
    dropped = frame.drop(columns="embedding")
    assert list(dropped.columns) == ["movie_rn", "movie_id", "movie_text", "score"]
    for rec in dropped.to_dict(orient="records"):
        assert isinstance(rec["movie_id"], int) and isinstance(rec["movie_text"], str) and isinstance(rec["score"], float)
    frame2 = index.search(torch.zeros(1, 8).numpy(), None, top_k=3)            # no exclusions: None is passed through
    assert calls[-1][1] is None and frame2["movie_text"].tolist() == ["thirty", "forty"]
    assert index._row_of_id == {40: 0, 10: 1, 30: 2, 20: 3}                     # noqa: SLF001  built once, reused
    with pytest.raises(ValueError, match="non-negative"):
        xfmr_b200.ItemProcessor().get_index(torch.zeros(2, 8), torch.tensor([3, -7]), device="cpu")
    raw = xfmr_b200.ItemProcessor(metric="dot").get_index(emb, device="cpu")
    assert torch.equal(raw.embeddings, emb)                                      # metric="dot": rows as given


def test_processors_json_follows_the_reference_schema(tmp_path) -> None:  # noqa: ANN001
    """``save`` writes ``processors.json["items"]`` with every field of the reference's ``ItemProcessor.model_dump()``
    (xfmr_rec/lightning.py:318-322; fields at data/lightning.py:79-81, :128-133, :154-165) and the item table with the
    reference's columns incl. ``movie_rn``; a reference-style entry constructs the processor."""
    import json  # noqa: PLC0415

    import pyarrow.parquet as pq  # noqa: PLC0415

    index = xfmr_b200.ItemProcessor().get_index(torch.eye(3, 8), torch.tensor([7, 8, 9]), ["a", "b", "c"],
                                                item_idx=torch.tensor([1, 2, 3]), device="cpu")
    index.save(tmp_path / "bundle")
    items = json.loads((tmp_path / "bundle" / "processors.json").read_text())["items"]
    reference_fields = ["batch_size", "data_dir", "idx_col", "id_col", "text_col", "lance_table_name", "lance_db_path",
                        "num_partitions", "num_sub_vectors", "num_probes", "refine_factor"]
    assert list(items)[: len(reference_fields)] == reference_fields
    assert items["idx_col"] == "movie_rn" and items["lance_table_name"] == "movies" and items["num_probes"] == 8
    table = pq.read_table(tmp_path / "bundle" / "items.parquet")
    assert table.column_names == ["movie_rn", "movie_id", "movie_text", "embedding"]
    import pyarrow as pa  # noqa: PLC0415

    etype = table.schema.field("embedding").type
    assert pa.types.is_fixed_size_list(etype) and etype.list_size == 8 and etype.value_type == pa.float32()
    # a reference export's entry (no keys of ours) is accepted as constructor arguments
    ref_entry = {k: items[k] for k in reference_fields}
    proc = xfmr_b200.ItemProcessor(**ref_entry)
    assert proc.idx_col == "movie_rn" and proc.reference_args["refine_factor"] == 4 and proc.metric == "cosine"
    with pytest.raises(TypeError, match="unknown"):
        xfmr_b200.ItemProcessor(bogus=1)


def test_shim_submodules_are_the_loaded_modules() -> None:
    """``from xfmr_b200.losses import ...`` must not execute ``losses.py`` a second time: re-defining the custom ops
    destroys the torch library of the first definition and leaves the loss classes with a dangling operator."""
    import importlib  # noqa: PLC0415
    import sys  # noqa: PLC0415

    import xfmr_b200  # noqa: PLC0415
    from xfmr_b200.losses import _loss_fwd  # noqa: PLC0415

    assert _loss_fwd is xfmr_b200.losses._loss_fwd  # noqa: SLF001
    for sub in ("losses", "retrieval", "hashing", "synthetic", "graphs", "distributed", "uniformity", "_lib"):
        assert importlib.import_module(f"xfmr_b200.{sub}") is sys.modules[f"matrix-factorization-torch_b200.{sub}"]


def test_committed_ncu_traffic_has_the_keys_bench_reads() -> None:
    """``bench.py`` fills ``roofline.traffic`` from ``profiles/ncu_traffic.json``: a renamed key would silently turn into null."""
    import json  # noqa: PLC0415
    import pathlib  # noqa: PLC0415
    import re  # noqa: PLC0415

    root = pathlib.Path(__file__).resolve().parents[1]
    keys = set(re.findall(r'ncu_traffic\("([a-z_]+)"\)', (root / "bench.py").read_text()))
    data = json.loads((root / "profiles" / "ncu_traffic.json").read_text())
    assert keys and keys <= set(data), keys - set(data)
    assert all(float(data[k]) > 0 for k in keys)
