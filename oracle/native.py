"""ctypes access to ``oracle/libxb_oracle.so`` (``topk_oracle.c``): exact top-k and XXH32 restatements.

CPU ORACLE — test infrastructure, not product.  Top-k parity is UNPINNED by the reference (its search is
an approximate third-party index, see the header of ``topk_oracle.c``); the XXH32 restatement is pinned
against python-xxhash 3.7.0 and the known answers of SURVEY.md Appendix C in ``tests/test_oracle_golden.py``.
"""

from __future__ import annotations

import ctypes
import pathlib
import subprocess

import numpy as np

_DIR = pathlib.Path(__file__).resolve().parent
_LIB = _DIR / "libxb_oracle.so"
PAD_ID = -(2**63)


def build() -> None:
    subprocess.run(["make", "-C", str(_DIR)], check=True, capture_output=True)  # noqa: S603, S607


def _load() -> ctypes.CDLL:
    if not _LIB.exists():
        build()
    lib = ctypes.CDLL(str(_LIB))
    lib.xbo_xxh32_i64.restype = ctypes.c_uint32
    lib.xbo_xxh32_i64.argtypes = [ctypes.c_int64, ctypes.c_uint32]
    return lib


_lib = _load()


def _p(a: np.ndarray | None) -> ctypes.c_void_p | None:
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def topk(queries, items, k, *, item_ids=None, id_base=0, exclude=None):  # noqa: ANN001, ANN201
    """``(scores [Q, k] float32, ids [Q, k] int64)``; ``exclude`` is ``[Q, E]`` int64 padded with ``PAD_ID``."""
    q = np.ascontiguousarray(queries, dtype=np.float32)
    it = np.ascontiguousarray(items, dtype=np.float32)
    ids = None if item_ids is None else np.ascontiguousarray(item_ids, dtype=np.int64)
    ex = None if exclude is None else np.ascontiguousarray(exclude, dtype=np.int64)
    nq, d = q.shape
    n = it.shape[0]
    scores = np.empty((nq, k), dtype=np.float32)
    out_ids = np.empty((nq, k), dtype=np.int64)
    _lib.xbo_topk(
        _p(q), _p(it), _p(ids), ctypes.c_int64(id_base), _p(ex), ctypes.c_int(0 if ex is None else ex.shape[1]),
        ctypes.c_int(nq), ctypes.c_int(n), ctypes.c_int(d), ctypes.c_int(k), _p(scores), _p(out_ids),
    )
    return scores, out_ids


def xxh32_i64(value: int, seed: int) -> int:
    return int(_lib.xbo_xxh32_i64(ctypes.c_int64(value), ctypes.c_uint32(seed)))


def hash_indices(ids, num_hashes, log2_rows, seed0=0):  # noqa: ANN001, ANN201
    a = np.ascontiguousarray(ids, dtype=np.int64)
    out = np.empty((a.size, num_hashes), dtype=np.int32)
    _lib.xbo_hash_indices(_p(a), ctypes.c_int64(a.size), ctypes.c_int(num_hashes), ctypes.c_uint32(seed0),
                          ctypes.c_int(log2_rows), _p(out))
    return out


def hash_gather(table_f32, ids, num_hashes, log2_rows, seed0=0):  # noqa: ANN001, ANN201
    """fp32 sum over the hashed rows in hash order (``embedding_bag(mode="sum")``); caller rounds to bf16."""
    idx = hash_indices(ids, num_hashes, log2_rows, seed0)
    acc = np.zeros((idx.shape[0], table_f32.shape[1]), dtype=np.float32)
    for h in range(num_hashes):
        acc = acc + table_f32[idx[:, h]]
    return acc, idx
