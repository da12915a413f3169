"""ctypes binding of ``libxfmr_b200.so`` (C ABI declared in ``include/xfmr_b200.h``).

The shared library is the product: there is no Python or CPU fallback for any entry point. If the
library has not been built, importing this module raises; if a CPU tensor reaches a compute call, the
caller raises (see ``require_cuda``).
"""

from __future__ import annotations

import ctypes
import os
import pathlib

import torch

_PKG_DIR = pathlib.Path(__file__).resolve().parent
# XB_LIB selects a debug build of the same library (e.g. the -DXB_TRACE one used by tests/quick_probe.py)
LIB_PATH = pathlib.Path(os.environ["XB_LIB"]) if os.environ.get("XB_LIB") else _PKG_DIR / "libxfmr_b200.so"

XB_OK = 0
XB_DTYPE_F32 = 0
XB_DTYPE_BF16 = 1
XB_COMPUTE_BF16 = 0
XB_COMPUTE_SPLIT = 1
XB_NUM_LOSSES = 7
XB_MINING_SEMI_HARD = 0
XB_MINING_HARD = 1

ERROR_NAMES = {-1: "invalid argument", -2: "unsupported shape", -3: "workspace too small", -4: "CUDA error"}


class XbError(RuntimeError):
    """A libxfmr_b200 entry point returned a negative status."""


class LossDesc(ctypes.Structure):
    """``xb_loss_desc`` (include/xfmr_b200.h)."""

    _fields_ = [
        ("batch", ctypes.c_int32),
        ("num_items", ctypes.c_int32),
        ("dim", ctypes.c_int32),
        ("num_pos", ctypes.c_int32),
        ("in_dtype", ctypes.c_int32),
        ("compute", ctypes.c_int32),
        ("num_negatives", ctypes.c_int32),
        ("loss_mask", ctypes.c_uint32),
        ("sigma", ctypes.c_float),
        ("margin", ctypes.c_float),
        ("has_log_q", ctypes.c_int32),
        ("mining", ctypes.c_int32),
    ]


class UniformityDesc(ctypes.Structure):
    """``xb_uniformity_desc`` (include/xfmr_b200.h)."""

    _fields_ = [
        ("n", ctypes.c_int32),
        ("dim", ctypes.c_int32),
        ("in_dtype", ctypes.c_int32),
        ("compute", ctypes.c_int32),
        ("t", ctypes.c_float),
        ("reserved", ctypes.c_int32),
    ]


class TopkDesc(ctypes.Structure):
    """``xb_topk_desc`` (include/xfmr_b200.h)."""

    _fields_ = [
        ("num_queries", ctypes.c_int32),
        ("num_items", ctypes.c_int32),
        ("dim", ctypes.c_int32),
        ("k", ctypes.c_int32),
        ("in_dtype", ctypes.c_int32),
        ("compute", ctypes.c_int32),
        ("has_exclusions", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("id_base", ctypes.c_int64),
    ]


# name -> (restype, argtypes); the single source of truth for the symbol-export test
_vp, _i32, _i64, _u32, _sz, _f = (
    ctypes.c_void_p,
    ctypes.c_int32,
    ctypes.c_int64,
    ctypes.c_uint32,
    ctypes.c_size_t,
    ctypes.c_float,
)
SIGNATURES = {
    "xb_loss_workspace_bytes": (_sz, [ctypes.POINTER(LossDesc)]),
    "xb_loss_forward": (_i32, [ctypes.POINTER(LossDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xb_loss_backward": (_i32, [ctypes.POINTER(LossDesc), _vp, _vp, _vp, _vp, _sz, _vp]),
    "xb_uniformity_workspace_bytes": (_sz, [ctypes.POINTER(UniformityDesc)]),
    "xb_uniformity_forward": (_i32, [ctypes.POINTER(UniformityDesc), _vp, _vp, _vp, _sz, _vp]),
    "xb_uniformity_backward": (_i32, [ctypes.POINTER(UniformityDesc), _vp, _vp, _vp, _sz, _vp]),
    "xb_topk_filter": (_i32, [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "xb_retrieval_metrics": (_i32, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "xb_topk_workspace_bytes": (_sz, [ctypes.POINTER(TopkDesc)]),
    "xb_mask_words": (_i32, [_i32]),
    "xb_topk_search": (_i32, [ctypes.POINTER(TopkDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xb_topk_merge": (_i32, [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "xb_pair_mask_workspace_bytes": (_sz, [_i32]),
    "xb_build_pair_mask": (_i32, [_i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xb_hash_indices": (_i32, [_vp, _i64, _i32, _u32, _i32, _vp, _vp]),
    "xb_hash_gather": (_i32, [_vp, _i64, _i32, _u32, _vp, _i32, _i32, _vp, _vp, _vp]),
    "xb_hash_scatter_grad": (_i32, [_vp, _i64, _i32, _u32, _vp, _i32, _i32, _vp, _vp]),
    "xb_debug_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "xb_debug_scores": (_i32, [_i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xb_debug_loss_region": (_i32, [ctypes.POINTER(LossDesc), _i32, ctypes.POINTER(_sz), ctypes.POINTER(_sz)]),
    "xb_debug_set_trace": (_i32, [_vp, _i32]),
    "xb_sweep_timing": (_i32, [_i32]),
    "xb_sweep_timing_read": (_i32, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
    "xb_last_error_string": (ctypes.c_char_p, []),
    "xb_version": (ctypes.c_char_p, []),
    "xb_launch_count": (_i64, [_i32]),
}


def _load() -> ctypes.CDLL:
    if not LIB_PATH.exists():
        msg = (
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C matrix-factorization-torch_b200/csrc -j8`). There is no fallback implementation."
        )
        raise ImportError(msg)
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the header and the library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


def check(status: int, what: str) -> None:
    if status != XB_OK:
        detail = lib.xb_last_error_string().decode()
        raise XbError(f"{what}: {ERROR_NAMES.get(status, status)}: {detail}")


def require_cuda(*tensors: torch.Tensor | None) -> torch.device:
    """All tensors must live on one CUDA device; returns it."""
    device = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            msg = "xfmr_b200 kernels run on CUDA (sm_100a) tensors only; there is no CPU path"
            raise RuntimeError(msg)
        if device is None:
            device = t.device
        elif t.device != device:
            msg = f"tensors on different devices: {device} vs {t.device}"
            raise RuntimeError(msg)
    if device is None:
        msg = "no tensor given"
        raise RuntimeError(msg)
    return device


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def dtype_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return XB_DTYPE_F32
    if dtype == torch.bfloat16:
        return XB_DTYPE_BF16
    msg = f"embeddings must be float32 or bfloat16, got {dtype}"
    raise TypeError(msg)


def compute_code(compute: str | None, dtype: torch.dtype) -> int:
    """``None`` picks the arithmetic from the input dtype: fp32 -> split-bf16 (fp32-grade), bf16 -> bf16."""
    if compute is None:
        return XB_COMPUTE_SPLIT if dtype == torch.float32 else XB_COMPUTE_BF16
    if compute in ("bf16", "bfloat16"):
        return XB_COMPUTE_BF16
    if compute in ("split", "fp32", "float32"):
        return XB_COMPUTE_SPLIT
    msg = f"compute must be None, 'bf16' or 'fp32', got {compute!r}"
    raise ValueError(msg)


def mining_code(mining: str) -> int:
    """``"semi_hard"`` (``semi_hard_mining``, what the reference losses call) or ``"hard"`` (``hard_mining``)."""
    if mining in ("semi_hard", "semi-hard"):
        return XB_MINING_SEMI_HARD
    if mining == "hard":
        return XB_MINING_HARD
    msg = f"mining must be 'semi_hard' or 'hard', got {mining!r}"
    raise ValueError(msg)


def sweep_timing(enable: bool) -> None:  # noqa: FBT001
    """Measurement hook: bracket every sweep launch with CUDA events (see ``xb_sweep_timing``)."""
    check(lib.xb_sweep_timing(1 if enable else 0), "xb_sweep_timing")


def sweep_timing_read() -> tuple[float, int]:
    """``(total sweep milliseconds, number of sweep launches)`` since ``sweep_timing(True)``; blocks the host."""
    total = ctypes.c_double(0.0)
    count = ctypes.c_int64(0)
    check(lib.xb_sweep_timing_read(ctypes.byref(total), ctypes.byref(count)), "xb_sweep_timing_read")
    return total.value, count.value


def launch_count(*, reset: bool = False) -> int:
    return int(lib.xb_launch_count(1 if reset else 0))
