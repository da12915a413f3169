"""The C-ABI library loads without a GPU and exports exactly what ``include/xfmr_b200.h`` declares."""

from __future__ import annotations

import ctypes
import pathlib
import re
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "xfmr_b200.h"


def declared_symbols() -> set[str]:
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(xb_[a-z0-9_]+)\s*\(", text))


def test_header_declares_entry_points() -> None:
    syms = declared_symbols()
    for required in ("xb_loss_forward", "xb_loss_backward", "xb_topk_search", "xb_topk_merge", "xb_hash_gather",
                     "xb_build_pair_mask", "xb_last_error_string"):
        assert required in syms


def test_library_exports_every_declared_symbol() -> None:
    import xfmr_b200  # noqa: PLC0415

    lib = ctypes.CDLL(str(xfmr_b200.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    # and the Python binding table covers the header, nothing more
    assert set(xfmr_b200._lib.SIGNATURES) == declared_symbols()  # noqa: SLF001


def test_exported_symbols_are_only_the_abi() -> None:
    import xfmr_b200  # noqa: PLC0415

    out = subprocess.run(["nm", "-D", "--defined-only", str(xfmr_b200.LIB_PATH)], capture_output=True, text=True, check=True)  # noqa: S603, S607
    exported = {line.split()[-1] for line in out.stdout.splitlines() if " T " in line}
    assert {s for s in exported if s.startswith("xb_")} == declared_symbols()


def test_sass_is_blackwell_native() -> None:
    """tcgen05.mma / tcgen05.ld / TMA must be present in the compiled kernels (UTCHMMA / LDTM / UTMALDG)."""
    import shutil  # noqa: PLC0415

    import xfmr_b200  # noqa: PLC0415

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", str(xfmr_b200.LIB_PATH)], capture_output=True, text=True, check=True).stdout  # noqa: S603, S607
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "sm_100a" in sass


def test_workspace_queries_need_no_gpu() -> None:
    import xfmr_b200  # noqa: PLC0415
    from xfmr_b200 import _lib  # noqa: PLC0415

    desc = _lib.LossDesc(batch=4096, num_items=87585, dim=128, num_pos=32, in_dtype=1, compute=0, num_negatives=0,
                         loss_mask=127, sigma=1.0, margin=1.0, has_log_q=0, mining=0)
    assert _lib.lib.xb_loss_workspace_bytes(ctypes.byref(desc)) > 0
    bad = _lib.LossDesc(batch=8, num_items=4, dim=128, num_pos=0, in_dtype=1, compute=0, num_negatives=0,
                        loss_mask=127, sigma=1.0, margin=1.0, has_log_q=0, mining=0)
    assert _lib.lib.xb_loss_workspace_bytes(ctypes.byref(bad)) == 0
    assert b"batch" in _lib.lib.xb_last_error_string()
    too_wide = _lib.LossDesc(batch=8, num_items=16, dim=512, num_pos=0, in_dtype=1, compute=0, num_negatives=0,
                             loss_mask=127, sigma=1.0, margin=1.0, has_log_q=0, mining=0)
    assert _lib.lib.xb_loss_workspace_bytes(ctypes.byref(too_wide)) == 0
    assert xfmr_b200._lib.lib.xb_mask_words(3706) == 4 * 29  # noqa: SLF001


def test_header_is_plain_c() -> None:
    """The boundary is a C ABI: the header must compile as C99 on its own (no C++ or torch types)."""
    import shutil  # noqa: PLC0415
    import tempfile  # noqa: PLC0415

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = ('#include "xfmr_b200.h"\n'
           "int main(void) { xb_loss_desc d; xb_topk_desc t; xb_uniformity_desc u; d.mining = XB_MINING_HARD; t.k = 1; u.n = 2;\n"
           "  (void)d; (void)t; (void)u; return XB_OK; }\n")
    with tempfile.NamedTemporaryFile("w", suffix=".c", delete=False) as f:
        f.write(src)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", str(HEADER.parent), f.name],  # noqa: S603, S607
                   check=True)
