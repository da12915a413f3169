"""Import shim: ``import xfmr_b200`` loads the package in ``matrix-factorization-torch_b200/``
(a directory name with hyphens cannot appear in an ``import`` statement)."""

import importlib
import pathlib
import sys

_root = str(pathlib.Path(__file__).resolve().parent)
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("matrix-factorization-torch_b200")
sys.modules[__name__] = _pkg
