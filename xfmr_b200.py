"""Import shim: ``import xfmr_b200`` loads the package in ``matrix-factorization-torch_b200/``
(a directory name with hyphens cannot appear in an ``import`` statement)."""

import importlib
import importlib.abc
import importlib.util
import pathlib
import sys

_root = str(pathlib.Path(__file__).resolve().parent)
if _root not in sys.path:
    sys.path.insert(0, _root)
_REAL = "matrix-factorization-torch_b200"
_ALIAS = __name__


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """``xfmr_b200.<sub>`` is the module ``matrix-factorization-torch_b200.<sub>`` - the same object, never a second copy.

    A second execution of ``losses.py`` would re-define the torch custom ops, which destroys the library the first
    definition lives in and leaves the classes of the first import with a dangling operator.
    """

    def find_spec(self, fullname, path=None, target=None):  # noqa: ANN001, ANN201, ARG002
        if fullname.startswith(_ALIAS + "."):
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):  # noqa: ANN001, ANN201
        return importlib.import_module(_REAL + spec.name[len(_ALIAS):])

    def exec_module(self, module) -> None:  # noqa: ANN001
        pass


sys.meta_path.insert(0, _AliasFinder())
_pkg = importlib.import_module(_REAL)
sys.modules[_ALIAS] = _pkg
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_REAL + "."):
        sys.modules[_ALIAS + _name[len(_REAL):]] = _mod
