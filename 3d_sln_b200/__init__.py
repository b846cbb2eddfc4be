"""Alias directory: the package lives in ``sln_b200/`` (an importable identifier).  ``importlib.import_module("3d_sln_b200")`` and
``importlib.import_module("3d_sln_b200.models.graph")`` keep working and return the SAME module objects as ``import sln_b200...``
(no second copy of any class, one load of libsln_b200.so)."""
import importlib
import importlib.abc
import importlib.machinery
import sys

_ALIAS, _REAL = __name__, "sln_b200"


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    _specs = {}

    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(_ALIAS + "."):
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        real = importlib.import_module(_REAL + spec.name[len(_ALIAS):])
        self._specs[spec.name] = real.__spec__
        return real

    def exec_module(self, module):
        # the import machinery re-stamped __spec__ with the alias spec: put the real one back
        real = self._specs.pop(module.__spec__.name, None)
        if real is not None:
            module.__spec__ = real


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
sys.modules[_ALIAS] = importlib.import_module(_REAL)
