"""Import alias: ``import sln_b200`` == ``importlib.import_module("3d_sln_b200")`` (the package name is not an identifier)."""
import importlib
import sys

_pkg = importlib.import_module("3d_sln_b200")
sys.modules[__name__] = _pkg
