"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference modules from /root/reference when that tree exists.

/root/reference is only present in the build container (not on the GPU box), so everything here is optional: callers
must handle ``available() == False`` (tests skip).  Used to pin the oracle restatements and to generate tests/golden/*.
"""
import importlib
import os
import sys

REF_ROOT = os.environ.get("SLN_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "Sg2ScVAE_model.py"))


def _import(name):
    if not available():
        raise ImportError("reference tree not present at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # the reference's top-level names (models, utils, data) are generic: import under a clean module cache entry
    return importlib.import_module(name)


def vae_model_class():
    return _import("models.Sg2ScVAE_model").Sg2ScVAEModel


def graph_module():
    return _import("models.graph")


def reference_losses():
    return _import("utils").calculate_model_losses


def spade_module():
    return _import("models.SPADE_related")
