"""sln_b200 — B200-native (sm_100a) implementation of the data-parallel hot path of aluo-x/3D_SLN.

``import sln_b200`` (the alias directory ``3d_sln_b200/`` maps ``importlib.import_module("3d_sln_b200...")`` onto the same
module objects).  Sub-modules mirror the reference layout:

    models.graph            <- reference models/graph.py          (GraphTripleConv, GraphTripleConvNet, make_mlp)
    models.Sg2ScVAE_model   <- reference models/Sg2ScVAE_model.py (Sg2ScVAEModel)
    utils                   <- reference utils.py hot-path pieces  (calculate_model_losses, tensor_aug) + FusedAdam, VAETrainStep
    data.synthetic          synthetic SUNCG-shaped scene graphs (format of data/suncg_dataset.py:295-337)

Everything computes in libsln_b200.so (hand-written CUDA, C ABI in include/sln_b200.h).  No CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
