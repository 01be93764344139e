"""cobweb-b200: B200-native engine for the Cobweb concept-tree hot path.

Drop-in for the reference's ``src.cobweb`` classes::

    from rag_cobweb_b200 import CobwebTorchTree, CobwebWrapper

Importing the package never touches CUDA; constructing a tree or wrapper requires a CUDA
device and the in-tree libcobweb_b200.so (there is no CPU fallback).
"""
from . import constants, evaluate, serialize, synth, topology  # noqa: F401
from ._lib import CobwebB200Error  # noqa: F401
from .tree import CobwebNode, CobwebTorchTree, default_prior_var  # noqa: F401
from .whitening import PCAICAWhiteningModel  # noqa: F401
from .wrapper import CobwebWrapper, DenseIndex  # noqa: F401

__all__ = ["CobwebTorchTree", "CobwebWrapper", "CobwebNode", "DenseIndex", "PCAICAWhiteningModel", "CobwebB200Error", "default_prior_var",
           "synth", "topology", "serialize", "evaluate", "constants"]
