"""Operator surface — the names the reference exports from ``sige.nn``
(reference sige/nn/__init__.py:1-4)."""
from . import modules as utils  # ``sige.nn.utils.activation`` (reference sige/nn/utils.py)
from .modules import Gather, Scatter, ScatterGather, ScatterWithBlockResidual, SIGEConv2d, activation
from .state import SIGEModel, SIGEModule, SIGEModuleWrapper

__all__ = [
    "SIGEConv2d", "SIGEModel", "SIGEModule", "SIGEModuleWrapper", "Gather", "Scatter", "ScatterWithBlockResidual",
    "ScatterGather", "activation", "utils",
]
