"""``sige.nn`` -> ``sige_b200.nn`` (reference sige/nn/__init__.py:1-4)."""
from sige_b200.nn import (  # noqa: F401
    Gather, Scatter, ScatterGather, ScatterWithBlockResidual, SIGEConv2d, SIGEModel, SIGEModule, SIGEModuleWrapper,
    activation, utils,
)
