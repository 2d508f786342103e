"""Stand-in for ``stochproc.timeseries`` (see package docstring)."""
from . import result  # noqa: F401
from .state import TimeseriesState
from .process import (
    StructuralStochasticProcess,
    AffineProcess,
    AffineEulerMaruyama,
    StateSpaceModel,
    LinearStateSpaceModel,
)
from . import models  # noqa: F401

__all__ = [
    "TimeseriesState",
    "StructuralStochasticProcess",
    "AffineProcess",
    "AffineEulerMaruyama",
    "StateSpaceModel",
    "LinearStateSpaceModel",
    "models",
    "result",
]
