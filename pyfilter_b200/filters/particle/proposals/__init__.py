"""Proposal plug-ins (reference filters/particle/proposals/): on the device a proposal is an enum baked into the fused
step kernel, so these classes only carry the choice and the compatibility check of the reference."""
from ....timeseries import StateSpaceModel


class Proposal:
    proposal_id = -1

    def __init__(self, pre_weight_func=None):
        if pre_weight_func is not None:
            raise NotImplementedError("custom pre-weight callables cannot run inside the fused kernel")
        self._model = None

    def set_model(self, model: StateSpaceModel):
        self._model = model
        return self

    def copy(self) -> "Proposal":
        return type(self)()


class Bootstrap(Proposal):
    """``proposals/bootstrap.py:4-17``: propose from the dynamics, weight by the observation density."""

    proposal_id = 0


class LinearGaussianObservations(Proposal):
    """``proposals/linear.py:13-89``: the optimal Gaussian kernel for ``y = b + a x + s nu`` (scalar state in the zoo)."""

    proposal_id = 1

    def set_model(self, model):
        if not getattr(model, "is_linear_gaussian", False):
            raise ValueError("Model combination not supported!")  # proposals/linear.py:32-36
        return super().set_model(model)


def _out_of_scope(name):
    class _Missing(Proposal):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is outside the B200 hot path (SURVEY.md section 2 row 5)")

    _Missing.__name__ = name
    return _Missing


Linearized = _out_of_scope("Linearized")
NestedProposal = _out_of_scope("NestedProposal")
GaussianLinear = _out_of_scope("GaussianLinear")
GaussianLinearized = _out_of_scope("GaussianLinearized")
GaussianProposal = _out_of_scope("GaussianProposal")
LocalLinearization = _out_of_scope("LocalLinearization")
