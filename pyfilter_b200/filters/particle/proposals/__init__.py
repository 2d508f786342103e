"""Proposal plug-ins (reference filters/particle/proposals/).  Inside the fused move a proposal is an enum baked into the kernel; the
plug-in METHODS of the reference - ``pre_weight`` and ``sample_and_weight`` (proposals/base.py:52-85) - are stand-alone device passes
of the same device functions (``smcb_filter_pre_weight`` / ``smcb_filter_sample_and_weight``, csrc/plugin.cuh), so code written
against the reference's split step (``filter.predict`` / ``filter.correct``, a custom APF loop) runs unchanged."""
from typing import Tuple

import torch

from ....timeseries import StateSpaceModel, TimeseriesState


class Proposal:
    proposal_id = -1

    def __init__(self, pre_weight_func=None):
        if pre_weight_func is not None:
            raise NotImplementedError("custom pre-weight callables cannot run inside the compiled kernels")
        self._model = None
        self._op_engine = None

    def set_model(self, model: StateSpaceModel):
        if model is not self._model:
            self._op_engine = None
        self._model = model
        return self

    def copy(self) -> "Proposal":
        return type(self)()

    def config(self) -> dict:
        """Settings the compiled proposal reads (smcb_config)."""
        return {}

    # ---- the plug-in methods
    def _engine_for(self, x: torch.Tensor):
        """A handle of the right shape for the stand-alone passes (kept until the shape or the model changes)."""
        from ..engine import Engine

        assert self._model is not None, "call set_model first"
        d = len(self._model.hidden.event_shape)
        n = int(x.shape[0])
        batch = torch.Size(x.shape[1: x.dim() - d])
        e = self._op_engine
        if e is None or e.N != n or e.batch_shape != batch:
            e = Engine(self._model, self.proposal_id, 0, 0, n, batch, 0.9, int(torch.randint(0, 2**62, (1,)).item()), 1,
                       proposal_config=self.config())
            self._op_engine = e
        return e

    def pre_weight(self, y: torch.Tensor, x: TimeseriesState) -> torch.Tensor:
        """``proposals/base.py:69-85`` (Bootstrap: ``log p(y | loc(x))`` with the affine pre-weight function, pre_weight_funcs.py:9-11)
        / ``proposals/linear.py:57-86``: the log-weights the APF uses to select the particles it propagates."""
        return self._engine_for(x.value).pre_weight(y, x.value)

    def sample_and_weight(self, y: torch.Tensor, prediction, eps: torch.Tensor = None) -> Tuple[TimeseriesState, torch.Tensor]:
        """``proposals/bootstrap.py:10-14`` / ``proposals/linear.py:38-55``: new particles and their weight increments.  ``eps``
        (same shape as the particles) injects the N(0, 1) draws - the parity hook; otherwise they come from Philox."""
        x = prediction.get_timeseries_state()
        e = self._engine_for(x.value)
        t = int(x.time_index)
        x_new, w = e.sample_and_weight(y, x.value, eps, t)
        return x.propagate_from(x_new), w


class Bootstrap(Proposal):
    """``proposals/bootstrap.py:4-17``: propose from the dynamics, weight by the observation density."""

    proposal_id = 0


class LinearGaussianObservations(Proposal):
    """``proposals/linear.py:13-89``: the optimal Gaussian kernel for ``y = b + A x + s nu`` - scalar state, and the vector state of the
    Lorenz model (examples/lorenz.ipynb:214), whose matrices are diagonal."""

    proposal_id = 1

    def set_model(self, model):
        if not getattr(model, "is_linear_gaussian", False):
            raise ValueError("Model combination not supported!")  # proposals/linear.py:32-36
        return super().set_model(model)


class Linearized(Proposal):
    """``proposals/linearized.py:9-73``: a Gaussian kernel around a (few-step) ascent of ``log p(y | x) + log p(x | x_prev)`` from the
    transition mean, first order (``x += alpha`` per step, as the reference's ``ModeFinder.find_mode`` does it) or with second-order
    information.  The derivatives are the closed forms of the compiled models (``Model::obs_grad_hess``, csrc/models.h) instead of
    functorch's; the legacy ``torch.autograd`` path (``use_functorch=False``) is not built."""

    proposal_id = 2

    def __init__(self, n_steps=1, alpha: float = 1e-4, use_second_order: bool = False, use_functorch: bool = True):
        assert n_steps > 0, "``n_steps`` must be >= 1"
        super().__init__()
        if not use_functorch:
            raise NotImplementedError("the legacy autograd path of ModeFinder (find_mode_legacy) is not compiled")
        self._alpha, self._n_steps, self._use_second_order = alpha, n_steps, use_second_order

    def set_model(self, model):
        if getattr(model, "model_id", None) == 4:   # a user model brings no derivatives
            raise ValueError("Hidden must be of type AffineProcess with compiled derivatives!")
        return super().set_model(model)

    def config(self) -> dict:
        return {"n_steps": self._n_steps, "alpha": self._alpha, "use_second_order": self._use_second_order}

    def copy(self) -> "Proposal":
        return Linearized(self._n_steps, self._alpha, self._use_second_order)


class NestedProposal(Proposal):
    """``proposals/nested.py:8-50`` (Naesseth et al.): ``num_samples`` draws from the transition density per particle, one of them -
    chosen with probabilities proportional to their observation densities - becomes the particle, the weight is the log of their mean
    observation density.  Compiled into the fused kernels (csrc/step.cuh); at most 256 inner samples."""

    proposal_id = 3

    def __init__(self, num_samples: int, **kwargs):
        super().__init__(**kwargs)
        self._num_samples = torch.Size([int(num_samples)])

    def set_model(self, model):
        if getattr(model, "model_id", None) == 4:
            raise ValueError("NestedProposal is compiled for the models of the zoo")
        return super().set_model(model)

    def config(self) -> dict:
        return {"num_samples": int(self._num_samples[0])}

    def copy(self) -> "Proposal":
        return NestedProposal(self._num_samples[0])


class GaussianProposal(Proposal):
    """``proposals/approximate.py:13-37``: the proposal of the Gaussian particle filter - every particle is a draw from the Gaussian fitted
    to the propagated, weighted cloud, weighted by the observation density.  It is the fit over the WHOLE cloud that makes it a step of
    the filter rather than a per-particle pass: it runs inside ``GPF``'s move (csrc/gpf.cuh) and has no stand-alone form."""

    proposal_id = 0
    gaussian = True

    def pre_weight(self, y, x):
        raise NotImplementedError("GaussianProposal belongs to GPF, which does not pre-weight")

    def sample_and_weight(self, y, prediction, eps=None):
        raise NotImplementedError("GaussianProposal runs inside GPF's move: use GPF.filter / GPF.correct")


def _out_of_scope(name):
    class _Missing(Proposal):
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is outside the B200 hot path (SURVEY.md section 2 row 5)")

    _Missing.__name__ = name
    return _Missing


GaussianLinear = _out_of_scope("GaussianLinear")
GaussianLinearized = _out_of_scope("GaussianLinearized")
LocalLinearization = _out_of_scope("LocalLinearization")
