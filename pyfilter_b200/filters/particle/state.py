"""``ParticleFilterCorrection`` (reference filters/particle/state.py:72-211): particles, log-weights, ancestors, log-likelihood
increment and moments of one filter move.  Big tensors are zero-copy views of the engine's device buffers while the state is
the live one; ``detach_copy`` turns them into owned tensors."""
from collections import OrderedDict
from typing import Any, Dict

import torch

from ...timeseries import TimeseriesState
from ...utils import normalize


class ParticleFilterPrediction:
    """Prediction state of the particle filters (reference filters/particle/state.py:14-69): the (resampled) previous particles,
    their log-weights and normalised weights, and the resampling indices."""

    def __init__(self, prev_x: TimeseriesState, weights: torch.Tensor, normalized_weights: torch.Tensor, indices: torch.Tensor):
        self.prev_x = prev_x
        self.weights = weights
        self.normalized_weights = normalized_weights
        self.indices = indices

    def get_timeseries_state(self) -> TimeseriesState:
        return self.prev_x


class SampledPath:
    """What ``model.sample_states`` returns in stochproc, reduced to the accessor the reference's callers use (``get_paths``)."""

    def __init__(self, x: torch.Tensor, y: torch.Tensor):
        self._x, self._y = x, y

    def get_paths(self):
        return self._x, self._y


class ParticleFilterCorrection(dict):
    def __init__(self, x: TimeseriesState, w: torch.Tensor, ll: torch.Tensor, prev_inds, mean: torch.Tensor, var: torch.Tensor,
                 engine=None, stamp=None):
        super().__init__()
        self["_x"], self["_w"], self["_ll"], self["_mean"], self["_var"] = x, w, ll, mean, var
        # same key as the reference (particle/state.py:155,162); an int64 tensor, or - while the state is the engine's live one - a
        # zero-argument callable that widens the device's int32 ancestors on first use
        dict.__setitem__(self, "_prev_inds", prev_inds)
        self._engine, self._stamp = engine, stamp

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if key == "_prev_inds" and callable(v):
            self._check_live("previous_indices")
            v = v()
            dict.__setitem__(self, key, v)
        return v

    def _check_live(self, what):
        e = self._engine
        if e is not None and self._stamp != e.stamp:
            raise RuntimeError(f"this state's {what} are views of device buffers that later filter moves have overwritten; keep a state "
                               "with .detach_copy() (FilterResult does so when record_states is set) before the filter moves again")

    # -- reference accessors
    @property
    def timeseries_state(self) -> TimeseriesState:
        self._check_live("particles")
        return self["_x"]

    def get_timeseries_state(self) -> TimeseriesState:
        return self["_x"]

    @property
    def weights(self) -> torch.Tensor:
        self._check_live("log-weights")
        return self["_w"]

    @property
    def previous_indices(self) -> torch.Tensor:
        return self["_prev_inds"]

    def get_loglikelihood(self):
        return self["_ll"]

    def get_mean(self):
        return self["_mean"]

    def get_variance(self):
        return self["_var"]

    def normalized_weights(self) -> torch.Tensor:
        return normalize(self.weights)

    # -- engine bookkeeping
    def is_live(self, engine) -> bool:
        return self._engine is engine and engine is not None and self._stamp == engine.stamp

    def detach_copy(self) -> "ParticleFilterCorrection":
        self._check_live("particles and weights")
        x = self.timeseries_state
        return ParticleFilterCorrection(TimeseriesState(x.time_index.clone(), x.value.clone(), x.event_shape), self.weights.clone(),
                                        self["_ll"].clone(), self.previous_indices.clone(), self["_mean"].clone(),
                                        self["_var"].clone())

    # -- theta-level operations used by SMC2 / PMMH (state.py:150-168)
    def resample(self, indices):
        self["_x"] = self.timeseries_state.copy(values=self.timeseries_state.value[:, indices])
        self["_w"] = self.weights[:, indices]
        self["_prev_inds"] = self.previous_indices[:, indices]
        self["_mean"] = self["_mean"][indices]
        self["_var"] = self["_var"][indices]
        self._engine = None

    def exchange(self, other, mask):
        if self._engine is not None:  # never write through a view into live engine buffers
            mine = self.detach_copy()
            for k in ("_x", "_w", "_ll", "_mean", "_var"):
                self[k] = mine[k]
            self["_prev_inds"], self._engine = mine.previous_indices, None
        self["_x"].value[:, mask] = other.timeseries_state.value[:, mask]
        self["_w"][:, mask] = other.weights[:, mask]
        self["_ll"][mask] = other.get_loglikelihood()[mask]
        self.previous_indices[:, mask] = other.previous_indices[:, mask]
        self["_mean"][mask] = other["_mean"][mask]
        self["_var"][mask] = other["_var"][mask]

    def predict_path(self, model, num_steps: int) -> SampledPath:
        """``particle/state.py:173-174``: ``model.sample_states(num_steps, x_0=self.timeseries_state)`` - every particle simulated
        ``num_steps`` transitions ahead together with its observations (one device pass, csrc/plugin.cuh)."""
        from .engine import Engine

        x = self.timeseries_state.value
        e = self._engine
        if e is None or e.model is not model:
            n = x.shape[0]
            d = len(model.hidden.event_shape)
            batch = tuple(x.shape[1: x.dim() - d])
            e = Engine(model, 0, 0, 0, n, torch.Size(batch), 0.9, int(torch.randint(0, 2**62, (1,)).item()), 1)
            e.t = int(self.timeseries_state.time_index)
        xp, yp = e.predict_path(int(num_steps), x)
        return SampledPath(xp, yp)

    def state_dict(self) -> Dict[str, Any]:
        res = OrderedDict()
        res["_w"], res["_ll"], res["_mean"], res["_var"] = self["_w"], self["_ll"], self["_mean"], self["_var"]
        res["_prev_inds"] = self.previous_indices
        res["_x"] = {"time_index": self.timeseries_state.time_index, "value": self.timeseries_state.value}
        return res

    def load_state_dict(self, state_dict: Dict[str, Any]):
        values = state_dict["_x"]["value"]
        assert self.timeseries_state.value.shape == values.shape, "Seems like you're loading a different shape"
        self["_x"] = TimeseriesState(state_dict["_x"]["time_index"], values, self.timeseries_state.event_shape)
        self["_w"], self["_ll"] = state_dict["_w"], state_dict["_ll"]
        self["_prev_inds"] = state_dict["_prev_inds"]
        self["_mean"], self["_var"] = state_dict["_mean"], state_dict["_var"]
        self._engine = None

    def __repr__(self):
        return f"ParticleFilterCorrection(time_index: {self.timeseries_state.time_index}, event_shape: {self.timeseries_state.event_shape})"
