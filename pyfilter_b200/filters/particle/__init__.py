from . import proposals  # noqa: F401
from .base import ParticleFilter
from .state import ParticleFilterCorrection, ParticleFilterPrediction  # noqa: F401

import torch

from ... import utils as _utils
from ..utils import batched_gather


class SISR(ParticleFilter):
    """Sequential Importance Sampling Resampling (reference filters/particle/sisr.py:7-56)."""

    algorithm_id = 0

    def predict(self, state):
        """``sisr.py:14-48``: columns whose ESS fell below the threshold are resampled (weights reset), the others pass through."""
        normalized = state.normalized_weights()
        ess = _utils.get_ess(normalized, normalized=True)
        mask = ess < self._resample_threshold
        ts_state = state.get_timeseries_state()
        weights, prev_inds = state.weights, state.previous_indices
        if not bool(mask.any()):
            return ParticleFilterPrediction(ts_state, weights, normalized, indices=prev_inds)
        n = weights.shape[0]
        if weights.dim() == 1:
            indices = self._resampler(normalized, normalized=True)
            x = batched_gather(ts_state.value, indices)
            return ParticleFilterPrediction(ts_state.copy(values=x), torch.zeros_like(weights), torch.full_like(normalized, 1.0 / n), indices)
        cols = mask.nonzero().flatten()
        sub_indices = self._resampler(normalized[:, cols].contiguous(), normalized=True)
        resampled_indices = prev_inds.clone()
        resampled_indices[:, cols] = sub_indices
        gather_idx = torch.arange(n, device=weights.device).unsqueeze(-1).expand(weights.shape).clone()
        gather_idx[:, cols] = sub_indices
        x = batched_gather(ts_state.value, gather_idx)
        um = mask.unsqueeze(0)
        return ParticleFilterPrediction(ts_state.copy(values=x), weights.masked_fill(um, 0.0), normalized.masked_fill(um, 1.0 / n),
                                        indices=resampled_indices)

    def correct(self, y, prediction):
        """``sisr.py:50-56``: propagate the (already resampled) prediction, add the weight increment, likelihood increment against the
        prediction's normalised weights - one fused move with the ESS rule switched off."""
        x = prediction.get_timeseries_state()
        return self._move_from(y, x.value, prediction.weights, prediction.indices, int(x.time_index), resample=False)


class APF(ParticleFilter):
    """Auxiliary Particle Filter of Pitt and Shephard (reference filters/particle/apf.py:9-46)."""

    algorithm_id = 1

    def predict(self, state):
        """``apf.py:16-23``: nothing moves yet - normalised weights and identity indices."""
        normalized = state.normalized_weights()
        old_indices = torch.arange(normalized.shape[0], device=normalized.device)
        if self.batch_shape:
            old_indices = old_indices.unsqueeze(-1).expand(self.particles)
        return ParticleFilterPrediction(state.timeseries_state, state.weights, normalized, old_indices)

    def correct(self, y, prediction):
        """``apf.py:25-46``: pre-weight, resample on ``g + log w``, propagate, second-stage weights - one fused move."""
        x = prediction.get_timeseries_state()
        return self._move_from(y, x.value, prediction.weights, prediction.indices, int(x.time_index), resample=True)


class GPF(ParticleFilter):
    """Gaussian particle filter of Kotecha and Djuric (reference filters/particle/gpf.py:10-36) with its default ``GaussianProposal``: per
    move the cloud is propagated, a Gaussian is fitted to it under the previous weights, every particle is redrawn from that Gaussian and
    weighted by the observation density (the weights are replaced); nothing resamples.  One difference from the reference: the time index
    of the state advances with every move (the reference's ``GaussianProposal`` copies the previous state's index, approximate.py:28)."""

    algorithm_id = 2

    def __init__(self, model, particles: int, proposal=None, **kwargs):
        proposal = proposal if proposal is not None else proposals.GaussianProposal()
        if not getattr(proposal, "gaussian", False):
            raise NotImplementedError("GPF is compiled with its default GaussianProposal (GaussianLinearized / GaussianLinear are not built)")
        super().__init__(model, particles, proposal=proposal, **kwargs)

    def predict(self, state):
        """``gpf.py:27-30``."""
        return ParticleFilterPrediction(state.timeseries_state, state.weights, state.normalized_weights(), state.previous_indices)

    def correct(self, y, prediction):
        """``gpf.py:32-36``: one move of csrc/gpf.cuh from the prediction's state."""
        x = prediction.get_timeseries_state()
        return self._move_from(y, x.value, prediction.weights, prediction.indices, int(x.time_index), resample=True)
