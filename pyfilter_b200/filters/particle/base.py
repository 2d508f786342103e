"""``ParticleFilter`` (reference filters/particle/base.py:14-174 on top of filters/base.py:17-221): same constructor, same
methods; the per-step work runs in the fused CUDA kernels behind the C ABI instead of ~100-220 ATen calls."""
from typing import Callable, Union

import torch

from ... import resampling as _resampling
from ...timeseries import StateSpaceModel
from ..result import FilterResult
from .engine import Engine
from .proposals import Bootstrap, Proposal
from .state import ParticleFilterCorrection

_RESAMPLERS = {_resampling.systematic: 0, _resampling.multinomial: 1}


class ParticleFilter:
    algorithm_id = -1

    def __init__(self, model, particles: int, resampling: Callable = _resampling.systematic, proposal: Union[str, Proposal] = None,
                 ess_threshold=0.9, record_states=False, record_moments=True, nan_strategy: str = "skip",
                 record_intermediary_states: bool = False, seed: int = None, fold_lookahead: bool = True,
                 exact_weights: bool = False):
        if not (isinstance(model, StateSpaceModel) or callable(model)):
            raise ValueError("`model` must be a `StateSpaceModel` or a callable that returns one!")
        builder = callable(model) and not isinstance(model, StateSpaceModel)
        self._model_builder = model if builder else (lambda _: model)
        self._model = None if builder else model
        self._batch_shape = torch.Size([])
        self.record_states, self.record_moments = record_states, record_moments
        if nan_strategy not in ["skip", "impute"]:
            raise NotImplementedError(f"Currently cannot handle strategy '{nan_strategy}'!")
        self._nan_strategy, self._record_intermediary = nan_strategy, record_intermediary_states
        self._base_particles = torch.Size([int(particles)])
        self._ess_threshold_arg = ess_threshold
        self._resample_threshold = ess_threshold * particles  # filters/particle/base.py:42
        if resampling not in _RESAMPLERS:
            raise NotImplementedError("only pyfilter_b200.resampling.systematic / multinomial run inside the device loop")
        self._resampler = resampling
        self._proposal: Proposal = proposal if proposal is not None else Bootstrap()
        self._seed = seed
        self._fold = fold_lookahead
        self._engine: Engine = None
        self._exact_weights = bool(exact_weights)  # see include/smcb200.h: smcb_config.exact_weights

    # ---- reference surface
    @property
    def ssm(self) -> StateSpaceModel:
        return self._model

    @property
    def batch_shape(self) -> torch.Size:
        return self._batch_shape

    def set_batch_shape(self, batch_shape: torch.Size):
        if len(batch_shape) > 1:
            raise NotImplementedError("Currently do not support nested batches!")
        self._batch_shape = torch.Size(batch_shape)
        self._engine = None

    @property
    def particles(self) -> torch.Size:
        return torch.Size([*self._base_particles, *self.batch_shape])

    @property
    def proposal(self) -> Proposal:
        return self._proposal

    def increase_particles(self, factor: int):
        self._base_particles = torch.Size([int(factor * self._base_particles[0])])
        self._resample_threshold *= factor
        self._engine = None

    def initialize_model(self, context):
        self._model = self._model_builder(context)
        self._proposal.set_model(self._model)
        self._engine = None

    def copy(self):
        res = type(self)(model=self._model_builder, particles=self._base_particles[0], resampling=self._resampler,
                         proposal=self._proposal.copy(), ess_threshold=self._resample_threshold,  # sic: base.py:165 (Appendix A-12)
                         record_states=self.record_states, record_moments=self.record_moments, nan_strategy=self._nan_strategy,
                         record_intermediary_states=self._record_intermediary, seed=self._seed, fold_lookahead=self._fold, exact_weights=self._exact_weights)
        res._model = self._model
        res.set_batch_shape(self.batch_shape)
        return res

    def predict(self, state):
        raise NotImplementedError("predict/correct are fused into one device step; call filter() or batch_filter()")

    correct = predict

    def smooth(self, states, method="ffbs"):
        raise NotImplementedError("smoothing is listed under 'next' (SURVEY.md 8(f) f3)")

    # ---- engine management
    def _get_engine(self, history_rows: int) -> Engine:
        assert self._model is not None, "Model has not been initialized!"
        self._proposal.set_model(self._model)
        e = self._engine
        if e is None or e.history_rows < history_rows:
            seed = self._seed if self._seed is not None else int(torch.randint(0, 2**62, (1,)).item())
            n = int(self._base_particles[0])
            e = Engine(self._model, self._proposal.proposal_id, self.algorithm_id, _RESAMPLERS[self._resampler], n,
                       self.batch_shape, self._resample_threshold / n, seed, history_rows, self._fold, self._exact_weights)
            self._engine = e
        return e

    def _adopt(self, engine: Engine, state: ParticleFilterCorrection):
        if not state.is_live(engine):
            engine.load_state(state.timeseries_state.value, state.weights, state.previous_indices,
                              int(state.timeseries_state.time_index))

    def initialize(self) -> ParticleFilterCorrection:
        e = self._get_engine(2)
        e.initialize()
        return e.make_state()

    def initialize_with_result(self, state=None) -> FilterResult:
        return FilterResult(state or self.initialize(), self.record_states, self.record_moments)

    def batch_filter(self, y, bar=True, init_state=None) -> FilterResult:
        """``BaseFilter.batch_filter`` (filters/base.py:140-158): the whole time loop stays on the device."""
        y = torch.as_tensor(y)
        T = int(y.shape[0])
        e = self._get_engine(T + 1)
        if init_state is None:
            e.initialize()
            init_state = e.make_state()
        else:
            self._adopt(e, init_state)
        result = FilterResult(init_state, self.record_states, self.record_moments)
        y_dev = y.to(device="cuda", dtype=torch.float32).reshape(T, -1).contiguous()
        t0 = e.t
        e.set_observations(y_dev, t0)
        keep_all = not (self.record_states is False)
        if keep_all:  # every intermediate state is wanted: one move at a time
            for _ in range(T):
                e.run(1)
                result.append(e.make_state())
            return result
        if bar:
            from tqdm import tqdm

            chunk = max(1, T // 20)
            for start in tqdm(range(0, T, chunk), desc=type(self).__name__):
                e.run(min(chunk, T - start))
        else:
            e.run(T)
        hist_rows = min(e.history_rows, t0 + T + 1)
        means, variances, ll = e.history(hist_rows)
        lo = t0 + 1
        if lo + T <= hist_rows:
            result.extend_moments(means[lo: lo + T], variances[lo: lo + T])
            result._loglikelihood = result._loglikelihood + ll[lo: lo + T].sum(0)
        result._states.append(e.make_state())
        return result

    def filter(self, y, correction: ParticleFilterCorrection, result: FilterResult = None) -> ParticleFilterCorrection:
        """``BaseFilter.filter`` (filters/base.py:188-221): one move."""
        e = self._get_engine(2)
        self._adopt(e, correction)
        y_dev = torch.as_tensor(y).to(device="cuda", dtype=torch.float32).reshape(1, -1).contiguous()
        e.set_observations(y_dev, e.t)
        e.run(1)
        state = e.make_state()
        if result is not None:
            result.append(state)
        return state
