"""``ParticleFilter`` (reference filters/particle/base.py:14-174 on top of filters/base.py:17-221): same constructor, same
methods; the per-step work runs in the fused CUDA kernels behind the C ABI instead of ~100-220 ATen calls."""
from typing import Callable, Union

import torch

from ... import resampling as _resampling
from ...timeseries import StateSpaceModel
from ..result import FilterResult
from .engine import Engine
from .proposals import Bootstrap, Proposal
from .state import ParticleFilterCorrection, ParticleFilterPrediction

_RESAMPLERS = {_resampling.systematic: 0, _resampling.multinomial: 1}


class ParticleFilter:
    algorithm_id = -1

    def __init__(self, model, particles: int, resampling: Callable = _resampling.systematic, proposal: Union[str, Proposal] = None,
                 ess_threshold=0.9, record_states=False, record_moments=True, nan_strategy: str = "skip",
                 record_intermediary_states: bool = False, seed: int = None, fold_lookahead: bool = True,
                 exact_weights: bool = False, column_offset: int = 0):
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
        # systematic / multinomial run inside the fused move.  Any other resampler - `residual`, a user's callable
        # `(weights, normalized=False) -> indices` (filters/particle/base.py:23,43) - is served by the SPLIT step: `predict` calls it
        # between the stand-alone device passes, `correct` is one fused move with the resampling rule switched off.  SISR only: the
        # APF resamples inside its correction.
        self._split = resampling not in _RESAMPLERS
        if self._split and not callable(resampling):
            raise ValueError("`resampling` must be a callable")
        if self._split and self.algorithm_id == 1:
            raise NotImplementedError("APF resamples inside the fused move: pyfilter_b200.resampling.systematic / multinomial only")
        self._resampler = resampling
        self._proposal: Proposal = proposal if proposal is not None else Bootstrap()
        if getattr(self._proposal, "gaussian", False) and self.algorithm_id != 2:
            raise NotImplementedError("GaussianProposal is the proposal of GPF")
        self._seed = seed
        self._fold = fold_lookahead
        self._engine: Engine = None
        self._exact_weights = bool(exact_weights)  # see include/smcb200.h: smcb_config.exact_weights
        # Random streams: the Philox key is `seed`, the counter (particle group, column_offset + column, move, purpose).  Shards of ONE
        # batch of filters (ranks of a theta-sharded SMC2 run, pyfilter_b200.sharding) pass the global index of their first column as
        # `column_offset` with the SAME seed: no two columns share a stream and the result does not depend on the split.  Every new
        # engine of a filter (history growth, set_batch_shape, increase_particles) and every copy() derives a fresh sub-seed.
        self._column_offset = int(column_offset)
        self._generation = 0   # engines created so far
        self._copies = 0

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
                         record_intermediary_states=self._record_intermediary, seed=self._derive_seed(0x9E3779B9 + self._copies),
                         fold_lookahead=self._fold, exact_weights=self._exact_weights, column_offset=self._column_offset)
        self._copies += 1
        res._model = self._model
        res.set_batch_shape(self.batch_shape)
        return res

    # ---- the split step of the reference (filters/base.py:160-186): predict, then correct.  `filter()` / `batch_filter()` run both in ONE
    #      fused kernel; these two exist for code written against the reference's plug-in surface and are built from the stand-alone
    #      device passes (normalize, get_ess, the resamplers, batched_gather) and a single fused move from a loaded state.
    def predict(self, state: ParticleFilterCorrection) -> ParticleFilterPrediction:
        raise NotImplementedError()

    def correct(self, y: torch.Tensor, prediction: ParticleFilterPrediction) -> ParticleFilterCorrection:
        raise NotImplementedError()

    def _move_from(self, y, x: torch.Tensor, weights: torch.Tensor, indices: torch.Tensor, t: int, resample: bool) -> ParticleFilterCorrection:
        """One fused move from an explicit state.  ``resample=False`` (SISR.correct: the prediction is already resampled) switches the
        ESS rule off for this move."""
        e = self._get_engine(2)
        n = int(self._base_particles[0])
        if not resample:
            e.set_ess_threshold(-1.0)
        try:
            e.load_state(x, weights, indices, int(t))
            e.set_observations(torch.as_tensor(y, dtype=torch.float32).reshape(1, -1).to("cuda").contiguous(), e.t)
            e.run(1)
            state = e.make_state().detach_copy()   # not a live view: a later filter() call reloads it and refreshes the statistics
        finally:
            e.set_ess_threshold(self._resample_threshold / n)
        return state

    def _filter_split(self, y, correction: ParticleFilterCorrection, result: FilterResult = None) -> ParticleFilterCorrection:
        """``BaseFilter.filter`` as the reference writes it (filters/base.py:201-221): predict, the propagate-only moves of
        ``observe_every_step``, then ``correct`` - or another propagate-only move when the observation is all NaN."""
        y = torch.as_tensor(y, dtype=torch.float32)
        k = int(getattr(self._model, "observe_every_step", 1))
        nan_y = torch.full_like(y.reshape(-1), float("nan"))

        def propagate_only(pred):   # ParticleFilterPrediction.create_state_from_prediction (particle/state.py:38-42)
            x = pred.get_timeseries_state()
            return self._move_from(nan_y, x.value, pred.weights, pred.indices, int(x.time_index), resample=False)

        prediction = self.predict(correction)
        while int(prediction.get_timeseries_state().time_index) % k != 0:
            correction = propagate_only(prediction)
            if result is not None and self._record_intermediary:
                result.append(correction)
            prediction = self.predict(correction)
        correction = propagate_only(prediction) if bool(torch.isnan(y).all()) else self.correct(y, prediction)
        if result is not None:
            result.append(correction)
        return correction

    def _do_sample_fl(self, states):
        """Fixed-lag smoothing = ancestral tracing over the recorded states (filters/particle/base.py:130-146)."""
        from ..utils import trace_back

        rev = list(reversed(list(states)))
        latest = rev[0]
        result = [latest.timeseries_state.value]
        n = int(self._base_particles[0])
        lineage = torch.arange(n, device=result[0].device)
        if self.batch_shape:
            lineage = lineage.unsqueeze(-1).expand(self.particles).contiguous()
        for s in rev[1:]:
            x_s, lineage = trace_back(s.timeseries_state.value, lineage, latest.previous_indices)
            result.append(x_s)
            latest = s
        return torch.stack(result[::-1], dim=0)

    def _do_sample_ffbs(self, states):
        """Forward filtering - backward sampling as the reference writes it (filters/particle/base.py:105-128)."""
        from .smoothing import ffbs

        return ffbs(self, states)

    def smooth(self, states, method="ffbs") -> torch.Tensor:
        lower_method = method.lower()
        if lower_method == "ffbs":
            return self._do_sample_ffbs(states)
        if method == "fl":
            return self._do_sample_fl(states)
        raise NotImplementedError(f"Currently do not support '{method}'!")

    # ---- engine management
    def _derive_seed(self, salt: int) -> int:
        """SplitMix64 of (seed, salt): sub-seeds of copies and of re-created engines (None stays None: drawn from torch's generator)."""
        if self._seed is None:
            return None
        z = (int(self._seed) + (salt + 1) * 0x9E3779B97F4A7C15) & (2**64 - 1)
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
        return (z ^ (z >> 31)) & (2**62 - 1)

    def _get_engine(self, history_rows: int) -> Engine:
        assert self._model is not None, "Model has not been initialized!"
        self._proposal.set_model(self._model)
        e = self._engine
        if e is None or e.history_rows < history_rows:
            if self._seed is None:
                seed = int(torch.randint(0, 2**62, (1,)).item())
            else:  # the first engine uses the seed itself (two filters built with the same seed agree), later ones a sub-seed
                seed = self._seed if self._generation == 0 else self._derive_seed(self._generation)
            self._generation += 1
            n = int(self._base_particles[0])
            e = Engine(self._model, self._proposal.proposal_id, self.algorithm_id, _RESAMPLERS.get(self._resampler, 0), n,
                       self.batch_shape, self._resample_threshold / n, seed, history_rows, self._fold, self._exact_weights,
                       self._column_offset, proposal_config=self._proposal.config())
            self._engine = e
        return e

    def _adopt(self, engine: Engine, state: ParticleFilterCorrection):
        if not state.is_live(engine):
            state._check_live("particles and weights")   # a stale view of this engine's buffers cannot be adopted: it raises
            engine.load_state(state.timeseries_state.value, state.weights, state.previous_indices,
                              int(state.timeseries_state.time_index))

    def initialize(self) -> ParticleFilterCorrection:
        e = self._get_engine(2)
        e.initialize()
        return e.make_state()

    def initialize_with_result(self, state=None) -> FilterResult:
        return FilterResult(state or self.initialize(), self.record_states, self.record_moments)

    def _expand_observations(self, y: torch.Tensor, t0: int):
        """``observe_every_step`` (filters/base.py:204-210): before an observation is used the filter makes propagate-only moves
        until the time index is a multiple of ``observe_every_step``.  On the device a propagate-only move is a move whose
        observation is NaN (filters/base.py:213-214 takes the same route), so the observations are interleaved with NaN rows.
        Returns the expanded ``(moves, obs_dim)`` tensor and the (0-based) move index of every real observation."""
        k = int(getattr(self._model, "observe_every_step", 1))
        T = int(y.shape[0])
        y2 = y.to(dtype=torch.float32).reshape(T, -1)
        if k == 1:
            return y2, list(range(T))
        rows, observed, t = [], [], int(t0)
        nan_row = torch.full((1, y2.shape[1]), float("nan"))
        for i in range(T):
            while t % k != 0:
                rows.append(nan_row)
                t += 1
            observed.append(len(rows))
            rows.append(y2[i: i + 1].cpu())
            t += 1
        return torch.cat(rows, 0), observed

    def batch_filter(self, y, bar=True, init_state=None) -> FilterResult:
        """``BaseFilter.batch_filter`` (filters/base.py:140-158): the whole time loop stays on the device."""
        y = torch.as_tensor(y)
        if self._split:   # a resampler the fused loop does not hold: the reference's own loop over filter() (filters/base.py:140-158)
            state = init_state if init_state is not None else self.initialize().detach_copy()
            result = FilterResult(state, self.record_states, self.record_moments)
            for yt in y:
                state = self.filter(yt, state, result=result)
            return result
        T = int(y.shape[0])
        k = int(getattr(self._model, "observe_every_step", 1))
        t_start = 0 if init_state is None else int(init_state.timeseries_state.time_index)
        e = self._get_engine(t_start + T * k + 1)
        if init_state is None:
            e.initialize()
            init_state = e.make_state()
        else:
            self._adopt(e, init_state)
        result = FilterResult(init_state, self.record_states, self.record_moments)
        t0 = e.t
        y_moves, observed = self._expand_observations(y, t0)
        M = int(y_moves.shape[0])  # moves, propagate-only ones included
        if t0 + M + 1 > e.history_rows:
            raise ValueError("the moment history of the handle is too short for this call; create the filter anew")
        y_dev = y_moves.to(device="cuda", dtype=torch.float32).contiguous()
        e.set_observations(y_dev, t0)
        keep_all = not (self.record_states is False)
        if keep_all:  # every recorded state is wanted: one move at a time
            obs = set(observed)
            for mv in range(M):
                e.run(1)
                if mv in obs or self._record_intermediary:
                    result.append(e.make_state())
            return result
        if bar:
            from tqdm import tqdm

            chunk = max(1, M // 20)
            for start in tqdm(range(0, M, chunk), desc=type(self).__name__):
                e.run(min(chunk, M - start))
        else:
            e.run(M)
        means, variances, ll = e.history(t0 + M + 1)
        # history row of a move = its index + 1; with record_intermediary_states the propagate-only moves of observe_every_step are
        # recorded as well (filters/base.py:207-208; their likelihood increments are zero)
        kept = range(M) if self._record_intermediary else observed
        rows = torch.as_tensor([t0 + 1 + mv for mv in kept], device=means.device)
        result.extend_moments(means[rows], variances[rows])
        result._loglikelihood = result._loglikelihood + ll[rows].sum(0)
        result._states.append(e.make_state())
        return result

    def filter(self, y, correction: ParticleFilterCorrection, result: FilterResult = None) -> ParticleFilterCorrection:
        """``BaseFilter.filter`` (filters/base.py:188-221): one observation - the propagate-only moves ``observe_every_step`` asks
        for, then the observed move."""
        if self._split:
            return self._filter_split(y, correction, result)
        e = self._get_engine(2)
        self._adopt(e, correction)
        y_moves, _ = self._expand_observations(torch.as_tensor(y).reshape(1, -1), e.t)
        M = int(y_moves.shape[0])
        e.set_observations(y_moves.to(device="cuda", dtype=torch.float32).contiguous(), e.t)
        if result is not None and self._record_intermediary and M > 1:
            for _ in range(M - 1):
                e.run(1)
                result.append(e.make_state())
            e.run(1)
        else:
            e.run(M)
        state = e.make_state()
        if result is not None:
            result.append(state)
        return state
