"""SMC2 (Chopin et al.) on the resident batch of filters - reference inference/sequential/smc2.py:53-65 (the step), its particle
Metropolis-Hastings rejuvenation kernel inference/sequential/kernels/mh.py:52-140 and inference/batch/mcmc/utils.py:14-77 (one PMMH
sweep), the symmetric proposal inference/batch/mcmc/proposals/symmetric_mh.py:13-23 with inference/utils.py:42-76, and the state
bookkeeping of inference/sequential/state.py:35-44.

What runs where: the theta-particles are the COLUMNS of one ``smcb_filter`` handle (``set_batch_shape``, filters/base.py:93-119); a
step is one fused move of all columns (csrc/column.cuh or csrc/move.cuh); theta-resampling and the accept/reject exchange permute /
copy columns of the resident state on the device (``smcb_filter_resample_columns`` / ``smcb_filter_exchange_columns``, csrc/plugin.cuh);
the proposal filter of a PMMH sweep is a second handle that re-filters the whole data with the candidate parameters
(``smcb_filter_set_params``).  The theta-level arithmetic itself - ``(B,)`` weights, the ``(B, p)`` parameter cloud, a ``p x p``
Cholesky factor - is torch on the device; the theta-level resampling and normalisation go through ``pyfilter_b200.resampling`` /
``pyfilter_b200.utils`` like the state level.  One host synchronisation per step (the ESS test), as in the reference (smc2.py:62)."""
import math
from typing import Callable, Dict, List, Optional

import torch

from .. import _lib, resampling as _resampling, utils as _utils
from ..filters.particle import APF
from ..filters.particle.engine import Engine
from .prior import ParameterContext, Prior


class TooManyIncreases(Exception):
    pass


class _Phases:
    """Wall time per phase of the algorithm (diagnostics: SMCB_SMC2_TIMING=1 makes every phase end with a device synchronisation)."""

    def __init__(self):
        import os

        self.on = bool(os.environ.get("SMCB_SMC2_TIMING"))
        self.t = {}

    def __call__(self, name):
        return _Phase(self, name)


class _Phase:
    def __init__(self, owner, name):
        self.o, self.n = owner, name

    def __enter__(self):
        if self.o.on:
            import time

            torch.cuda.synchronize()
            self.t0 = time.perf_counter()

    def __exit__(self, *a):
        if self.o.on:
            import time

            torch.cuda.synchronize()
            self.o.t[self.n] = self.o.t.get(self.n, 0.0) + time.perf_counter() - self.t0


# ---- theta-level pieces (each one callable on its own: tests/test_gpu_smc2.py compares them with oracle/smc2_oracle.py) -------------
def calc_mean_chol(x: torch.Tensor, w: torch.Tensor):
    """``inference/utils.py:42-58``: weighted mean and the Cholesky factor of the weighted covariance of the rows of ``x`` ``(B, p)``;
    a failed factorisation falls back to the diagonal."""
    mean = w @ x
    centralized = x - mean
    cov = (w * centralized.t()).matmul(centralized)
    chol, info = torch.linalg.cholesky_ex(cov)
    if bool((info > 0).any()):
        chol = cov.diag().sqrt().diag()
    return mean, chol


def construct_mvn(x: torch.Tensor, w: torch.Tensor, scale: float = 1.0):
    """``inference/utils.py:61-76``: ``(mean, scale_tril)`` of the multivariate normal fitted to the weighted samples."""
    mean, chol = calc_mean_chol(x, w)
    return mean, scale * chol


def mvn_sample(mean: torch.Tensor, scale_tril: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """``MultivariateNormal.sample``: ``mean + scale_tril @ eps`` for every row of ``eps`` ``(B, p)``."""
    return mean + eps @ scale_tril.t()


def mvn_log_prob(mean: torch.Tensor, scale_tril: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """``MultivariateNormal(mean, scale_tril=...).log_prob(x)`` for the rows of ``x``."""
    p = mean.shape[0]
    z = torch.linalg.solve_triangular(scale_tril, (x - mean).t(), upper=False)
    return -0.5 * (z * z).sum(0) - scale_tril.diagonal().log().sum() - 0.5 * p * math.log(2.0 * math.pi)


def pmmh_accept(diff_logl: torch.Tensor, diff_prior: torch.Tensor, diff_prop: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """``inference/batch/mcmc/utils.py:66-67``: accept where ``log u < diff_prop + diff_prior + diff_logl``."""
    return u.log() < (diff_prop + diff_prior + diff_logl)


class SMC2State:
    """``SMC2State`` (inference/sequential/state.py:10-95): theta log-weights, the ESS history, the parsed observations."""

    def __init__(self, weights: torch.Tensor, engine: Engine):
        self.w = weights          # (B,) theta log-weights of the WHOLE batch (replicated on every rank of a sharded run)
        self.engine = engine      # the handle that holds this process' columns
        self.ess: List[float] = [float(_utils.get_ess(weights))]
        self.parsed_data: List[torch.Tensor] = []
        self.current_iteration = 0
        self.rejuvenations = 0
        self.acceptance: List[float] = []

    @property
    def loglikelihood(self) -> torch.Tensor:
        return self.engine.raw(_lib.PTR_LL_TOTAL, (self.engine.B,))

    def normalized_weights(self) -> torch.Tensor:
        return _utils.normalize(self.w)


class SMC2:
    def __init__(self, model_builder: Callable[[Dict[str, torch.Tensor]], object], priors: Dict[str, Prior], particles: int,
                 state_particles: int, filter_cls=APF, proposal=None, threshold: float = 0.2, num_steps: int = 1,
                 acceptance_threshold: float = 0.2, max_increases: int = 5, seed: Optional[int] = None, max_observations: int = 1024,
                 resampling=_resampling.systematic):
        self._builder = model_builder
        self.context = ParameterContext(priors)
        self.particles = torch.Size([int(particles)])
        self._n_state = int(state_particles)
        self._filter_cls, self._proposal = filter_cls, proposal
        self._threshold = float(threshold)            # ConstantThreshold (inference/sequential/threshold.py)
        self._n_steps = int(num_steps)
        self._acceptance_threshold = float(acceptance_threshold)
        self._max_increases, self._increases = int(max_increases), 0
        self._seed = seed
        self._max_obs, self._oes = int(max_observations), 1
        self._rows = self._max_obs + 2
        self._resampler = resampling                    # kernels/base.py:15-23
        self._filter = None
        self._proposal_filter = None
        self._y_dev: torch.Tensor = None
        self._gen = torch.Generator().manual_seed(seed if seed is not None else int(torch.randint(0, 2**62, (1,)).item()))
        self.phases = _Phases()

    # ---- where the columns live: one process holds all of them here; ShardedSMC2 overrides the three hooks
    def _columns(self) -> slice:
        """The theta-particles whose filters this process holds."""
        return slice(0, int(self.particles[0]))

    def _gather(self, local: torch.Tensor) -> torch.Tensor:
        """A per-column quantity of this process' columns -> the whole batch's."""
        return local

    def _migrate(self, engine: Engine, indices: torch.Tensor):
        """theta-resampling of the filter states: column b <- column indices[b] of the whole batch (state.filter_state.resample)."""
        engine.resample_columns(indices, entire_history=True)

    def _local_model(self, context: ParameterContext):
        cols = self._columns()
        return self._builder({k: v[cols] for k, v in context.constrained().items()})

    # ---- filters
    def _make_filter(self, context: ParameterContext, n_state: int, salt: int):
        cols = self._columns()
        f = self._filter_cls(self._local_model(context), n_state, proposal=self._proposal.copy() if self._proposal is not None else None,
                             seed=None if self._seed is None else self._seed + 7919 * salt, column_offset=cols.start)
        f.set_batch_shape(torch.Size([cols.stop - cols.start]))
        # filters/base.py:204-210: observation k >= 1 is consumed by move k * observe_every_step, the moves in between only propagate
        self._oes = int(getattr(f.ssm, "observe_every_step", 1))
        self._rows = self._max_obs * self._oes + 2
        return f, f._get_engine(self._rows)

    def _moves(self, n_obs: int) -> int:
        """Moves the filter has made once ``n_obs`` observations are in."""
        return 0 if n_obs == 0 else (n_obs - 1) * self._oes + 1

    def _draw_seed(self) -> int:
        return int(torch.randint(0, 2**62, (1,), generator=self._gen).item())

    def initialize(self) -> SMC2State:
        """``SequentialParticleAlgorithm.initialize`` (inference/sequential/base.py:56-67) + ``SMC2.initialize`` (smc2.py:48-51)."""
        B = int(self.particles[0])
        self.context.initialize_parameters(B, self._gen)
        self._filter, e = self._make_filter(self.context, self._n_state, 0)
        e.initialize()
        self._y_dev = torch.full((self._rows, e.OD), float("nan"), device="cuda", dtype=torch.float32)
        return SMC2State(torch.zeros(B, device="cuda"), e)

    # ---- one observation (smc2.py:53-65)
    def step(self, y: torch.Tensor, state: SMC2State) -> SMC2State:
        e = state.engine
        t = len(state.parsed_data)
        if t + 1 > self._max_obs:
            raise ValueError("more observations than `max_observations`")
        yt = torch.as_tensor(y, dtype=torch.float32).reshape(-1)
        state.parsed_data.append(yt)
        done, upto = self._moves(t), self._moves(t + 1)
        self._y_dev[upto - 1] = yt.to("cuda")                            # the rows of the propagate-only moves stay NaN
        with self.phases("filter move"):
            e.set_observations(self._y_dev[:upto], 0)
            e.run(upto - done)
        with self.phases("gather increments"):
            state.w += self._gather(e.raw(_lib.PTR_LL, (e.B,)))      # SequentialAlgorithmState.append (state.py:35-44)
        with self.phases("ess"):   # the ESS and the finiteness flag come to the host together: ONE synchronisation per observation (smc2.py:59-62)
            ess, finite = torch.stack((_utils.get_ess(state.w).reshape(()), torch.isfinite(state.w).all().float())).tolist()
        state.ess.append(ess)
        any_nans = finite < 0.5
        if ess < self._threshold * self.particles[0] or any_nans:
            with self.phases("rejuvenate (all of it)"):
                state = self.rejuvenate(state)
        state.current_iteration += 1
        return state

    def fit(self, y: torch.Tensor) -> SMC2State:
        state = self.initialize()
        for yt in torch.as_tensor(y):
            state = self.step(yt, state)
        return state

    # ---- ParticleMetropolisHastings.update (kernels/mh.py:52-108)
    def rejuvenate(self, state: SMC2State) -> SMC2State:
        ctx, e = self.context, state.engine
        B = int(self.particles[0])
        with self.phases("theta resample + proposal fit"):
            W = state.normalized_weights()
        u = torch.rand(1, generator=self._gen)                            # the offset from the algorithm's own generator: every rank of a
        indices = self._resampler(W, normalized=True, u=u.cuda()) if self._resampler is _resampling.systematic else self._resampler(W, normalized=True)
        kernel = construct_mvn(ctx.stack_parameters(), W, scale=1.1)      # SymmetricMH.build (symmetric_mh.py:13-23)      sharded run draws the same
        ctx.resample(indices)
        with self.phases("migrate columns"):
            self._migrate(e, indices)                                     # state.filter_state.resample(indices)
        T = self._moves(len(state.parsed_data))
        if self._proposal_filter is None or self._proposal_filter[1].N != e.N:
            with self.phases("create proposal filter"):
                self._proposal_filter = self._make_filter(ctx, e.N, 1 + self._increases)
        sub_context = ctx.make_new()
        pe = self._proposal_filter[1]
        acceptance_rate = 0.0
        for i in range(self._n_steps):
            accepted = self._run_pmmh(ctx, state, kernel, pe, sub_context, T)
            acceptance_rate = (float(accepted.float().mean()) + i * acceptance_rate) / (i + 1)
            if acceptance_rate < self._acceptance_threshold:
                state.acceptance.append(acceptance_rate)
                return self._increase_states(state)
        state.acceptance.append(acceptance_rate)
        e.set_params(self._local_model(ctx))                              # filter_.initialize_model(context)
        state.w.fill_(0.0)
        state.rejuvenations += 1
        return state

    # ---- run_pmmh (inference/batch/mcmc/utils.py:14-77)
    def _run_pmmh(self, ctx: ParameterContext, state: SMC2State, kernel, pe: Engine, sub_context: ParameterContext, T: int) -> torch.Tensor:
        B = int(self.particles[0])
        mean, scale_tril = kernel
        with self.phases("pmmh propose"):
            eps = torch.randn(B, mean.shape[0], generator=self._gen).to("cuda")
            rvs = mvn_sample(mean, scale_tril, eps)
            sub_context.unstack_parameters(rvs)
        cols = self._columns()
        with self.phases("proposal filter run"):
            pe.set_params(self._local_model(sub_context))
            pe.set_seed(self._draw_seed())                               # a fresh random stream for every re-filtering
            pe.initialize()
            pe.set_observations(self._y_dev[:T], 0)
            pe.run(T)
        ph = self.phases("pmmh accept arithmetic + exchange")
        ph.__enter__()
        new_ll = pe.raw(_lib.PTR_LL_TOTAL, (pe.B,))
        diff_logl = new_ll - state.loglikelihood                         # this process' columns
        diff_prior = sub_context.eval_priors() - ctx.eval_priors()       # (B,) arithmetic, replicated
        uniform = torch.full((B,), 1.0 / B, device="cuda")               # state.replicate(new_res): zero log-weights
        new_kernel = construct_mvn(sub_context.stack_parameters(), uniform, scale=1.1)
        diff_prop = mvn_log_prob(*new_kernel, ctx.stack_parameters()) - mvn_log_prob(mean, scale_tril, rvs)
        u = torch.rand(B, generator=self._gen).to("cuda")
        accepted_local = pmmh_accept(diff_logl, diff_prior[cols], diff_prop[cols], u[cols])
        state.engine.exchange_columns(pe, accepted_local)                 # state.filter_state.exchange(new_res, accepted)
        accepted = self._gather(accepted_local.float()) > 0.5            # the one collective of a sweep on a sharded batch
        ctx.exchange(sub_context, accepted)
        ph.__exit__()
        return accepted

    # ---- _increase_states (kernels/mh.py:110-140): twice the state particles, re-filter, weights = change of the log-likelihoods
    def _increase_states(self, state: SMC2State) -> SMC2State:
        self._increases += 1
        if self._increases > self._max_increases:
            raise TooManyIncreases(f"Configuration only allows {self._max_increases}!")
        old_ll = state.loglikelihood.clone()
        T = self._moves(len(state.parsed_data))
        self._filter, e = self._make_filter(self.context, 2 * state.engine.N, 100 + self._increases)
        e.initialize()
        e.set_observations(self._y_dev[:T], 0)
        e.run(T)
        weight = self._gather(e.raw(_lib.PTR_LL_TOTAL, (e.B,)) - old_ll)
        res = SMC2State(weight.clone(), e)
        res.ess, res.parsed_data, res.current_iteration = state.ess, state.parsed_data, state.current_iteration
        res.rejuvenations, res.acceptance = state.rejuvenations, state.acceptance
        self._proposal_filter = None
        return res

    # ---- summaries
    def posterior_mean(self, state: SMC2State) -> Dict[str, torch.Tensor]:
        W = state.normalized_weights()
        return {k: (W * v).sum() for k, v in self.context.constrained().items()}


class ShardedSMC2(SMC2):
    """SMC2 with the theta-particles sharded over the ranks of a ``torch.distributed`` group (one process per GPU):
    every rank holds a block of the columns (``pyfilter_b200.sharding.column_shard``, Philox streams keyed by the GLOBAL column index),
    runs the filter moves and the proposal filters of its block, and keeps a replica of the theta-level state - the ``(B,)`` weights and
    the ``(B, p)`` parameter cloud - which it advances with the same arithmetic from the same generator (``seed`` is mandatory), so no
    rank ever has to be told what another one decided.  Communication: per step the ``(B_local,)`` likelihood increments (all-gather);
    per PMMH sweep the accepted flags; per rejuvenation the filter states, because the ancestor of a column may live on another GPU
    (export of the resident columns as records -> all-gather -> import of the records a rank now owns).  A sharded run reproduces the
    single-process run of the same seed bit for bit (tools/smc2_sharded_check.py)."""

    def __init__(self, *args, group=None, **kwargs):
        import torch.distributed as dist

        from ..sharding import column_shard

        if kwargs.get("seed") is None:
            raise ValueError("a sharded run needs `seed`: every rank replays the same theta-level random numbers")
        super().__init__(*args, **kwargs)
        self._dist, self._group = dist, group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self._group), dist.get_rank(self._group)
        self._lo, self._hi = column_shard(int(self.particles[0]), self.rank, self.world)
        if (self._hi - self._lo) * self.world != int(self.particles[0]):
            raise ValueError("the number of theta-particles must be a multiple of the world size")
        self._gather_bufs = {}

    def initialize(self) -> SMC2State:
        state = super().initialize()
        # the proposal filter and the migration buffers exist before the first rejuvenation (allocation is slow once the ranks have mapped
        # each other's memory, see _gather)
        self._proposal_filter = self._make_filter(self.context, self._n_state, 1)
        self._proposal_filter[1].initialize()
        self._gather(state.engine.export_columns())
        return state

    def _columns(self) -> slice:
        return slice(self._lo, self._hi)

    def _gather(self, local: torch.Tensor) -> torch.Tensor:
        # one output buffer per (shape, dtype), kept: with peer access enabled between the GPUs every NEW device allocation is mapped
        # into all peers and costs milliseconds - the loop must not allocate
        local = local.contiguous()
        key = (tuple(local.shape), local.dtype)
        out = self._gather_bufs.get(key)
        if out is None:
            out = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
            self._gather_bufs[key] = out
        self._dist.all_gather_into_tensor(out, local, group=self._group)
        return out.clone() if out.numel() <= 1 << 16 else out   # small results are handed out as copies (the buffer is reused)

    def _migrate(self, engine: Engine, indices: torch.Tensor):
        records = engine.export_columns()                    # (B_local, record)
        everything = self._gather(records)                   # (B, record): the ranks' blocks in column order
        engine.import_columns(everything, indices[self._lo:self._hi])
