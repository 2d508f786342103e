"""NESS (Crisan & Miguez) and its fixed-width variant on the resident batch of filters - reference inference/sequential/ness.py:16-109,
the online kernel inference/sequential/kernels/online.py:26-53 and the jittering kernels inference/sequential/kernels/jittering.py:11-225.

A step is: (when the theta-level ESS has fallen) resample the theta-particles, jitter the parameters, permute the resident filter
columns on the device (``smcb_filter_resample_columns``, latest state only: ``entire_history=False``), hand the filters their new
parameters (``smcb_filter_set_params``) - then one fused move of all columns.  The theta-level arithmetic is ``(B, p)`` torch on the
device; the state particles never leave it."""
import math
from typing import Callable, Dict, Optional

import torch

from .. import _lib, resampling as _resampling, utils as _utils
from ..filters.particle import APF
from .prior import ParameterContext, Prior
from .smc2 import SMC2State

EPS = math.sqrt(torch.finfo(torch.float32).eps)   # constants.py: EPS


def robust_var(x: torch.Tensor, w: torch.Tensor, mean: torch.Tensor = None) -> torch.Tensor:
    """``jittering.py:49-83``: ``min(IQR / 1.349, sigma)^2`` per parameter, the quartiles read off the weighted empirical distribution."""
    srt, sort_indices = x.sort(0)
    ws = w[sort_indices]                                            # (B, p)
    B = x.shape[0]
    tri = torch.ones(B, B, device=x.device, dtype=x.dtype).tril()   # running sums of the sorted weights as one (B, B) product
    cumulative_weights = tri @ ws
    low_indices = (cumulative_weights - 0.25).abs().argmin(0)
    high_indices = (cumulative_weights - 0.75).abs().argmin(0)
    cols = torch.arange(x.shape[1], device=x.device)
    iqr = (srt[high_indices, cols] - srt[low_indices, cols]) / 1.349
    iqr2 = iqr**2
    wu = w.unsqueeze(-1)
    if mean is None:
        mean = (wu * x).sum(0)
    var = (wu * (x - mean) ** 2).sum(0)
    return torch.where(iqr2 <= var, iqr2, var)


class JitterKernel:
    """``jittering.py:86-137``."""

    def __init__(self, std_threshold: float = EPS):
        self._min_std = std_threshold

    def fit(self, x, w, indices):
        raise NotImplementedError()

    def jitter(self, x: torch.Tensor, w: torch.Tensor, indices: torch.Tensor, eps: torch.Tensor = None) -> torch.Tensor:
        mean, scale = self.fit(x, w, indices)
        std = scale.clamp(min=self._min_std) if isinstance(scale, torch.Tensor) else max(scale, self._min_std)
        if eps is None:
            eps = torch.randn(mean.shape, device=mean.device)
        return mean + std * eps                                     # _jitter (jittering.py:11-22)

    @staticmethod
    def get_ess(w):
        return 1.0 / (w * w).sum()


class ShrinkingKernel(JitterKernel):
    """``jittering.py:140-158`` (Flury & Shephard)."""

    def fit(self, x, w, indices):
        ess = self.get_ess(w)
        bw_fac = (1.59 * ess ** (-1 / 3)).clamp(EPS, 1 - EPS)
        mean = (w.unsqueeze(-1) * x).sum(0)
        var = robust_var(x, w, mean)
        beta = (1.0 - bw_fac**2).sqrt()
        return (mean + beta * (x - mean))[indices], bw_fac * var.sqrt()


class NonShrinkingKernel(ShrinkingKernel):
    """``jittering.py:161-173``."""

    def fit(self, x, w, indices):
        ess = self.get_ess(w)
        bw_fac = (1.59 * ess ** (-1 / 3)).clamp(EPS, 1 - EPS)
        return x[indices], bw_fac * robust_var(x, w).sqrt()


class LiuWestShrinkage(ShrinkingKernel):
    """``jittering.py:176-203``."""

    def __init__(self, a=0.98):
        super().__init__()
        self._a = a
        self._bw_fac = math.sqrt(1 - a**2)

    def fit(self, x, w, indices):
        mean = (w.unsqueeze(-1) * x).sum(0)
        var = robust_var(x, w, mean)
        return (x * self._a + (1 - self._a) * mean)[indices], self._bw_fac * var.sqrt()


class ConstantKernel(ShrinkingKernel):
    """``jittering.py:206-225``."""

    def __init__(self, scale):
        super().__init__()
        self._scale = scale

    def fit(self, x, w, indices):
        return x[indices], self._scale


class BaseOnlineAlgorithm:
    """``BaseOnlineAlgorithm`` (ness.py:16-56) with ``OnlineKernel.update`` (kernels/online.py:26-53)."""

    def __init__(self, model_builder: Callable[[Dict[str, torch.Tensor]], object], priors: Dict[str, Prior], particles: int,
                 state_particles: int, filter_cls=APF, proposal=None, kernel: JitterKernel = None, discrete: bool = False,
                 seed: Optional[int] = None, max_observations: int = 1024, resampling=_resampling.systematic):
        self._builder = model_builder
        self.context = ParameterContext(priors)
        self.particles = torch.Size([int(particles)])
        self._n_state = int(state_particles)
        self._filter_cls, self._proposal = filter_cls, proposal
        self._kernel = kernel or NonShrinkingKernel()
        self._disc = bool(discrete)
        self._seed = seed
        self._max_obs, self._oes = int(max_observations), 1
        self._rows = self._max_obs + 2
        self._resampler = resampling
        self._gen = torch.Generator().manual_seed(seed if seed is not None else int(torch.randint(0, 2**62, (1,)).item()))
        self._y_dev = None
        self.updates = 0

    def initialize(self) -> SMC2State:
        B = int(self.particles[0])
        self.context.initialize_parameters(B, self._gen)
        f = self._filter_cls(self._builder(self.context.constrained()), self._n_state,
                             proposal=self._proposal.copy() if self._proposal is not None else None, seed=self._seed)
        f.set_batch_shape(self.particles)
        self._filter = f
        # filters/base.py:204-210: observation k >= 1 is consumed by move k * observe_every_step, the moves in between only propagate
        self._oes = int(getattr(f.ssm, "observe_every_step", 1))
        self._rows = self._max_obs * self._oes + 2
        e = f._get_engine(self._rows)
        e.initialize()
        self._y_dev = torch.full((self._rows, e.OD), float("nan"), device="cuda", dtype=torch.float32)
        return SMC2State(torch.zeros(B, device="cuda"), e)

    def do_update_particles(self, state: SMC2State) -> bool:
        raise NotImplementedError()

    # OnlineKernel.update (kernels/online.py:26-53)
    def _update(self, state: SMC2State) -> SMC2State:
        ctx, e = self.context, state.engine
        weights = state.normalized_weights()
        stacked = ctx.stack_parameters()
        indices = self._resampler(weights, normalized=True)
        eps = torch.randn(stacked.shape, generator=self._gen).to("cuda")
        jittered = self._kernel.jitter(stacked, weights, indices, eps)
        ctx.resample(indices)
        e.resample_columns(indices, entire_history=False)
        if self._disc:
            p = 1.0 / weights.shape[0] ** 0.5
            to_jitter = (torch.rand(jittered.shape[0], generator=self._gen) < p).float().to("cuda").unsqueeze(-1)
            jittered = (1 - to_jitter) * stacked[indices] + to_jitter * jittered
        ctx.unstack_parameters(jittered)
        e.set_params(self._builder(ctx.constrained()))
        state.w.fill_(0.0)
        self.updates += 1
        return state

    def step(self, y: torch.Tensor, state: SMC2State) -> SMC2State:
        """``BaseOnlineAlgorithm._step`` (ness.py:50-56): update the particles when due, then one filter move."""
        if self.do_update_particles(state):
            state = self._update(state)
        e = state.engine
        t = len(state.parsed_data)
        if t + 1 > self._max_obs:
            raise ValueError("more observations than `max_observations`")
        yt = torch.as_tensor(y, dtype=torch.float32).reshape(-1)
        state.parsed_data.append(yt)
        done = 0 if t == 0 else (t - 1) * self._oes + 1
        upto = t * self._oes + 1
        self._y_dev[upto - 1] = yt.to("cuda")                            # the rows of the propagate-only moves stay NaN
        e.set_observations(self._y_dev[:upto], 0)
        e.run(upto - done)
        state.w += e.raw(_lib.PTR_LL, (e.B,))
        ess, finite = torch.stack((_utils.get_ess(state.w).reshape(()), torch.isfinite(state.w).all().float())).tolist()   # one synchronisation
        state.ess.append(ess)
        state.finite = finite > 0.5
        state.current_iteration += 1
        return state

    def fit(self, y: torch.Tensor) -> SMC2State:
        state = self.initialize()
        for yt in torch.as_tensor(y):
            state = self.step(yt, state)
        return state

    def posterior_mean(self, state: SMC2State) -> Dict[str, torch.Tensor]:
        W = state.normalized_weights()
        return {k: (W * v).sum() for k, v in self.context.constrained().items()}


class NESS(BaseOnlineAlgorithm):
    """``NESS`` (ness.py:59-83): particles are updated when the relative ESS falls below ``threshold`` (or a weight is not finite)."""

    def __init__(self, *args, threshold: float = 0.9, **kwargs):
        super().__init__(*args, **kwargs)
        self._threshold = threshold * int(self.particles[0])

    def do_update_particles(self, state):
        ess = state.ess
        return (any(ess) and ess[-1] < self._threshold) or not getattr(state, "finite", True)


class FixedWidthNESS(BaseOnlineAlgorithm):
    """``FixedWidthNESS`` (ness.py:86-109): particles are updated every ``block_len`` observations."""

    def __init__(self, *args, block_len: int = 125, **kwargs):
        super().__init__(*args, **kwargs)
        self._bl = int(block_len)
        self._num_iterations = 0

    def do_update_particles(self, state):
        self._num_iterations += 1
        return (self._num_iterations % self._bl == 0) or not getattr(state, "finite", True)
