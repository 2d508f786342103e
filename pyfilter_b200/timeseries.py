"""Model objects for the compiled zoo - the stand-in for the ``stochproc.timeseries`` objects a pyfilter user builds.

In the reference the state-space model is a bundle of Python callables (``mean_scale``, ``build_observation``; SURVEY.md
Appendix C) evaluated by torch.  A fused CUDA kernel cannot call Python, so ``pyfilter_b200`` ships the four BASELINE.json models as
device functions (``csrc/models.h``) and this module describes WHICH one to run and with what parameters.  The class and
attribute names follow stochproc so that user code reads the same; a model outside the zoo raises ``NotImplementedError``
(there is no torch/CPU fallback).

Parameters are float32 scalars or ``(B,)`` tensors - one value per parallel filter, exactly like the reference's batched
parameters under ``set_batch_shape``.
"""
import math
from typing import Sequence, Tuple

import torch

MODEL_LG_AR1, MODEL_SINE_EM, MODEL_SV_AR1, MODEL_LORENZ63_EM, MODEL_USER = 0, 1, 2, 3, 4


def _t(v) -> torch.Tensor:
    return v.detach().float().cpu() if isinstance(v, torch.Tensor) else torch.as_tensor(v, dtype=torch.float32)


class TimeseriesState(dict):
    """``stochproc.timeseries.TimeseriesState``: values + time index + event shape."""

    def __init__(self, time_index, values: torch.Tensor, event_shape: torch.Size):
        super().__init__()
        self.time_index = time_index if isinstance(time_index, torch.Tensor) else torch.tensor(time_index)
        self.value = values
        self.event_shape = event_shape

    @property
    def batch_shape(self):
        return self.value.shape[: self.value.dim() - len(self.event_shape)]

    def copy(self, values):
        return TimeseriesState(self.time_index, values, self.event_shape)

    def propagate_from(self, values, time_increment=1):
        return TimeseriesState(self.time_index + time_increment, values, self.event_shape)


class HiddenProcess:
    """A latent process of the zoo (``AffineProcess`` in stochproc terms): ``x_t = loc(x) + scale(x) * inc``."""

    name = "hidden"
    event_shape = torch.Size([])

    def __init__(self, parameters: Sequence):
        self.parameters = tuple(_t(p) for p in parameters)

    @property
    def n_dim(self) -> int:
        return len(self.event_shape)

    # --- torch-CPU evaluation, used only to SIMULATE synthetic data (never by the filters)
    def mean_scale(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    inc_scale = 1.0

    def initial_loc_scale(self):
        raise NotImplementedError


class AR(HiddenProcess):
    """``stochproc.timeseries.models.AR(alpha, beta, sigma)`` (reference tests/filters/models.py:12-14)."""

    name = "ar1"

    def __init__(self, alpha, beta, sigma):
        super().__init__((alpha, beta, sigma))

    def mean_scale(self, x):
        a, b, s = self.parameters
        return a + b * x, s

    def initial_loc_scale(self):
        a, b, s = self.parameters
        return a, s / (1.0 - b**2.0).sqrt()


class SineDiffusion(HiddenProcess):
    """Euler-Maruyama discretisation of ``dx = sin(x - gamma) dt + sigma dW`` (reference README.md:44-62: an
    ``AffineEulerMaruyama`` built from ``f(x, gamma, sigma) = (sin(x - gamma), sigma)`` with ``N(0, sqrt(dt))`` increments)."""

    name = "sine_em"

    def __init__(self, gamma, sigma, dt: float = 0.1):
        super().__init__((gamma, sigma))
        self.dt = float(dt)
        self.inc_scale = math.sqrt(self.dt)

    def mean_scale(self, x):
        g, s = self.parameters
        return x + torch.sin(x - g) * self.dt, s

    def initial_loc_scale(self):
        g, _ = self.parameters
        return torch.zeros_like(g), torch.ones_like(g)


class StochasticVolatilityAR(HiddenProcess):
    """Log-volatility AR(1): ``x_t = mu + phi (x_{t-1} - mu) + sigma_v eps`` with the stationary initial law."""

    name = "sv_ar1"

    def __init__(self, mu, phi, sigma_v):
        super().__init__((mu, phi, sigma_v))

    def mean_scale(self, x):
        m, p, s = self.parameters
        return m + p * (x - m), s

    def initial_loc_scale(self):
        m, p, s = self.parameters
        return m, s / (1.0 - p**2.0).sqrt()


class Lorenz63(HiddenProcess):
    """Euler-Maruyama Lorenz-63 (reference examples/lorenz.ipynb:53-107)."""

    name = "lorenz63_em"
    event_shape = torch.Size([3])

    def __init__(self, s, r, b, sigma=1.0, dt: float = 0.01):
        super().__init__((s, r, b, sigma))
        self.dt = float(dt)
        self.inc_scale = math.sqrt(self.dt)

    def mean_scale(self, x):
        s, r, b, sig = self.parameters
        f0 = -s * (x[..., 0] - x[..., 1])
        f1 = r * x[..., 0] - x[..., 1] - x[..., 0] * x[..., 2]
        f2 = x[..., 0] * x[..., 1] - b * x[..., 2]
        return x + torch.stack((f0, f1, f2), -1) * self.dt, sig

    def initial_loc_scale(self):
        return torch.tensor([-5.91652, -5.52332, 24.5723]), math.sqrt(10.0) * torch.ones(3)


class models:  # namespace like stochproc.timeseries.models
    AR = AR
    SineDiffusion = SineDiffusion
    StochasticVolatilityAR = StochasticVolatilityAR
    Lorenz63 = Lorenz63


class StateSpaceModel:
    """A (hidden process, observation) pair from the compiled zoo.  Mirrors the attributes the reference filters read:
    ``hidden``, ``parameters``, ``event_shape``, ``n_dim``, ``observe_every_step``."""

    def __init__(self, hidden: HiddenProcess, model_id: int, obs_parameters: Sequence, event_shape=torch.Size([]),
                 observe_every_step: int = 1, linear: bool = False):
        self.hidden = hidden
        self.model_id = model_id
        self.parameters = tuple(_t(p) for p in obs_parameters)
        self.event_shape = torch.Size(event_shape)
        self.observe_every_step = int(observe_every_step)  # filters/base.py:204-210: propagate-only moves between observations
        self.is_linear_gaussian = linear
        if self.observe_every_step < 1:
            raise ValueError("observe_every_step must be >= 1")

    @property
    def n_dim(self) -> int:
        return len(self.event_shape)

    # ---- what the C ABI needs
    def raw_parameters(self) -> Tuple[torch.Tensor, ...]:
        h, o = self.hidden.parameters, self.parameters
        if self.model_id == MODEL_LG_AR1:
            return (*h, *o)
        if self.model_id == MODEL_SINE_EM:
            return (h[0], h[1], _t(self.hidden.dt), *o)
        if self.model_id == MODEL_SV_AR1:
            return tuple(h)
        if self.model_id == MODEL_LORENZ63_EM:
            return (*h, _t(self.hidden.dt), *o)
        if self.model_id == MODEL_USER:
            return tuple(h)
        raise NotImplementedError

    def parameter_matrix(self, batch: int) -> torch.Tensor:
        """``(n_raw, cols)`` float32 host matrix, ``cols`` = 1 when every parameter is shared, else ``batch``."""
        raw = self.raw_parameters()
        cols = 1
        for p in raw:
            if p.dim() > 1 or (p.dim() == 1 and p.shape[0] not in (1, batch)):
                raise ValueError(f"parameter of shape {tuple(p.shape)} does not match the batch shape ({batch},)")
            if p.dim() == 1 and p.shape[0] == batch and batch > 1:
                cols = batch
        return torch.stack([p.reshape(-1).expand(cols) if p.numel() == 1 else p.reshape(-1) for p in raw]).contiguous()

    # ---- torch-CPU simulation of synthetic data (plumbing; not used by the filters)
    def obs_loc_scale(self, x: torch.Tensor):
        if self.model_id in (MODEL_LG_AR1, MODEL_SINE_EM):
            a, b, s = self.parameters
            return b + a * x, s
        if self.model_id == MODEL_SV_AR1:
            return torch.zeros_like(x), (x / 2.0).exp()
        a, s = self.parameters
        return torch.stack((a * x[..., 0], a * x[..., 2]), -1), s

    def sample_states(self, steps: int, generator: torch.Generator = None):
        """Draws ``(x_{1:T}, y_{1:T})`` on the CPU for ONE path (parameters must be scalars)."""
        h = self.hidden
        loc, scale = h.initial_loc_scale()
        x = loc + scale * torch.empty(loc.shape).normal_(generator=generator)
        xs, ys = [], []
        for _ in range(steps):
            m, s = h.mean_scale(x)
            x = m + s * (torch.empty(x.shape).normal_(generator=generator) * h.inc_scale)
            ol, os_ = self.obs_loc_scale(x)
            ys.append(ol + os_ * torch.empty(ol.shape).normal_(generator=generator))
            xs.append(x)
        return torch.stack(xs), torch.stack(ys)


class LinearStateSpaceModel(StateSpaceModel):
    """``stochproc.timeseries.LinearStateSpaceModel(hidden, (a, s) | (a, b, s), event_shape)``: ``y = b + a x + s nu``.
    Both parameter tuples occur in the reference (tests/filters/models.py:16, proposals/linear.py:48)."""

    def __init__(self, hidden: HiddenProcess, parameters: Sequence, event_shape=torch.Size([]), observe_every_step: int = 1):
        if isinstance(hidden, AR):
            mid = MODEL_LG_AR1
        elif isinstance(hidden, SineDiffusion):
            mid = MODEL_SINE_EM
        elif isinstance(hidden, Lorenz63):
            mid = MODEL_LORENZ63_EM
        else:
            raise NotImplementedError(f"{type(hidden).__name__} with linear-Gaussian observations is not in the compiled model zoo")
        if mid == MODEL_LORENZ63_EM:
            a, s = parameters
            super().__init__(hidden, mid, (a, s), torch.Size([2]), observe_every_step, linear=True)   # y = a (x1, x3) + s nu (lorenz.ipynb:105-117)
            return
        if len(event_shape) != 0:
            raise NotImplementedError("only scalar linear-Gaussian observations are in the compiled zoo")
        if len(parameters) == 2:
            a, s = parameters
            parameters = (a, 0.0, s)
        super().__init__(hidden, mid, parameters, event_shape, observe_every_step, linear=True)


class StochasticVolatilityModel(StateSpaceModel):
    """``y_t ~ N(0, exp(x_t / 2))`` on top of :class:`StochasticVolatilityAR` (BASELINE.json config 3)."""

    def __init__(self, hidden: StochasticVolatilityAR, observe_every_step: int = 1):
        if not isinstance(hidden, StochasticVolatilityAR):
            raise NotImplementedError("the exp(x/2) observation is compiled for StochasticVolatilityAR only")
        super().__init__(hidden, MODEL_SV_AR1, (), torch.Size([]), observe_every_step)


def build(name: str, **params) -> StateSpaceModel:
    """Factory keyed by the zoo names used in tests/golden and bench.py (``observe_every_step=k`` is passed through)."""
    every = int(params.pop("observe_every_step", 1))
    if name == "lg_ar1":
        p = dict(alpha=0.0, beta=0.99, sigma=0.05, a=1.0, b=0.0, s=0.15); p.update(params)
        return LinearStateSpaceModel(AR(p["alpha"], p["beta"], p["sigma"]), (p["a"], p["b"], p["s"]), observe_every_step=every)
    if name == "sine_em":
        p = dict(gamma=0.0, sigma=1.0, dt=0.1, a=1.0, b=0.0, s=0.1); p.update(params)
        return LinearStateSpaceModel(SineDiffusion(p["gamma"], p["sigma"], p["dt"]), (p["a"], p["b"], p["s"]), observe_every_step=every)
    if name == "sv_ar1":
        p = dict(mu=-1.0, phi=0.97, sigma_v=0.2); p.update(params)
        return StochasticVolatilityModel(StochasticVolatilityAR(p["mu"], p["phi"], p["sigma_v"]), observe_every_step=every)
    if name == "lorenz63_em":
        p = dict(s=10.0, r=28.0, b=8.0 / 3.0, sigma=1.0, dt=0.01, obs_a=0.8, obs_s=math.sqrt(0.1)); p.update(params)
        return LinearStateSpaceModel(Lorenz63(p["s"], p["r"], p["b"], p["sigma"], p["dt"]), (p["obs_a"], p["obs_s"]), torch.Size([2]),
                                     observe_every_step=every)
    raise NotImplementedError(f"'{name}' is not in the compiled model zoo")


# ---- user-supplied models ---------------------------------------------------------------------------------------
class UserProcess(HiddenProcess):
    name = "user"

    def __init__(self, parameters: Sequence, state_dim: int):
        super().__init__(parameters)
        self.event_shape = torch.Size([state_dim]) if state_dim > 1 else torch.Size([])


class UserStateSpaceModel(StateSpaceModel):
    """A model whose ``mean_scale`` / observation density the USER wrote as CUDA device functions (``compile_user_model``).  It runs through
    a build of the library that carries that code; everything else - filters, proposals' Bootstrap path, resamplers, SMC2 / NESS - is the
    same.  There is no torch evaluation of a user model: ``sample_states`` is not available."""

    def __init__(self, library_path: str, parameters: Sequence, state_dim: int, obs_dim: int, observe_every_step: int = 1):
        super().__init__(UserProcess(parameters, state_dim), MODEL_USER, (), torch.Size([obs_dim]) if obs_dim > 1 else torch.Size([]),
                         observe_every_step)
        self.library_path = library_path

    def library(self):
        from . import _lib

        return _lib.load_user_library(self.library_path)

    def sample_states(self, *a, **k):
        raise NotImplementedError("a user model exists as device code only")


def compile_user_model(source: str, state_dim: int = 1, obs_dim: int = 1):
    """Compiles ``source`` - CUDA C++ defining ``struct UserModel`` (see csrc/models.h for the contract: ``D``, ``OD``, ``NRAW``,
    ``loc_scale``, ``obs_lp``, ``obs_sample`` as ``__device__`` functions and the host function ``derive``) - into its own build of the
    library (nvcc, about a minute the first time; cached in-tree by the hash of the source) and returns a factory
    ``make(*parameters, observe_every_step=1) -> UserStateSpaceModel`` (parameters: floats or ``(B,)`` tensors, ``NRAW`` of them)."""
    import hashlib
    import os

    from . import _lib

    here = os.path.dirname(os.path.abspath(__file__))
    tmp_dir = os.path.join(here, "_user", "src")
    os.makedirs(tmp_dir, exist_ok=True)
    header = os.path.join(tmp_dir, hashlib.sha256(source.encode()).hexdigest()[:16] + ".h")
    if not os.path.exists(header):
        with open(header, "w") as fh:
            fh.write(source)
    so = _lib.build_user_library(header)

    def make(*parameters, observe_every_step: int = 1) -> UserStateSpaceModel:
        return UserStateSpaceModel(so, parameters, state_dim, obs_dim, observe_every_step)

    make.library_path = so
    return make
