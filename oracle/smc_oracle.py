"""CPU oracle for the Sequential-Monte-Carlo inner loop of tingiskhan/pyfilter (reference v0.29.0).

THIS FILE IS TEST INFRASTRUCTURE.  It restates, on torch-CPU / numpy, the algorithm behind
``pyfilter.filters.particle.{SISR,APF}.batch_filter`` and ``pyfilter.resampling.{systematic,multinomial}``
so that the CUDA path in ``pyfilter_b200`` can be checked against it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product path never does (it fails loudly when the CUDA library is missing).

Pinning: ``tests/test_oracle_pinned.py`` checks every function below against (a) the unmodified reference
package imported from /root/reference (build container only; needs the stand-ins under ``oracle/standins``),
(b) the golden vectors under ``tests/golden/`` that ``oracle/make_golden.py`` generated from that same
reference, and (c) the reference's own known-answer test ``tests/test_resampling.py:31-47``.
The state-space-model arithmetic itself lives in the un-vendored dependency ``stochproc==0.3.0``
(pyproject.toml:32) - its semantics (``x_t = loc + scale*eps``; Euler-Maruyama ``loc = x + f(x) dt``) are
restated from its published behaviour and are pinned only through the stand-in, see DESIGN.md ("parity
unpinned at the stochproc boundary").

Layout follows the reference: particle-major ``x:(N,[B],[d])``, ``log w:(N,[B])``; float32 everywhere, int64
ancestor indices.  All file:line citations are into /root/reference/pyfilter/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch

INFTY = float("inf")
_LOG_SQRT_2PI = math.log(math.sqrt(2.0 * math.pi))


# ----------------------------------------------------------------------------------------------------------------------
# weights: normalize / ESS                                                                      utils.py:8-20, 49-64
# ----------------------------------------------------------------------------------------------------------------------
def normalize(weights: torch.Tensor) -> torch.Tensor:
    """``utils.py:49-64``: NaN-safe softmax over dim 0.  Mutates ``weights`` in place exactly like the reference
    (NaN and +inf become -inf, -inf becomes -FLT_MAX); columns whose soft-max sums to zero become uniform."""
    weights = weights.nan_to_num_(-INFTY, posinf=-INFTY)
    normalized = (weights - weights.max(dim=0)[0]).softmax(dim=0)
    ax_sum = normalized.sum(dim=0)
    normalized.masked_fill_(ax_sum == 0.0, 1.0 / normalized.shape[0])
    return normalized


def get_ess(weights: torch.Tensor, normalized: bool = False) -> torch.Tensor:
    """``utils.py:8-20``: ``1 / sum_i W_i^2`` over dim 0."""
    if not normalized:
        weights = normalize(weights)
    return weights.pow(2.0).sum(dim=0).reciprocal()


# ----------------------------------------------------------------------------------------------------------------------
# resampling                                                                                     resampling.py:8-105
# ----------------------------------------------------------------------------------------------------------------------
def _wrapped(f, w: torch.Tensor, normalized: bool, **kwargs) -> torch.Tensor:
    """``resampling.py:8-21``.  1-D input calls ``f(w, *kwargs)`` - i.e. keyword *names* are splatted
    positionally, so ``u=`` is effectively ignored for 1-D input: the positional
    argument that arrives is the string ``"u"`` bound to ``normalized`` and ``u`` stays ``None``."""
    if not normalized:
        w = normalize(w)
    if w.dim() == 1:
        return f(w)
    return f(w.moveaxis(0, 1), **kwargs).moveaxis(0, 1)


def _systematic_rows(w: torch.Tensor, u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``resampling.py:35-52`` on a ``(B,N)`` (or ``(N,)``) tensor of normalised weights."""
    is_1d = w.dim() == 1
    if is_1d:
        w = w.unsqueeze(0)
    shape = (w.shape[0], 1)
    u = u if u is not None else torch.empty(shape, device=w.device).uniform_()
    n = w.shape[1]
    index_range = torch.arange(n, dtype=u.dtype, device=w.device).unsqueeze(0)
    probs = (index_range + u) / n
    cumsum = w.cumsum(-1)
    cumsum[..., -1] = 1.0
    res = torch.searchsorted(cumsum, probs)
    return res.squeeze(0) if is_1d else res


def systematic(w: torch.Tensor, normalized: bool = False, u: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``pyfilter.resampling.systematic`` (``resampling.py:24-52``).  ``w``: ``(N,)`` or ``(N,B)``; ``u``: ``(B,1)``."""
    return _wrapped(_systematic_rows, w, normalized, **({} if u is None else {"u": u}))


def multinomial(w: torch.Tensor, normalized: bool = False) -> torch.Tensor:
    """``pyfilter.resampling.multinomial`` (``resampling.py:55-65``): ``torch.multinomial(W, N, replacement=True)``."""
    return _wrapped(lambda ww: torch.multinomial(ww, ww.shape[-1], replacement=True), w, normalized)


def _residual_1d(w: torch.Tensor) -> torch.Tensor:
    """``resampling.py:78-105`` on a 1-D tensor of normalised weights: ``floor(N w_i)`` deterministic copies of every particle
    (in particle order), the remaining ``N - sum floor`` slots drawn by ``torch.multinomial`` from the fractional parts
    (divided by the number of deterministic copies, resampling.py:88 - multinomial normalises anyway)."""
    if w.dim() > 1:
        raise NotImplementedError("Not implemented for multidimensional arrays!")
    n = w.shape[-1]
    mw = n * w
    floored = mw.floor()
    res = mw - floored
    out = torch.ones_like(w, dtype=torch.long)
    numelems = floored.sum(-1)
    res = res / numelems
    intpart = floored.long()
    ranged = torch.arange(n, dtype=intpart.dtype) * out
    modded = ranged.repeat_interleave(intpart)
    aslong = int(numelems.long())
    out[:aslong] = modded
    if aslong == n:
        return out
    out[aslong:] = torch.multinomial(res, n - aslong, replacement=True)
    return out


def residual(w: torch.Tensor, normalized: bool = False) -> torch.Tensor:
    """``pyfilter.resampling.residual`` (``resampling.py:66-105``); 1-D only, like the reference."""
    return _wrapped(_residual_1d, w, normalized)


def residual_restated(W: np.ndarray, U: np.ndarray) -> np.ndarray:
    """Numpy restatement of ``residual`` for one column of normalised float32 weights, given the float64 uniforms ``U`` the
    multinomial part consumes in draw order (``len(U) = N - sum floor(N w)``): float32 products ``N w`` and floors, the
    fractional parts divided (float32) by the number of deterministic copies, then ``multinomial_restated``'s rule on them
    (sequential float32 prefix sum divided by its last element, left search in double).  What a CUDA kernel must reproduce."""
    W = np.asarray(W, dtype=np.float32)
    n = W.shape[0]
    mw = (np.float32(n) * W).astype(np.float32)
    floored = np.floor(mw).astype(np.float32)
    counts = floored.astype(np.int64)
    k = int(counts.sum())
    out = np.ones(n, dtype=np.int64)
    out[:k] = np.repeat(np.arange(n, dtype=np.int64), counts)
    if k == n:
        return out
    res = ((mw - floored).astype(np.float32) / np.float32(k)).astype(np.float32)
    c = sequential_cumsum(res, np.float32)
    c = (c / c[-1]).astype(np.float32)
    out[k:] = np.searchsorted(c.astype(np.float64), np.asarray(U, dtype=np.float64)[: n - k], side="left")
    return out


def sequential_cumsum(w: np.ndarray, acc_dtype, out_dtype=np.float32) -> np.ndarray:
    """Sequential prefix sum accumulated in ``acc_dtype`` and rounded to ``out_dtype`` per element.

    ``acc_dtype=float64`` is what ``torch.cumsum`` does for float32 CPU tensors (ATen ``cumsum_cpu_kernel``
    accumulates in ``acc_type<float>=double``; SURVEY.md Appendix A-5, measured).  ``acc_dtype=float32`` is the
    prefix sum inside ``torch.multinomial``'s CPU kernel (Appendix A-6).  ``np.cumsum`` is a plain sequential loop
    in the accumulator dtype, which is exactly these semantics."""
    return np.cumsum(np.asarray(w).astype(acc_dtype), dtype=acc_dtype).astype(out_dtype)


def systematic_probes(n: int, u: float) -> np.ndarray:
    """``resampling.py:44-46``: ``p_i = fl32(fl32(i + u) / n)`` with true IEEE division (torch CPU ``div``)."""
    i = np.arange(n, dtype=np.float32)
    return ((i + np.float32(u)).astype(np.float32) / np.float32(n)).astype(np.float32)


def systematic_restated(W: np.ndarray, u: float) -> np.ndarray:
    """Numpy restatement of ``systematic`` for ONE column of normalised float32 weights (no torch ops):
    sequential fp64-accumulated prefix sum, last element forced to 1, left-insertion search.  Used to pin what
    the CUDA kernel must reproduce bit-for-bit."""
    W = np.asarray(W, dtype=np.float32)
    c = sequential_cumsum(W, np.float64)
    c[-1] = np.float32(1.0)
    return np.searchsorted(c, systematic_probes(W.shape[0], u), side="left").astype(np.int64)


def multinomial_restated(W: np.ndarray, U: np.ndarray) -> np.ndarray:
    """Numpy restatement of ``torch.multinomial(W, n, replacement=True)`` on CPU for one row, given the float64
    uniforms ``U`` in draw order (ATen ``multinomial_with_replacement_kernel``):
    sequential **float32** prefix sum, divided by its last element, then a left binary search evaluated in double."""
    W = np.asarray(W, dtype=np.float32)
    c = sequential_cumsum(W, np.float32)
    c = (c / c[-1]).astype(np.float32)
    return np.searchsorted(c.astype(np.float64), np.asarray(U, dtype=np.float64), side="left").astype(np.int64)


def systematic_loop(W, u) -> np.ndarray:
    """Pure-Python two-pointer walk (small inputs only): ancestor of probe i is the first j with c_j >= p_i."""
    W = np.asarray(W, dtype=np.float32)
    n = W.shape[0]
    p = systematic_probes(n, u)
    out = np.zeros(n, dtype=np.int64)
    acc, j = np.float64(W[0]), 0
    c = np.float32(acc) if n > 1 else np.float32(1.0)
    for i in range(n):
        while c < p[i]:
            j += 1
            acc = acc + np.float64(W[j])
            c = np.float32(1.0) if j == n - 1 else np.float32(acc)
        out[i] = j
    return out


# ----------------------------------------------------------------------------------------------------------------------
# estimators                                                                              filters/particle/utils.py
# ----------------------------------------------------------------------------------------------------------------------
def log_likelihood(importance_weights: torch.Tensor, weights: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``filters/particle/utils.py:7-22``: ``max + log sum_i omega_i exp(w_i - max)``; ``omega = 1/N`` if omitted."""
    max_w, _ = importance_weights.max(dim=0)
    temp = (importance_weights - max_w).exp()
    if weights is None:
        weights = 1.0 / importance_weights.shape[0]
    return max_w + (weights * temp).sum(dim=0).log()


def filter_mean_and_variance(x: torch.Tensor, W: torch.Tensor, event_dims: int):
    """``filters/particle/utils.py:26-65`` (``covariance=False, keep_dim=True``): two-pass weighted mean and
    variance over dim 0; scalar states get a trailing axis of length 1."""
    values = x.unsqueeze(-1) if event_dims == 0 else x
    w = W.unsqueeze(-1)
    mean = (w * values).sum(dim=0)
    centered = values - mean
    var = (w * centered.pow(2.0)).sum(dim=0)
    return mean, var


def normal_log_prob(value, loc, scale):
    """``torch.distributions.Normal.log_prob`` (the density every reference model callable ends in)."""
    scale = torch.as_tensor(scale, dtype=torch.float32)
    var = scale**2
    return -((value - loc) ** 2) / (2 * var) - scale.log() - _LOG_SQRT_2PI


# ----------------------------------------------------------------------------------------------------------------------
# model zoo (the four BASELINE.json state-space models)
#   hidden process:  x_t = loc(x_{t-1}) + scale(x_{t-1}) * inc,   inc = inc_scale * z,  z ~ N(0, I)      [stochproc AffineProcess]
#   observation:     log p(y | x)
# ----------------------------------------------------------------------------------------------------------------------
def _t(v):
    return torch.as_tensor(v, dtype=torch.float32)


@dataclass
class Model:
    """Base class.  Parameters are float32 tensors of shape ``()`` or ``(B,)`` (one value per parallel filter)."""

    name: str = "model"
    state_dim: int = 0  # 0 = scalar state (event_shape ()), else event_shape (state_dim,)
    obs_dim: int = 0
    inc_scale: float = 1.0  # std of the increment distribution: 1 for unit increments, sqrt(dt) for Euler-Maruyama
    linear_obs: Optional[Tuple] = None  # (a, b, s) when the observation is y = b + a x + s nu (LinearStateSpaceModel)

    def mean_scale(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    def obs_log_prob(self, y: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def initial_loc_scale(self) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    # -- helpers shared by all models
    def initial_sample(self, shape: Tuple[int, ...], z: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``filters/particle/base.py:91``: ``x_0 = loc_0 + scale_0 * z`` with ``z ~ N(0,1)`` of shape ``(N,[B],[d])``."""
        loc, scale = self.initial_loc_scale()
        full = tuple(shape) + ((self.state_dim,) if self.state_dim else ())
        if z is None:
            z = torch.empty(full).normal_()
        return loc + scale * z

    def propagate(self, x: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        """stochproc ``AffineProcess.propagate``: ``AffineTransform(loc, scale)(inc)`` = ``loc + scale * inc``."""
        loc, scale = self.mean_scale(x)
        inc = z * self.inc_scale if self.inc_scale != 1.0 else z
        return loc + scale * inc

    def simulate(self, T: int, generator: Optional[torch.Generator] = None):
        """Draw one path ``(x_{1:T}, y_{1:T})`` from the model (synthetic data for tests and the bench)."""
        g = generator
        d = (self.state_dim,) if self.state_dim else ()
        x = self.initial_sample((), torch.empty(d).normal_(generator=g))
        xs, ys = [], []
        for _ in range(T):
            x = self.propagate(x, torch.empty(d).normal_(generator=g))
            xs.append(x)
            ys.append(self.sample_obs(x, g))
        return torch.stack(xs), torch.stack(ys)

    def sample_obs(self, x, g):
        raise NotImplementedError


@dataclass
class LinearGaussianAR1(Model):
    """Config 1 (reference ``tests/filters/models.py:12-16``): ``x_t = alpha + beta x_{t-1} + sigma eps``,
    stationary ``x_0``; ``y_t = b + a x_t + s nu``."""

    alpha: torch.Tensor = field(default_factory=lambda: _t(0.0))
    beta: torch.Tensor = field(default_factory=lambda: _t(0.99))
    sigma: torch.Tensor = field(default_factory=lambda: _t(0.05))
    a: torch.Tensor = field(default_factory=lambda: _t(1.0))
    b: torch.Tensor = field(default_factory=lambda: _t(0.0))
    s: torch.Tensor = field(default_factory=lambda: _t(0.15))
    name: str = "lg_ar1"

    def __post_init__(self):
        for k in ("alpha", "beta", "sigma", "a", "b", "s"):
            setattr(self, k, _t(getattr(self, k)))
        self.linear_obs = (self.a, self.b, self.s)

    def mean_scale(self, x):
        return torch.broadcast_tensors(self.alpha + self.beta * x, self.sigma)

    def obs_log_prob(self, y, x):
        return normal_log_prob(y, self.b + self.a * x, self.s)

    def initial_loc_scale(self):
        return self.alpha, self.sigma / (1.0 - self.beta**2.0).sqrt()

    def sample_obs(self, x, g):
        return self.b + self.a * x + self.s * torch.empty(()).normal_(generator=g)


@dataclass
class SineDiffusion(Model):
    """Configs 2 and 5 (reference README.md:44-67): Euler-Maruyama of ``dx = sin(x - gamma) dt + sigma dW``,
    ``x_0 ~ N(0,1)``; ``y_t = b + a x_t + s nu``."""

    gamma: torch.Tensor = field(default_factory=lambda: _t(0.0))
    sigma: torch.Tensor = field(default_factory=lambda: _t(1.0))
    dt: float = 0.1
    a: torch.Tensor = field(default_factory=lambda: _t(1.0))
    b: torch.Tensor = field(default_factory=lambda: _t(0.0))
    s: torch.Tensor = field(default_factory=lambda: _t(0.1))
    name: str = "sine_em"

    def __post_init__(self):
        for k in ("gamma", "sigma", "a", "b", "s"):
            setattr(self, k, _t(getattr(self, k)))
        self.inc_scale = math.sqrt(self.dt)
        self.linear_obs = (self.a, self.b, self.s)

    def mean_scale(self, x):
        return torch.broadcast_tensors(x + torch.sin(x - self.gamma) * _t(self.dt), self.sigma)

    def obs_log_prob(self, y, x):
        return normal_log_prob(y, self.b + self.a * x, self.s)

    def initial_loc_scale(self):
        return torch.zeros_like(self.gamma), torch.ones_like(self.gamma)

    def sample_obs(self, x, g):
        return self.b + self.a * x + self.s * torch.empty(()).normal_(generator=g)


@dataclass
class StochasticVolatility(Model):
    """Config 3: ``x_t = mu + phi (x_{t-1} - mu) + sigma_v eps``, stationary ``x_0``; ``y_t ~ N(0, exp(x_t / 2))``."""

    mu: torch.Tensor = field(default_factory=lambda: _t(-1.0))
    phi: torch.Tensor = field(default_factory=lambda: _t(0.97))
    sigma_v: torch.Tensor = field(default_factory=lambda: _t(0.2))
    name: str = "sv_ar1"

    def __post_init__(self):
        for k in ("mu", "phi", "sigma_v"):
            setattr(self, k, _t(getattr(self, k)))

    def mean_scale(self, x):
        return torch.broadcast_tensors(self.mu + self.phi * (x - self.mu), self.sigma_v)

    def obs_log_prob(self, y, x):
        return normal_log_prob(y, 0.0, (x / 2.0).exp())

    def initial_loc_scale(self):
        return self.mu, self.sigma_v / (1.0 - self.phi**2.0).sqrt()

    def sample_obs(self, x, g):
        return (x / 2.0).exp() * torch.empty(()).normal_(generator=g)


@dataclass
class VerhulstSinhArcsinh(Model):
    """The model of the reference's ``examples/stochastic-volatility.ipynb:60-83``: a Verhulst volatility process
    ``dV = kappa V (gamma - V) dt + sigma V dU`` stepped by Euler-Maruyama (``dt = 0.2``, observed every ``1 / dt`` steps) and
    ``Y = mu + V s(W)``, ``W ~ N(0, 1)``, ``s(w) = sinh((asinh(w) + nu) tau)`` - the notebook's ``AffineTransform(mean, scale)`` after a
    ``SinhArcsinhTransform(skew, kurt)``.  ``Verhulst`` and the transform are classes of the absent stochproc package: the equations are
    the ones the notebook's first cell states, the initial law ``V_0 ~ N(gamma, sigma)`` is this repository's choice (parity unpinned
    at that boundary)."""

    kappa: torch.Tensor = field(default_factory=lambda: _t(0.1))
    gamma: torch.Tensor = field(default_factory=lambda: _t(1.0))
    sigma: torch.Tensor = field(default_factory=lambda: _t(0.1))
    mu: torch.Tensor = field(default_factory=lambda: _t(0.0))
    nu: torch.Tensor = field(default_factory=lambda: _t(0.0))
    tau: torch.Tensor = field(default_factory=lambda: _t(1.0))
    dt: float = 0.2
    name: str = "verhulst_sas"

    def __post_init__(self):
        for k in ("kappa", "gamma", "sigma", "mu", "nu", "tau"):
            setattr(self, k, _t(getattr(self, k)))
        self.inc_scale = math.sqrt(self.dt)

    def mean_scale(self, x):
        return torch.broadcast_tensors(x + self.kappa * x * (self.gamma - x) * self.dt, self.sigma * x)

    def obs_log_prob(self, y, x):
        # change of variables through the inverse transform w = sinh(asinh(r) / tau - nu), r = (y - mu) / v
        r = (y - self.mu) / x
        a = torch.asinh(r) / self.tau - self.nu
        w = torch.sinh(a)
        return (-0.5 * w * w - _LOG_SQRT_2PI + torch.cosh(a).log() - self.tau.log() - 0.5 * torch.log1p(r * r) - x.abs().log())

    def initial_loc_scale(self):
        return self.gamma, self.sigma

    def sample_obs(self, x, g):
        w = torch.empty(()).normal_(generator=g)
        return self.mu + x * torch.sinh((torch.asinh(w) + self.nu) * self.tau)


@dataclass
class Lorenz63(Model):
    """Config 4 (reference examples/lorenz.ipynb:53-117): Euler-Maruyama of the Lorenz-63 drift, ``dt=0.01``, unit
    diffusion, ``x_0 ~ N(m_0, sqrt(10) I)``; ``y_t = 0.8 (x^1, x^3) + sqrt(0.1) nu`` (2-D observation)."""

    s: torch.Tensor = field(default_factory=lambda: _t(10.0))
    r: torch.Tensor = field(default_factory=lambda: _t(28.0))
    b: torch.Tensor = field(default_factory=lambda: _t(8.0 / 3.0))
    sigma: torch.Tensor = field(default_factory=lambda: _t(1.0))
    dt: float = 0.01
    obs_a: float = 0.8
    obs_s: torch.Tensor = field(default_factory=lambda: _t(0.1).sqrt())
    name: str = "lorenz63_em"
    state_dim: int = 3
    obs_dim: int = 2

    def __post_init__(self):
        for k in ("s", "r", "b", "sigma", "obs_s"):
            setattr(self, k, _t(getattr(self, k)))
        self.inc_scale = math.sqrt(self.dt)
        self.x0_mean = _t([-5.91652, -5.52332, 24.5723])
        self.x0_scale = math.sqrt(10.0) * torch.ones(3)
        # the same observation as a LinearStateSpaceModel (examples/lorenz.ipynb:105-117): y = b + A x + s nu, A (2, 3), b, s (1,)
        s1 = self.obs_s.reshape(-1)[:1] if self.obs_s.dim() else self.obs_s.unsqueeze(-1)
        self.linear_obs = (_t([[self.obs_a, 0.0, 0.0], [0.0, 0.0, self.obs_a]]), torch.zeros_like(s1), s1)

    def _p(self, p):  # parameter (B,) -> broadcast against (N,B)
        return p

    def mean_scale(self, x):
        x0, x1, x2 = x[..., 0], x[..., 1], x[..., 2]
        f0 = -self.s * (x0 - x1)
        f1 = self.r * x0 - x1 - x0 * x2
        f2 = x0 * x1 - self.b * x2
        drift = torch.stack((f0, f1, f2), dim=-1)
        sig = self.sigma.unsqueeze(-1) if self.sigma.dim() > 0 else self.sigma
        return torch.broadcast_tensors(x + drift * _t(self.dt), sig)

    def obs_loc(self, x):
        return torch.stack((self.obs_a * x[..., 0], self.obs_a * x[..., 2]), dim=-1)

    def obs_log_prob(self, y, x):
        return normal_log_prob(y, self.obs_loc(x), self.obs_s).sum(-1)

    def initial_loc_scale(self):
        return self.x0_mean, self.x0_scale

    def sample_obs(self, x, g):
        return self.obs_loc(x) + self.obs_s * torch.empty(2).normal_(generator=g)


# ----------------------------------------------------------------------------------------------------------------------
# proposals                                                               filters/particle/proposals/{bootstrap,linear}.py
# ----------------------------------------------------------------------------------------------------------------------
def bootstrap_sample_and_weight(model: Model, y, x_prev, z):
    """``proposals/bootstrap.py:10-14``: propagate through the dynamics, weight by the observation density."""
    x_new = model.propagate(x_prev, z)
    return x_new, model.obs_log_prob(y, x_new)


def affine_pre_weight(model: Model, y, x_prev):
    """``proposals/base.py:69-85`` with ``pre_weight_funcs.py:9-11``: ``log p(y | loc(x_{t-1}))``."""
    loc, _ = model.mean_scale(x_prev)
    return model.obs_log_prob(y, loc)


def lgo_sample_and_weight(model: Model, y, x_prev, z):
    """``proposals/linear.py:38-55`` + ``proposals/utils.py:219-267`` + ``proposals/base.py:45-50`` for a SCALAR
    state and observation: optimal Gaussian kernel ``N(k, P)``, ``P = 1/(sigma^-2 + a^2 s^-2)``,
    ``k = P (sigma^-2 m + a s^-2 (y - b))``; weight ``log p(y|x') + log p(x'|x) - log N(x'; k, P)``."""
    assert model.linear_obs is not None
    if model.state_dim > 0:
        return _lgo_sample_and_weight_nd(model, y, x_prev, z)
    assert model.obs_dim == 0
    a, b, s = model.linear_obs
    mean, scale = model.mean_scale(x_prev)
    h_var_inv = scale.pow(-2.0)
    o_var_inv = s.pow(-2.0)
    # find_optimal_density, hidden_is_1d and obs_is_1d: 1x1 "matrices"
    t_2 = a * o_var_inv * a
    cov = (h_var_inv + t_2).reciprocal()  # .inverse() of a 1x1 matrix
    t_1 = h_var_inv * mean
    t_3 = a * (o_var_inv * (y - b))
    k_mean = cov * (t_1 + t_3)
    k_std = cov.sqrt()
    x_new = z * k_std + k_mean  # torch.normal(mean, std)
    # _weight_with_kernel
    y_lp = normal_log_prob(y, b + a * x_new, s)
    inc_std = _t(model.inc_scale)
    # TransformedDistribution(inc, Affine(loc, scale)).log_prob(x) = inc.log_prob((x - loc)/scale) - log|scale|
    x_lp = normal_log_prob((x_new - mean) / scale, 0.0, inc_std) - scale.abs().log()
    k_lp = normal_log_prob(x_new, k_mean, k_std)
    return x_new, y_lp + x_lp - k_lp


def lgo_pre_weight(model: Model, y, x_prev):
    """``proposals/linear.py:57-86`` for scalar state/observation: ``log N(y; b + a x_{t-1}, sqrt(s^2 + a^2 sigma^2))``
    (centred on the PREVIOUS state, Appendix A-8)."""
    if model.state_dim > 0:
        return _lgo_pre_weight_nd(model, y, x_prev)
    a, b, s = model.linear_obs
    _, h_scale = model.mean_scale(x_prev)
    cov = s.pow(2.0) + a * h_scale.pow(2.0) * a
    return normal_log_prob(y, b + a * x_prev, cov.sqrt())


def _diag_from_flat(v: torch.Tensor, dim: int) -> torch.Tensor:
    """``utils.py:23-46`` for a vector event shape: ``eye(dim) * v.unsqueeze(-1)``."""
    return torch.eye(dim, dtype=v.dtype) * v.unsqueeze(-1)


def _lgo_sample_and_weight_nd(model: Model, y, x_prev, z):
    """``proposals/linear.py:38-55`` + ``proposals/utils.py:219-267`` + ``proposals/base.py:45-50`` for a VECTOR state (d) and a
    vector observation (m), as in examples/lorenz.ipynb:214 (oracle only so far).  With ``A (m, d)``,
    ``(mean, sigma) = mean_scale(x)``: ``P = (diag(sigma^-2) + A^T diag(s^-2) A)^-1``, ``k = P (sigma^-2 mean + A^T diag(s^-2) (y - b))``,
    ``x' = k + chol(P) z``; weight ``log p(y|x') + log p(x'|x) - log N(x'; k, P)``.  Same torch ops in the same order as the
    reference (batched ``inverse``, ``cholesky_ex``, ``MultivariateNormal.log_prob``)."""
    from torch.distributions import AffineTransform, Independent, MultivariateNormal, Normal, TransformedDistribution
    from torch.linalg import cholesky_ex

    a, b, s = model.linear_obs
    d, m = model.state_dim, model.obs_dim
    mean, scale = model.mean_scale(x_prev)
    h_var_inv = scale.pow(-2.0)
    o_var_inv = s.pow(-2.0)
    yc = y - b
    # find_optimal_density
    c_t = a.transpose(-2, -1)
    o_inv_cov = _diag_from_flat(o_var_inv, m)
    t_2 = c_t.matmul(o_inv_cov).matmul(a)
    cov = (_diag_from_flat(h_var_inv, d) + t_2).inverse()
    t_1 = h_var_inv * mean
    t_2 = o_inv_cov.matmul(yc)
    t_3 = c_t.matmul(t_2.unsqueeze(-1))
    k_mean = cov.matmul(t_1.unsqueeze(-1) + t_3).squeeze(-1)
    kernel = MultivariateNormal(k_mean, scale_tril=cholesky_ex(cov)[0], validate_args=False)
    # MultivariateNormal.rsample: loc + scale_tril @ eps
    x_new = kernel.loc + torch.matmul(kernel._unbroadcasted_scale_tril, z.unsqueeze(-1)).squeeze(-1)
    # _weight_with_kernel
    y_lp = Independent(Normal(b + (a @ x_new.unsqueeze(-1)).squeeze(-1), s, validate_args=False), 1).log_prob(y)
    inc = Independent(Normal(torch.zeros(d), _t(model.inc_scale).expand(d), validate_args=False), 1)
    x_lp = TransformedDistribution(inc, AffineTransform(mean, scale, event_dim=1), validate_args=False).log_prob(x_new)
    return x_new, y_lp + x_lp - kernel.log_prob(x_new)


def _lgo_pre_weight_nd(model: Model, y, x_prev):
    """``proposals/linear.py:57-86``, vector case: ``log N(y; b + A x_{t-1}, diag(s^2) + A diag(sigma^2) A^T)``."""
    from torch.distributions import MultivariateNormal
    from torch.linalg import cholesky_ex

    a, b, s = model.linear_obs
    d, m = model.state_dim, model.obs_dim
    _, h_scale = model.mean_scale(x_prev)
    cov = _diag_from_flat(s.pow(2.0), m) + a.matmul(_diag_from_flat(h_scale.pow(2.0), d)).matmul(a.transpose(-2, -1))
    o_loc = b + (a @ x_prev.unsqueeze(-1)).squeeze(-1)
    return MultivariateNormal(o_loc, scale_tril=cholesky_ex(cov)[0], validate_args=False).log_prob(y)


def transition_log_prob(model: Model, x_new, x_prev):
    """``model.hidden.build_density(state).log_prob(x_new)`` (stochproc ``AffineProcess.build_density``): the increment distribution
    ``N(0, inc_scale)`` pushed through ``AffineTransform(loc, scale)``, summed over the event dimension."""
    loc, scale = model.mean_scale(x_prev)
    inc_std = _t(model.inc_scale)
    lp = normal_log_prob((x_new - loc) / scale, 0.0, inc_std) - scale.abs().log()
    return lp.sum(-1) if model.state_dim else lp


def linearized_sample_and_weight(model: Model, y, x_prev, z, n_steps: int = 1, alpha: float = 1e-4, second_order: bool = False):
    """``Linearized.sample_and_weight`` (``proposals/linearized.py:53-70``) with ``ModeFinder.find_mode`` (``proposals/utils.py:96-146``, the
    default functorch path), as written: starting from ``x = mean``, ``n_steps`` times ``x += step`` where ``step`` is the CONSTANT
    ``alpha`` for the first-order variant (the gradient is evaluated but not used: proposals/utils.py:119,137) and the Newton-like
    ``cov * gradient`` with ``cov = -(H - clip(2 H, 0))^-1`` (scalar state) / ``-pinv(H - clip(2 lambda_min, 0) I)`` (vector state) for
    ``use_second_order``, ``H`` and the gradient those of ``log p(y | x) + log p(x | x_prev)`` at the current ``x`` (torch.func here,
    functorch there).  Kernel ``N(x, std)`` with ``std`` the hidden scale (first order) or ``sqrt(cov)`` / ``chol(cov)`` of the LAST step;
    weight ``log p(y|x') + log p(x'|x_prev) - kernel.log_prob(x')`` (``proposals/base.py:45-50``)."""
    from torch.func import grad, hessian, vmap

    d = model.state_dim
    mean, std0 = model.mean_scale(x_prev)
    mean, std0 = torch.broadcast_tensors(mean, std0)
    shape = x_prev.shape
    flat = lambda t: t.reshape((-1, d) if d else (-1,))
    # per-particle parameters: models with (B,) parameters are evaluated column by column through a tiny per-particle model
    n_part = flat(mean).shape[0]

    def joint(xi, xpi, col):
        mo = model if col is None else _column_model(model, col)
        return mo.obs_log_prob(y, xi) + transition_log_prob(mo, xi, xpi)

    batched = x_prev.dim() - (1 if d else 0) == 2
    B = shape[1] if batched else 1
    x = mean.clone()
    std = std0.clone()
    chol = None
    for _ in range(n_steps):
        if second_order:
            cols = []
            for b in range(B):
                xb = x[:, b] if batched else x
                pb = x_prev[:, b] if batched else x_prev
                fn = lambda xi, xpi: joint(xi, xpi, b if batched else None)
                g = vmap(grad(fn))(xb, pb).to(xb.dtype)
                H = vmap(hessian(fn))(xb, pb).to(xb.dtype)   # (float32 like the reference's: python-float constants of this module upcast it)
                if d == 0:
                    d_h = (2.0 * H).clip(min=0.0)
                    cov = -(H - d_h).pow(-1)
                    cols.append((cov * g, cov.sqrt(), None))
                else:
                    lam = torch.linalg.eigvalsh(H).real.min(dim=-1).values
                    d_h = (2.0 * lam).clip(min=0.0).view(-1, 1, 1) * torch.eye(d)
                    cov = -torch.linalg.pinv(H - d_h)
                    cols.append(((cov @ g.unsqueeze(-1)).squeeze(-1), None, torch.linalg.cholesky_ex(cov)[0]))
            step = torch.stack([c[0] for c in cols], 1) if batched else cols[0][0]
            if d == 0:
                std = torch.stack([c[1] for c in cols], 1) if batched else cols[0][1]
            else:
                chol = torch.stack([c[2] for c in cols], 1) if batched else cols[0][2]
            x = x + step
        else:
            x = x + alpha   # (sic) proposals/utils.py:119,137
    if d and second_order:
        kernel = torch.distributions.MultivariateNormal(x, scale_tril=chol, validate_args=False)
        x_new = x + (chol @ z.unsqueeze(-1)).squeeze(-1)
        k_lp = kernel.log_prob(x_new)
    else:
        x_new = x + std * z
        k_lp = normal_log_prob(x_new, x, std)
        if d:
            k_lp = k_lp.sum(-1)
    w = model.obs_log_prob(y, x_new) + transition_log_prob(model, x_new, x_prev) - k_lp
    return x_new, w


def _column_model(model: Model, b: int) -> Model:
    """The same model with every ``(B,)`` parameter reduced to its ``b``-th entry (per-particle autograd wants scalars)."""
    import copy

    mo = copy.copy(model)
    for k, v in vars(model).items():
        if isinstance(v, torch.Tensor) and v.dim() == 1 and v.shape[0] > 1 and k not in ("x0_mean", "x0_scale"):
            setattr(mo, k, v[b])
    return mo


PROPOSALS = {
    "bootstrap": (bootstrap_sample_and_weight, affine_pre_weight),
    "linear_gaussian": (lgo_sample_and_weight, lgo_pre_weight),
}


def nested_sample_and_weight(model: Model, y, x_prev, noise, num_samples: int):
    """``NestedProposal.sample_and_weight`` (``proposals/nested.py:27-47``) on replayed draws: ``noise = (zs, E)`` with ``zs`` the standard
    normals behind ``hidden_density.sample(num_samples)`` - shape ``(M, N, [B], [d])`` - and ``E`` the float32 Exp(1) values behind
    ``Categorical(probs).sample()``, shape ``(N, [B], M)``: ``torch.multinomial`` draws a single sample per row as
    ``argmax(probs / q)``, ``q ~ Exp(1)`` (aten/src/ATen/native/Distributions.cpp, the n_sample == 1 path)."""
    zs, E = noise
    d = model.state_dim
    mean, scale = model.mean_scale(x_prev)
    samples = mean + scale * (zs * _t(model.inc_scale))
    log_prob = model.obs_log_prob(y, samples).nan_to_num(-float("inf"), -float("inf"))
    probs = log_prob.softmax(dim=0)
    probs = probs.masked_fill(probs.isnan(), 1.0 / num_samples)
    rows = probs.moveaxis(0, -1)
    best = (rows / torch.as_tensor(E, dtype=rows.dtype).reshape(rows.shape)).argmax(dim=-1)
    idx = best.unsqueeze(0)
    if d:
        idx = idx.unsqueeze(-1).expand((1,) + tuple(samples.shape[1:]))
    x_new = samples.gather(0, idx).squeeze(0)
    return x_new, log_prob.exp().mean(dim=0).log()


class _ProposalTable(dict):
    """``"linearized:<n_steps>:<alpha>:<0|1 second order>"`` names a configured ``Linearized`` proposal (default pre-weight)."""

    def __missing__(self, key):
        if isinstance(key, str) and key.startswith("linearized"):
            parts = key.split(":")
            n_steps = int(parts[1]) if len(parts) > 1 else 1
            alpha = float(parts[2]) if len(parts) > 2 else 1e-4
            second = bool(int(parts[3])) if len(parts) > 3 else False
            fn = lambda model, y, x_prev, z: linearized_sample_and_weight(model, y, x_prev, z, n_steps, alpha, second)
            self[key] = (fn, affine_pre_weight)
            return self[key]
        if isinstance(key, str) and key.startswith("nested"):   # "nested:<num_samples>"; the step's noise argument is (zs, U)
            m = int(key.split(":")[1])
            self[key] = (lambda model, y, x_prev, noise: nested_sample_and_weight(model, y, x_prev, noise, m), affine_pre_weight)
            return self[key]
        raise KeyError(key)


PROPOSALS = _ProposalTable(PROPOSALS)


# ----------------------------------------------------------------------------------------------------------------------
# one filter step, noise injected (teacher-forced parity protocol)
# ----------------------------------------------------------------------------------------------------------------------
def _gather0(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """``filters/utils.py:4-21``."""
    if x.dim() > idx.dim():
        idx = idx.unsqueeze(-1).expand_as(x)
    return x.gather(0, idx)


def _as_2d_resample(W: torch.Tensor, u, resampler: str, U=None) -> torch.Tensor:
    """Resample every column of ``W`` (``(N,)`` or ``(N,B)``) with injected randomness."""
    squeeze = W.dim() == 1
    W2 = W.unsqueeze(-1) if squeeze else W
    if resampler == "systematic":
        uu = torch.as_tensor(u, dtype=torch.float32).reshape(-1, 1)
        idx = _systematic_rows(W2.moveaxis(0, 1), uu).moveaxis(0, 1)
    elif resampler == "multinomial":
        cols = [torch.from_numpy(multinomial_restated(W2[:, b].numpy(), np.asarray(U)[..., b] if np.asarray(U).ndim > 1 else U))
                for b in range(W2.shape[1])]
        idx = torch.stack(cols, dim=1)
    else:
        raise ValueError(resampler)
    return idx.squeeze(-1) if squeeze else idx


def sisr_step(model: Model, proposal: str, x, lw, prev_inds, y, z, u=None, ess_threshold=0.9,
              resampler="systematic", U=None, force_idx=None) -> Dict[str, torch.Tensor]:
    """One ``SISR`` move (``filters/particle/sisr.py:14-56``, ``filters/base.py:188-221``) with injected noise.

    ``x:(N,[B],[d])``, ``lw:(N,[B])``, ``z`` unit normals shaped like ``x``, ``u:(B,)`` systematic offsets for ALL
    columns (only the resampled ones are used).  A NaN ``y`` propagates only (``particle/state.py:38-42``).
    ``force_idx`` (test hook): ancestors to use instead of the resampler's (parity tests feed the device's ancestors so that every
    other quantity of the move can be compared even when an ulp of a weight moved one ancestor)."""
    lw = lw.clone()
    n = lw.shape[0]
    W = normalize(lw)  # sisr.py:16 (mutates lw like the reference)
    ess = get_ess(W, normalized=True)
    mask = ess < ess_threshold * n  # sisr.py:19, threshold base.py:42
    inds = prev_inds.clone()
    x_res, lw_res = x, lw
    if bool(mask.any()):
        all_idx = _as_2d_resample(W, u, resampler, U) if force_idx is None else force_idx
        if W.dim() == 1:
            inds, x_res = all_idx, _gather0(x, all_idx)
            lw_res = torch.zeros_like(lw)
            W = torch.full_like(W, 1.0 / n)
        else:
            um = mask.unsqueeze(0)
            inds = torch.where(um, all_idx, prev_inds)
            lw_res = lw.masked_fill(um, 0.0)
            W = W.masked_fill(um, 1.0 / n)
            sel = torch.where(um, all_idx, torch.arange(n).unsqueeze(-1).expand_as(all_idx))
            x_res = _gather0(x, sel)
    out = {"resampled": mask.clone(), "ess": ess, "prev_inds": inds, "x_resampled": x_res}
    if bool(torch.isnan(torch.as_tensor(y)).all()):
        x_new = model.propagate(x_res, z)
        lw_new, ll = lw_res, torch.zeros(lw.shape[1:])
    else:
        x_new, inc = PROPOSALS[proposal][0](model, y, x_res, z)
        lw_new = inc + lw_res  # sisr.py:52
        ll = log_likelihood(inc, W)  # sisr.py:55
    mean, var = filter_mean_and_variance(x_new, normalize(lw_new), model.state_dim)
    out.update(x=x_new, lw=lw_new, ll=ll, mean=mean, var=var)
    return out


def apf_step(model: Model, proposal: str, x, lw, prev_inds, y, z, u=None, resampler="systematic", U=None, force_idx=None):
    """One ``APF`` move (``filters/particle/apf.py:16-46``) with injected noise; resamples every step.  ``force_idx``: see
    :func:`sisr_step`."""
    lw = lw.clone()
    n = lw.shape[0]
    W = normalize(lw)  # apf.py:17
    if bool(torch.isnan(torch.as_tensor(y)).all()):  # filters/base.py:213, predict does not resample
        x_new = model.propagate(x, z)
        mean, var = filter_mean_and_variance(x_new, normalize(lw.clone()), model.state_dim)
        idx = torch.arange(n) if lw.dim() == 1 else torch.arange(n).unsqueeze(-1).expand(lw.shape)
        return {"x": x_new, "lw": lw, "ll": torch.zeros(lw.shape[1:]), "mean": mean, "var": var, "prev_inds": idx,
                "resample_W": W}
    sample_and_weight, pre_weight = PROPOSALS[proposal]
    g = pre_weight(model, y, x)  # apf.py:27
    rw = normalize(g + lw)  # apf.py:29-31 -> resampling.py:10-11
    idx = _as_2d_resample(rw, u, resampler, U) if force_idx is None else force_idx
    x_res = _gather0(x, idx)  # apf.py:34
    x_new, inc = sample_and_weight(model, y, x_res, z)  # apf.py:41
    lw_new = inc - g.gather(0, idx)  # apf.py:43
    ll = log_likelihood(lw_new) + (W * g.exp()).sum(dim=0).log()  # apf.py:44
    mean, var = filter_mean_and_variance(x_new, normalize(lw_new), model.state_dim)
    return {"x": x_new, "lw": lw_new, "ll": ll, "mean": mean, "var": var, "prev_inds": idx, "resample_W": rw,
            "pre_weight": g, "x_resampled": x_res}


def gpf_step(model: Model, proposal: str, x, lw, prev_inds, y, z, u=None, resampler="systematic", U=None, force_idx=None, **_):
    """One move of the Gaussian particle filter: ``GPF.predict`` / ``.correct`` (``filters/particle/gpf.py:27-36``) with the default
    ``GaussianProposal.sample_and_weight`` (``proposals/approximate.py:19-35``) and ``get_predictive_density(approximate=True)``
    (``particle/state.py:59-70``).  ``z = (z1, z2)``: the N(0, 1) draws of ``hidden.propagate`` and of the sample from the Gaussian
    approximation (one draw for a propagate-only move: ``z`` may then be ``z1`` alone).  The weights are replaced, not accumulated;
    the ancestors stay what they were."""
    z1, z2 = z if isinstance(z, (tuple, list)) else (z, None)
    lw = lw.clone()
    W = normalize(lw)                                  # state.normalized_weights() (mutates like the reference)
    d = model.state_dim
    x_prop = model.propagate(x, z1)
    if bool(torch.isnan(torch.as_tensor(y)).all()):     # filters/base.py:213-214 -> particle/state.py:38-42
        x_new, lw_new, ll = x_prop, lw, torch.zeros(lw.shape[1:])
    else:
        values = x_prop if d else x_prop.unsqueeze(-1)  # particle/utils.py:42-58 with covariance=True, keep_dim=False
        w = W.unsqueeze(-1)
        mean = (w * values).sum(dim=0)
        centered = values - mean
        if d == 0:
            var = (w * centered.pow(2.0)).sum(dim=0).squeeze(-1)
            x_new = mean.squeeze(-1) + var.sqrt() * z2  # Normal(mean, var.sqrt()).expand(...).sample()
        else:
            cov = torch.einsum("b...,b...ij->...ij", W, centered.unsqueeze(-1) @ centered.unsqueeze(-2))
            L = torch.linalg.cholesky(cov)              # MultivariateNormal(mean, covariance_matrix=cov).sample()
            x_new = mean + (L @ z2.unsqueeze(-1)).squeeze(-1)
        lw_new = model.obs_log_prob(y, x_new)
        ll = log_likelihood(lw_new)
    fm, fv = filter_mean_and_variance(x_new, normalize(lw_new), d)
    return dict(x=x_new, lw=lw_new, ll=ll, mean=fm, var=fv, prev_inds=prev_inds)


STEPS = {"sisr": sisr_step, "apf": apf_step, "gpf": gpf_step}


# ----------------------------------------------------------------------------------------------------------------------
# free-running driver (CPU baseline + statistical checks)                                   filters/base.py:140-158
# ----------------------------------------------------------------------------------------------------------------------
def batch_filter(model: Model, algorithm: str, proposal: str, y: torch.Tensor, particles: int,
                 batch_shape: Tuple[int, ...] = (), resampler: str = "systematic", ess_threshold: float = 0.9,
                 x0: Optional[torch.Tensor] = None, observe_every_step: int = 1, record_states: bool = False) -> Dict[str, torch.Tensor]:
    """``BaseFilter.batch_filter`` for SISR/APF using torch's global CPU generator in the reference's draw order
    (Appendix A-15: ``u`` - only when something resamples - then the transition noise).  ``record_states=True`` also returns every
    recorded state ``(x, log w, previous indices)``, the initial one included (``filters/result.py:39,119-133``)."""
    shape = (particles,) + tuple(batch_shape)
    d = (model.state_dim,) if model.state_dim else ()
    x = model.initial_sample(shape) if x0 is None else x0
    lw = torch.zeros(shape)
    n = particles
    inds = torch.arange(n) if not batch_shape else torch.arange(n).unsqueeze(-1).expand(shape)
    ll_total = torch.zeros(tuple(batch_shape))
    mean, var = filter_mean_and_variance(x, normalize(lw.clone()), model.state_dim)
    means, variances = [mean], [var]
    states = [(x, lw, inds)]
    nb = int(np.prod(batch_shape)) if batch_shape else 1
    # filters/base.py:204-210: before an observation is used, the filter propagates (predict + propagate, no weighting, no
    # likelihood, nothing recorded) until the time index of the state is a multiple of `observe_every_step`
    time_index = 0
    moves = []
    for y_t in y:
        while time_index % observe_every_step != 0:
            moves.append((torch.full_like(torch.as_tensor(y_t), float("nan")), False))
            time_index += 1
        moves.append((y_t, True))
        time_index += 1
    for y_t, recorded in moves:
        # which columns draw resampling randomness (reference order: u first, then the transition noise)
        if algorithm == "gpf":   # gpf.py: nothing resamples; the propagation's draws, then (observed moves) the Gaussian sample's
            z1 = torch.empty(shape + d).normal_()
            z2 = None if bool(torch.isnan(y_t).all()) else torch.empty(shape + d).normal_()
            out = gpf_step(model, proposal, x, lw, inds, y_t, (z1, z2))
            x, lw, inds = out["x"], out["lw"], out["prev_inds"]
            if recorded:
                ll_total = ll_total + out["ll"]
                means.append(out["mean"])
                variances.append(out["var"])
                states.append((x, lw, inds))
            continue
        if algorithm == "sisr":
            mask = (get_ess(normalize(lw.clone()), True) < ess_threshold * n).reshape(-1)  # sisr.py:16-19
        else:
            mask = torch.full((nb,), not bool(torch.isnan(y_t).all()))  # apf.py:29-31 resamples every observed step
        u = U = None
        if bool(mask.any()):
            k = int(mask.sum())
            if resampler == "systematic":
                u = torch.zeros(nb)
                u[mask] = torch.empty((k, 1)).uniform_().reshape(-1)  # resampling.py:40-41
            else:
                # torch.multinomial draws row by row (one row per resampled column), N float64 uniforms each
                Uf = np.zeros((n, nb))
                Uf[:, mask.numpy()] = torch.empty((k, n), dtype=torch.float64).uniform_().numpy().T
                U = Uf if batch_shape else Uf[:, 0]
        z = torch.empty(shape + d).normal_()
        if algorithm == "sisr":
            out = sisr_step(model, proposal, x, lw, inds, y_t, z, u, ess_threshold, resampler, U)
        else:
            out = apf_step(model, proposal, x, lw, inds, y_t, z, u, resampler, U)
        x, lw, inds = out["x"], out["lw"], out["prev_inds"]
        if not recorded:
            continue
        ll_total = ll_total + out["ll"]
        means.append(out["mean"])
        variances.append(out["var"])
        states.append((x, lw, inds))
    res = {"loglikelihood": ll_total, "filter_means": torch.stack(means), "filter_variance": torch.stack(variances),
           "x": x, "lw": lw, "prev_inds": inds}
    if record_states:
        res["states"] = states
    return res


def smooth_fixed_lag(states) -> torch.Tensor:
    """``ParticleFilter._do_sample_fl`` (``filters/particle/base.py:130-146``; SURVEY.md 8(f) f3, oracle only): ancestral tracing.
    Walk the recorded states backwards; the lineage of final particle i at time t-1 is ``previous_indices_t[lineage_t[i]]``;
    returns the ``(T+1, N, [B], [d])`` paths.  ``states``: list of ``(x, log w, previous indices)``."""
    x_last, _, prev_last = states[-1]
    n = x_last.shape[0]
    result = [x_last]
    lineage = torch.arange(n) if prev_last.dim() == 1 else torch.arange(n).unsqueeze(-1).expand(prev_last.shape)
    latest_prev = prev_last
    for x_s, _, prev_s in reversed(states[:-1]):
        lineage = _gather0(latest_prev, lineage)
        result.append(_gather0(x_s, lineage))
        latest_prev = prev_s
    return torch.stack(result[::-1], dim=0)


# ----------------------------------------------------------------------------------------------------------------------
# closed-form Kalman filter for config 1 (replaces the absent pykalman in tests/filters/test_particle.py:64-111)
# ----------------------------------------------------------------------------------------------------------------------
def kalman_filter_1d(y: np.ndarray, alpha, beta, sigma, a, b, s, m0, p0):
    """Scalar Kalman recursion; ``y`` may contain NaN (skipped).  Returns filtered means, variances, total log-lik."""
    y = np.asarray(y, dtype=np.float64)
    m, p = float(m0), float(p0)
    means, variances, ll = [], [], 0.0
    for yt in y:
        m, p = alpha + beta * m, beta * beta * p + sigma * sigma
        if not np.isnan(yt):
            sv = a * a * p + s * s
            resid = yt - (b + a * m)
            ll += -0.5 * (math.log(2.0 * math.pi * sv) + resid * resid / sv)
            k = p * a / sv
            m, p = m + k * resid, (1.0 - k * a) * p
        means.append(m)
        variances.append(p)
    return np.array(means), np.array(variances), ll


# ----------------------------------------------------------------------------------------------------------------------
# factory shared by tests / golden files / bench
# ----------------------------------------------------------------------------------------------------------------------
DEFAULT_PARAMS = {
    "lg_ar1": dict(alpha=0.0, beta=0.99, sigma=0.05, a=1.0, b=0.0, s=0.15),
    "sine_em": dict(gamma=0.0, sigma=1.0, dt=0.1, a=1.0, b=0.0, s=0.1),
    "sv_ar1": dict(mu=-1.0, phi=0.97, sigma_v=0.2),
    "lorenz63_em": dict(s=10.0, r=28.0, b=8.0 / 3.0, sigma=1.0, dt=0.01, obs_a=0.8, obs_s=math.sqrt(0.1)),
    "verhulst_sas": dict(kappa=0.1, gamma=1.0, sigma=0.1, mu=0.05, nu=-0.1, tau=1.1, dt=0.2),
}
_CLASSES = {"lg_ar1": LinearGaussianAR1, "sine_em": SineDiffusion, "sv_ar1": StochasticVolatility,
            "lorenz63_em": Lorenz63, "verhulst_sas": VerhulstSinhArcsinh}


def build_model(name: str, params: Optional[dict] = None) -> Model:
    p = dict(DEFAULT_PARAMS[name])
    p.update(params or {})
    return _CLASSES[name](**p)


def smooth_ffbs(model: Model, states, resampler: str = "systematic") -> torch.Tensor:
    """``ParticleFilter._do_sample_ffbs`` (``filters/particle/base.py:105-128``; SURVEY.md 8(f) f3, oracle only), restated as written in
    the reference (its own comment says "Something is wrong here": the backward weights are taken as they are).  Draws from torch's
    global generator in the reference's order: the resampler's offset for the last state, then one Categorical draw per earlier
    state.  ``states``: list of ``(x, log w, previous indices)``; non-batched filters only (the reference's batched branch moves axes
    in a way its own shapes do not support)."""
    from torch.distributions import AffineTransform, Categorical, Independent, Normal, TransformedDistribution

    x_last, lw_last, _ = states[-1]
    assert lw_last.dim() == 1, "non-batched only"
    idx = systematic(lw_last.clone()) if resampler == "systematic" else multinomial(lw_last.clone())
    res = [_gather0(x_last, idx)]
    d = model.state_dim
    for x_s, lw_s, _ in reversed(states[:-1]):
        loc, scale = model.mean_scale(x_s)
        if d:
            inc = Independent(Normal(torch.zeros(d), _t(model.inc_scale).expand(d), validate_args=False), 1)
        else:
            inc = Normal(0.0, _t(model.inc_scale), validate_args=False)
        density = TransformedDistribution(inc, AffineTransform(loc, scale, event_dim=1 if d else 0), validate_args=False)
        w_state = density.log_prob(res[-1].unsqueeze(1))   # (N_smoothed, N_particles)
        weights = lw_s.unsqueeze(0) + w_state
        indices = Categorical(logits=weights).sample()
        res.append(_gather0(x_s, indices))
    return torch.stack(res[::-1], dim=0)


def ffbs_backward_indices(model: Model, x_s, lw_s, x_later, U) -> torch.Tensor:
    """One backward step of ``_do_sample_ffbs`` (``filters/particle/base.py:112-122``) with the ``Categorical(logits=...).sample()`` draw
    replaced by the inversion it stands for, on replayed uniforms: the logits are computed exactly as in :func:`smooth_ffbs` (float32,
    the reference's own ops), the probabilities are accumulated in float64 and smoothed particle ``i`` takes the first index whose
    cumulative probability reaches ``U_i`` times the total.  What the CUDA kernel (csrc/plugin.cuh, ffbs_step_kernel) is compared with;
    the two may differ where ``U_i`` lies within float32 rounding of a boundary."""
    from torch.distributions import AffineTransform, Independent, Normal, TransformedDistribution

    d = model.state_dim
    loc, scale = model.mean_scale(x_s)
    if d:
        inc = Independent(Normal(torch.zeros(d), _t(model.inc_scale).expand(d), validate_args=False), 1)
    else:
        inc = Normal(0.0, _t(model.inc_scale), validate_args=False)
    density = TransformedDistribution(inc, AffineTransform(loc, scale, event_dim=1 if d else 0), validate_args=False)
    w_state = density.log_prob(x_later.unsqueeze(1))   # (N_smoothed, N_particles)
    weights = (lw_s.unsqueeze(0) + w_state).double()
    p = (weights - weights.max(-1, keepdim=True).values).exp()
    c = p.cumsum(-1)
    target = (torch.as_tensor(U, dtype=torch.float64) * c[:, -1]).unsqueeze(-1)
    return torch.searchsorted(c, target).squeeze(-1).clamp(max=x_s.shape[0] - 1)
