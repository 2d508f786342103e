"""Priors of the theta-particles and the parameter context (reference inference/prior.py:24-133, inference/context.py): what SMC2 /
PMMH need of them - sampling, the bijection to the unconstrained space the proposal kernel works in, ``eval_priors`` in that space,
``stack_parameters`` / ``unstack_parameters``, ``resample`` and ``exchange``.  Everything here is ``(B,)``-sized torch arithmetic on the
device (plumbing); the filters' particles are never touched."""
import math
from collections import OrderedDict
from typing import Dict

import torch


class Prior:
    """A univariate prior with its bijection from the unconstrained space (``biject_to(support)``, inference/prior.py:33-44)."""

    def sample(self, n: int, generator=None) -> torch.Tensor:   # constrained draws, on the CPU generator (torch.manual_seed governs them)
        raise NotImplementedError

    def get_constrained(self, u: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def get_unconstrained(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def eval_unconstrained(self, u: torch.Tensor) -> torch.Tensor:
        """``unconstrained_prior().log_prob(u)`` (inference/prior.py:81-90 with ``constrained=False``)."""
        raise NotImplementedError


def _normal_lp(v, loc, scale):
    return -((v - loc) ** 2) / (2.0 * scale * scale) - math.log(scale) - 0.5 * math.log(2.0 * math.pi)


class Normal(Prior):
    def __init__(self, loc: float, scale: float):
        self.loc, self.scale = float(loc), float(scale)

    def sample(self, n, generator=None):
        return self.loc + self.scale * torch.randn(n, generator=generator)

    def get_constrained(self, u):
        return u

    def get_unconstrained(self, x):
        return x

    def eval_unconstrained(self, u):
        return _normal_lp(u, self.loc, self.scale)


class LogNormal(Prior):
    """``LogNormal(loc, scale)``: positive support, bijection ``exp``; the unconstrained parameter is ``Normal(loc, scale)``."""

    def __init__(self, loc: float, scale: float):
        self.loc, self.scale = float(loc), float(scale)

    def sample(self, n, generator=None):
        return (self.loc + self.scale * torch.randn(n, generator=generator)).exp()

    def get_constrained(self, u):
        return u.exp()

    def get_unconstrained(self, x):
        return x.log()

    def eval_unconstrained(self, u):
        return _normal_lp(u, self.loc, self.scale)


class Exponential(Prior):
    """``Exponential(rate)``: positive support, bijection ``exp``; the unconstrained parameter ``u = log x`` has the density
    ``rate exp(u - rate exp(u))``."""

    def __init__(self, rate: float):
        self.rate = float(rate)

    def sample(self, n, generator=None):
        return torch.empty(n).exponential_(self.rate, generator=generator)

    def get_constrained(self, u):
        return u.exp()

    def get_unconstrained(self, x):
        return x.log()

    def eval_unconstrained(self, u):
        return math.log(self.rate) - self.rate * u.exp() + u


class ParameterContext:
    """The parameters of the theta-particles: ``values`` is the ``(B, p)`` matrix of UNCONSTRAINED values
    (``stack_parameters(constrained=False)``, inference/context.py), one column per named prior in declaration order."""

    def __init__(self, priors: Dict[str, Prior], device="cuda"):
        self.priors = OrderedDict(priors)
        self.device = device
        self.values: torch.Tensor = None

    @property
    def names(self):
        return list(self.priors)

    def initialize_parameters(self, batch: int, generator=None):
        cols = [p.get_unconstrained(p.sample(batch, generator)) for p in self.priors.values()]
        self.values = torch.stack(cols, dim=1).to(self.device, torch.float32)
        return self

    def stack_parameters(self) -> torch.Tensor:
        return self.values

    def unstack_parameters(self, values: torch.Tensor):
        self.values = values.to(self.device, torch.float32)

    def constrained(self) -> Dict[str, torch.Tensor]:
        return {k: p.get_constrained(self.values[:, i]) for i, (k, p) in enumerate(self.priors.items())}

    def eval_priors(self) -> torch.Tensor:
        """Sum of the unconstrained priors' log-densities, ``(B,)`` (``eval_priors(constrained=False)``)."""
        return sum(p.eval_unconstrained(self.values[:, i]) for i, p in enumerate(self.priors.values()))

    def resample(self, indices: torch.Tensor):
        self.values = self.values[indices]

    def exchange(self, other: "ParameterContext", mask: torch.Tensor):
        self.values = torch.where(mask.unsqueeze(-1), other.values, self.values)

    def make_new(self) -> "ParameterContext":
        return ParameterContext(self.priors, self.device)
