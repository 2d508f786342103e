import torch
from torch.distributions import AffineTransform, Distribution, Independent, Normal, TransformedDistribution

from .result import StateSpacePath
from .state import TimeseriesState


def _as_tensor(p):
    return p if isinstance(p, torch.Tensor) else torch.tensor(p, dtype=torch.get_default_dtype())


class StructuralStochasticProcess(object):
    def __init__(self, kernel, parameters, initial_kernel, initial_parameters=None):
        self._kernel = kernel
        self.parameters = tuple(_as_tensor(p) for p in parameters)
        self._initial_kernel = initial_kernel
        self.initial_parameters = (
            self.parameters if initial_parameters is None else tuple(_as_tensor(p) for p in initial_parameters)
        )
        self._event_shape = None

    @property
    def initial_distribution(self) -> Distribution:
        return self._initial_kernel(*self.initial_parameters)

    @property
    def event_shape(self) -> torch.Size:
        return self.initial_distribution.event_shape

    @property
    def n_dim(self) -> int:
        return len(self.event_shape)

    def initial_sample(self, shape=torch.Size([])) -> TimeseriesState:
        dist = self.initial_distribution
        if len(shape) > 0:
            dist = dist.expand(shape)
        return TimeseriesState(0, dist.sample(), self.event_shape)

    def build_density(self, x: TimeseriesState) -> Distribution:
        return self._kernel(x, *self.parameters)

    def propagate(self, x: TimeseriesState, time_increment=1) -> TimeseriesState:
        density = self.build_density(x)
        return x.propagate_from(values=density.sample, time_increment=time_increment)

    def yield_parameters(self):
        return {"parameters": self.parameters, "initial_parameters": self.initial_parameters}


class AffineProcess(StructuralStochasticProcess):
    def __init__(self, mean_scale, parameters, increment_distribution, initial_kernel, initial_parameters=None):
        super().__init__(None, parameters, initial_kernel, initial_parameters)
        self.mean_scale_fun = mean_scale
        self.increment_distribution = increment_distribution

    def mean_scale(self, x: TimeseriesState, parameters=None):
        mean, scale = self.mean_scale_fun(x, *(parameters or self.parameters))
        return torch.broadcast_tensors(mean, scale)

    def build_density(self, x):
        loc, scale = self.mean_scale(x)
        return TransformedDistribution(
            self.increment_distribution, AffineTransform(loc, scale, event_dim=self.n_dim), validate_args=False
        )


class AffineEulerMaruyama(AffineProcess):
    def __init__(self, dynamics, parameters, increment_distribution, dt, initial_kernel, initial_parameters=None):
        self.dt = _as_tensor(dt)
        self._dynamics = dynamics

        def _mean_scale(x, *params):
            drift, diffusion = dynamics(x, *params)
            return x.value + drift * self.dt, diffusion

        super().__init__(_mean_scale, parameters, increment_distribution, initial_kernel, initial_parameters)


class StateSpaceModel(object):
    def __init__(self, hidden, f, parameters, observe_every_step=1):
        self.hidden = hidden
        self._f = f
        self.parameters = tuple(_as_tensor(p) for p in parameters)
        self.observe_every_step = observe_every_step
        self._event_shape = None

    def build_density(self, x: TimeseriesState) -> Distribution:
        return self._f(x, *self.parameters)

    def yield_parameters(self):   # (call site: proposals/utils.py:70-71, ModeFinder.initialize)
        return {"parameters": self.parameters}

    @property
    def event_shape(self):
        if self._event_shape is None:
            self._event_shape = self.build_density(self.hidden.initial_sample()).event_shape
        return self._event_shape

    @property
    def n_dim(self):
        return len(self.event_shape)

    def sample_states(self, steps, x_0=None, samples=torch.Size([])):
        x = x_0 if x_0 is not None else self.hidden.initial_sample(samples)
        xs, ys = [], []
        for _ in range(steps):
            x = self.hidden.propagate(x)
            x.value  # materialise
            xs.append(x)
            ys.append(self.build_density(x).sample())
        return StateSpacePath(xs, ys)


class LinearStateSpaceModel(StateSpaceModel):
    """``y = b + a x + s nu``; parameters are ``(a, s)`` or ``(a, b, s)`` (both occur in the reference:
    tests/filters/models.py:16, proposals/linear.py:48)."""

    def __init__(self, hidden, parameters, event_shape, observe_every_step=1):
        if len(parameters) == 2:
            a, s = parameters
            parameters = (a, torch.zeros(event_shape), s)
        super().__init__(hidden, self._linear, parameters, observe_every_step)
        self._event_shape = event_shape

    def _linear(self, x, a, b, s):
        if self.hidden.n_dim == 0:
            loc = b + a * x.value if len(self._event_shape) == 0 else b + a * x.value.unsqueeze(-1)
        else:
            loc = b + (a @ x.value.unsqueeze(-1)).squeeze(-1) if a.dim() > 1 else b + (a * x.value).sum(-1)
        dist = Normal(loc, s)
        return Independent(dist, 1) if len(self._event_shape) > 0 else dist
