"""Reference-side (stand-in ``stochproc``) definitions of the BASELINE.json state-space models.

TEST INFRASTRUCTURE, build container only: these build the *user model objects* a pyfilter user would write
(Python callables on top of stochproc), mirroring reference README.md:44-67, tests/filters/models.py:12-16 and
examples/lorenz.ipynb:53-117, so that the unmodified reference filters can be run to generate golden vectors.
"""
import math

import torch


def build_reference_model(name: str, params: dict, observe_every_step: int = 1):
    from pyro.distributions import Normal  # stand-in
    from stochproc import timeseries as ts  # stand-in

    def t(v):
        return torch.as_tensor(v, dtype=torch.float32)

    if name == "lg_ar1":
        ar = ts.models.AR(t(params["alpha"]), t(params["beta"]), t(params["sigma"]))
        return ts.LinearStateSpaceModel(ar, (t(params["a"]), t(params["b"]), t(params["s"])), torch.Size([]))

    if name == "sine_em":
        dt = params["dt"]

        def f(x, gamma, sigma):
            return torch.sin(x.value - gamma), sigma

        def initial_kernel(gamma, sigma):
            return Normal(torch.zeros_like(gamma), torch.ones_like(gamma))

        hidden = ts.AffineEulerMaruyama(
            f, (t(params["gamma"]), t(params["sigma"])), Normal(loc=0.0, scale=math.sqrt(dt)), dt=dt,
            initial_kernel=initial_kernel,
        )
        return ts.LinearStateSpaceModel(hidden, (t(params["a"]), t(params["b"]), t(params["s"])), torch.Size([]),
                                        observe_every_step=observe_every_step)

    if name == "sv_ar1":
        def mean_scale(x, mu, phi, sigma_v):
            return mu + phi * (x.value - mu), sigma_v

        def initial_kernel(mu, phi, sigma_v):
            return Normal(mu, sigma_v / (1.0 - phi**2.0).sqrt())

        hidden = ts.AffineProcess(
            mean_scale, (t(params["mu"]), t(params["phi"]), t(params["sigma_v"])), Normal(0.0, 1.0), initial_kernel
        )

        def build_observation(x):
            return Normal(loc=0.0, scale=(x.value / 2.0).exp())

        return ts.StateSpaceModel(hidden, build_observation, ())

    if name == "lorenz63_em":
        dt = params["dt"]

        def f(x, s, r, b, sigma):
            x_t = -s * (x.value[..., 0] - x.value[..., 1])
            y_t = r * x.value[..., 0] - x.value[..., 1] - x.value[..., 0] * x.value[..., 2]
            z_t = x.value[..., 0] * x.value[..., 1] - b * x.value[..., 2]
            return torch.stack((x_t, y_t, z_t), dim=-1), (sigma.unsqueeze(-1) if sigma.dim() > 0 else sigma)

        def initial_kernel(x0, s0):
            return Normal(loc=x0, scale=s0).to_event(1)

        mean = torch.tensor([-5.91652, -5.52332, 24.5723])
        scale = math.sqrt(10) * torch.ones(3)
        increment_dist = Normal(loc=0.0, scale=math.sqrt(dt)).expand(mean.shape).to_event(1)
        hidden = ts.AffineEulerMaruyama(
            f, (t(params["s"]), t(params["r"]), t(params["b"]), t(params["sigma"])), increment_dist, dt=dt,
            initial_kernel=initial_kernel, initial_parameters=(mean, scale),
        )
        a = params["obs_a"]
        mat = torch.tensor([[a, 0.0, 0.0], [0.0, 0.0, a]])
        s = t(params["obs_s"]).reshape(-1)[:1] if t(params["obs_s"]).dim() else t(params["obs_s"]).unsqueeze(-1)
        offset = torch.zeros_like(s)
        return ts.LinearStateSpaceModel(hidden, (mat, offset, s), torch.Size([2]))

    raise ValueError(name)
