"""GPU tests of user-supplied models (pyfilter_b200.timeseries.compile_user_model, csrc/models.h: SMCB_USER_MODEL_HEADER):
the user's mean_scale / observation density as device code, compiled into its own build of the library."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import smc_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _src(name):
    return open(os.path.join(HERE, "user_models", name)).read()


def test_user_model_reproduces_builtin_bit_for_bit():
    """The SV model written as a user model runs through the same kernels as the built-in one: same seed, same bits - for the resident
    column kernel (2000 particles x 3 columns) and for move_kernel (300,000 particles)."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF

    make = ts.compile_user_model(_src("sv_user.h"), state_dim=1, obs_dim=1)
    torch.manual_seed(1)
    _, y = O.build_model("sv_ar1").simulate(25)
    for N, B in ((2000, 3), (300_000, 0)):
        out = []
        for model in (ts.build("sv_ar1"), make(-1.0, 0.97, 0.2)):
            f = APF(model, N, seed=11)
            if B:
                f.set_batch_shape(torch.Size([B]))
            r = f.batch_filter(y, bar=False)
            out.append((r.loglikelihood.clone(), r.filter_means.clone(), r.latest_state.timeseries_state.value.clone()))
        assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2]), (N, B)


class _Ricker(O.Model):
    """Oracle-side twin of tests/user_models/ricker_user.h (torch CPU)."""

    def __init__(self, r, K, sigma, tau):
        self.r, self.K, self.sigma, self.tau = (torch.as_tensor(v, dtype=torch.float32) for v in (r, K, sigma, tau))
        self.name, self.state_dim, self.obs_dim, self.inc_scale, self.linear_obs = "ricker", 0, 0, 1.0, None

    def mean_scale(self, x):
        return x + self.r * (1.0 - x.exp() / self.K), self.sigma

    def obs_log_prob(self, y, x):
        return O.normal_log_prob(y, x.exp(), self.tau * (0.5 * x).exp())

    def initial_loc_scale(self):
        return self.K.log(), torch.tensor(0.5)


def test_user_model_outside_the_zoo_vs_oracle():
    """A model the zoo does not have (noisy Ricker map, heteroscedastic observation): teacher-forced APF and SISR moves against the
    oracle's step functions evaluated on the oracle-side twin of the model, then a free-running filter's likelihood against the oracle's."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, SISR

    make = ts.compile_user_model(_src("ricker_user.h"), state_dim=1, obs_dim=1)
    pars = (0.8, 20.0, 0.15, 0.3)
    mo = _Ricker(*pars)
    gen = torch.Generator().manual_seed(3)
    N = 20_000
    loc, sc = mo.initial_loc_scale()
    x0 = loc + sc * torch.randn(N, generator=gen)
    lw0 = torch.randn(N, generator=gen) * 0.3
    z = torch.randn(N, generator=gen)
    y = torch.tensor(17.0)
    for alg, cls in (("apf", APF), ("sisr", SISR)):
        f = cls(make(*pars), N, seed=5)
        e = f._get_engine(4)
        e.load_state(x0, lw0, torch.arange(N), 0)
        u = torch.tensor([0.37])
        eps = torch.zeros(e.D, e.B, e.ld)
        eps[0, 0, :N] = z
        e.set_noise(eps.cuda(), u.cuda(), None)
        e.set_observations(y.reshape(1, 1).cuda(), 0)
        e.run(1)
        torch.cuda.synchronize()
        st = e.make_state()
        inds = st.previous_indices.cpu()
        step = O.apf_step if alg == "apf" else O.sisr_step
        kw = dict(force_idx=inds)
        ref = step(mo, "bootstrap", x0, lw0, torch.arange(N), y, z, u, **({"ess_threshold": 0.9} if alg == "sisr" else {}), resampler="systematic", **kw)
        ref0 = step(mo, "bootstrap", x0, lw0, torch.arange(N), y, z, u, **({"ess_threshold": 0.9} if alg == "sisr" else {}), resampler="systematic")
        assert int((ref0["prev_inds"] != inds).sum()) <= N // 200, alg        # ancestors: ulp-level weight differences only (expf vs torch.exp)
        assert torch.allclose(st.timeseries_state.value.cpu(), ref["x"], rtol=0, atol=2e-5), alg
        fin = torch.isfinite(ref["lw"])
        assert ((st.weights.cpu()[fin] - ref["lw"][fin]).abs() <= 2e-4 + 1e-5 * ref["lw"][fin].abs()).all(), alg
        assert np.allclose(st.get_loglikelihood().cpu().numpy(), ref["ll"].numpy(), rtol=1e-4, atol=1e-4), alg
        assert np.allclose(st.get_mean().cpu().numpy().reshape(-1), ref["mean"].numpy().reshape(-1), rtol=1e-4, atol=1e-4), alg
    # free running: simulate on the CPU twin, filter on the device and on the oracle
    torch.manual_seed(9)
    xs, ys, x = [], [], loc + sc * torch.randn(())
    for _ in range(40):
        m, s_ = mo.mean_scale(x)
        x = m + s_ * torch.randn(())
        ys.append(x.exp() + mo.tau * (0.5 * x).exp() * torch.randn(()))
    yv = torch.stack(ys)
    # the likelihood estimate is a random variable on both sides: several independent runs each, means within four standard errors
    dev = np.array([float(APF(make(*pars), 20_000, seed=100 + k).batch_filter(yv, bar=False).loglikelihood) for k in range(8)])
    ora = []
    for k in range(8):
        torch.manual_seed(200 + k)
        ora.append(float(O.batch_filter(mo, "apf", "bootstrap", yv, 20_000)["loglikelihood"]))
    ora = np.array(ora)
    se = math.sqrt(dev.var(ddof=1) / 8 + ora.var(ddof=1) / 8)
    assert abs(dev.mean() - ora.mean()) < 4.0 * se + 0.05, (dev.mean(), ora.mean(), se, dev, ora)


def test_notebook_stochastic_volatility_model_vs_oracle():
    """The model of examples/stochastic-volatility.ipynb:60-83 (Verhulst volatility observed through a sinh-arcsinh transformed normal,
    ``observe_every_step = 1 / dt = 5``, days without a price change masked to NaN) as a user model: teacher-forced APF / SISR moves and a
    free-running filter against the oracle's restatement, then the notebook's SMC2 fit (six parameters with its priors) on a small cloud."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, SISR
    from pyfilter_b200.inference import SMC2, Exponential, LogNormal, Normal

    make = ts.compile_user_model(_src("verhulst_sas_user.h"), state_dim=1, obs_dim=1)
    p = O.DEFAULT_PARAMS["verhulst_sas"]
    pars = (p["kappa"], p["gamma"], p["sigma"], p["mu"], p["nu"], p["tau"], p["dt"])
    mo = O.build_model("verhulst_sas")
    gen = torch.Generator().manual_seed(3)
    N = 20_000
    loc, sc = mo.initial_loc_scale()
    x0 = loc + sc * torch.randn(N, generator=gen)
    lw0 = torch.randn(N, generator=gen) * 0.3
    z = torch.randn(N, generator=gen)
    y = torch.tensor(-1.3)
    for alg, cls in (("apf", APF), ("sisr", SISR)):
        f = cls(make(*pars), N, seed=5)
        e = f._get_engine(4)
        e.load_state(x0, lw0, torch.arange(N), 0)
        u = torch.tensor([0.61])
        eps = torch.zeros(e.D, e.B, e.ld)
        eps[0, 0, :N] = z
        e.set_noise(eps.cuda(), u.cuda(), None)
        e.set_observations(y.reshape(1, 1).cuda(), 0)
        e.run(1)
        torch.cuda.synchronize()
        st = e.make_state()
        inds = st.previous_indices.cpu()
        step = O.apf_step if alg == "apf" else O.sisr_step
        kw = {"ess_threshold": 0.9} if alg == "sisr" else {}
        ref = step(mo, "bootstrap", x0, lw0, torch.arange(N), y, z, u, resampler="systematic", force_idx=inds, **kw)
        ref0 = step(mo, "bootstrap", x0, lw0, torch.arange(N), y, z, u, resampler="systematic", **kw)
        assert int((ref0["prev_inds"] != inds).sum()) <= N // 200, alg
        assert torch.allclose(st.timeseries_state.value.cpu(), ref["x"], rtol=0, atol=2e-5), alg
        fin = torch.isfinite(ref["lw"])
        assert ((st.weights.cpu()[fin] - ref["lw"][fin]).abs() <= 2e-4 + 1e-5 * ref["lw"][fin].abs()).all(), alg
        assert np.allclose(st.get_loglikelihood().cpu().numpy(), ref["ll"].numpy(), rtol=1e-4, atol=1e-4), alg
    # data on the filter's schedule (filters/base.py:204-210): the first observation after one step, then one every five steps
    torch.manual_seed(9)
    x, ys = mo.initial_sample(()), []
    for k in range(40):
        for _ in range(1 if k == 0 else 5):
            x = mo.propagate(x, torch.randn(()))
        ys.append(mo.sample_obs(x, None))
    yv = torch.stack(ys)
    yv[[7, 19, 20]] = float("nan")
    dev = np.array([float(APF(make(*pars, observe_every_step=5), 20_000, seed=100 + k).batch_filter(yv, bar=False).loglikelihood) for k in range(8)])
    ora = []
    for k in range(8):
        torch.manual_seed(200 + k)
        ora.append(float(O.batch_filter(mo, "apf", "bootstrap", yv, 20_000, observe_every_step=5)["loglikelihood"]))
    ora = np.array(ora)
    se = math.sqrt(dev.var(ddof=1) / 8 + ora.var(ddof=1) / 8)
    assert abs(dev.mean() - ora.mean()) < 4.0 * se + 0.05, (dev.mean(), ora.mean(), se, dev, ora)
    # the notebook's fit, small: APF(build_model, 400) under SMC2 with the notebook's priors
    priors = {"kappa": Exponential(10.0), "gamma": LogNormal(0.0, 1.0), "sigma": LogNormal(math.log(0.05), 1.0),
              "mu": Normal(0.0, 0.5), "nu": Normal(0.0, 0.15), "tau": LogNormal(0.0, 0.1)}
    alg = SMC2(lambda q: make(q["kappa"], q["gamma"], q["sigma"], q["mu"], q["nu"], q["tau"], 0.2, observe_every_step=5), priors,
               particles=256, state_particles=400, threshold=0.2, num_steps=2, seed=4, max_observations=64, max_increases=8)
    state = alg.fit(yv)
    ll = state.loglikelihood
    assert state.engine.t == (len(yv) - 1) * 5 + 1 and torch.isfinite(ll).all()
    post = {k: float(v) for k, v in alg.posterior_mean(state).items()}
    assert all(math.isfinite(v) for v in post.values()) and post["gamma"] > 0 and post["tau"] > 0, post
    assert len(state.ess) == len(yv) + 1 and state.ess[-1] > 1.0        # the initial cloud's ESS, then one per observation


def test_user_model_rejects_unsupported_pairings():
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, proposals

    make = ts.compile_user_model(_src("sv_user.h"), state_dim=1, obs_dim=1)
    with pytest.raises(ValueError):
        APF(make(-1.0, 0.97, 0.2), 100, proposal=proposals.LinearGaussianObservations()).initialize()
    with pytest.raises(ValueError):
        APF(make(-1.0, 0.97), 100).initialize()      # wrong number of parameters
