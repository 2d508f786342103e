"""CPU: the theta-level host arithmetic of pyfilter_b200.inference (priors, parameter context, proposal fit, acceptance rule, jittering
kernels) is plain torch and runs without a device: checked against the oracle (oracle/smc2_oracle.py) and the reference-generated golden
vectors (tests/golden/smc2_theta.npz).  The filters behind SMC2 / NESS need the GPU (tests/test_gpu_smc2.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import smc2_oracle as S
from pyfilter_b200.inference import Exponential, LogNormal, Normal, ParameterContext
from pyfilter_b200.inference import ness as NS
from pyfilter_b200.inference import smc2 as M

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "smc2_theta.npz")


def _g():
    z = np.load(GOLDEN)
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.mark.parametrize("name", ["b64_p2", "b1024_p2", "b256_p3"])
def test_proposal_fit_and_log_prob_vs_reference_golden(name):
    g = _g()
    mean, tril = M.construct_mvn(g[f"{name}_x"], g[f"{name}_W"], scale=1.1)
    assert torch.allclose(mean, g[f"{name}_mean"], rtol=1e-6, atol=1e-7) and torch.allclose(tril, g[f"{name}_tril"], rtol=1e-5, atol=1e-6)
    lp = M.mvn_log_prob(mean, tril, g[f"{name}_pts"])
    assert torch.allclose(lp, g[f"{name}_lp"], rtol=1e-5, atol=1e-4)
    eps = torch.randn(g[f"{name}_x"].shape, generator=torch.Generator().manual_seed(1))
    ref = torch.distributions.MultivariateNormal(mean, scale_tril=tril)
    assert torch.allclose(M.mvn_sample(mean, tril, eps), ref.loc + (ref.scale_tril @ eps.unsqueeze(-1)).squeeze(-1), atol=1e-5)


def test_priors_context_and_acceptance():
    g = _g()
    u = g["prior_u"]
    assert torch.allclose(Normal(0.0, 1.0).eval_unconstrained(u), g["prior_normal"], rtol=1e-5, atol=1e-5)
    assert torch.allclose(LogNormal(0.0, 0.5).eval_unconstrained(u), g["prior_lognormal"], rtol=1e-5, atol=1e-5)
    ctx = ParameterContext({"gamma": Normal(0.0, 1.0), "sigma": LogNormal(0.0, 0.5)}, device="cpu").initialize_parameters(300, torch.Generator().manual_seed(3))
    c = ctx.constrained()
    assert (c["sigma"] > 0).all() and torch.allclose(c["sigma"].log(), ctx.values[:, 1]) and torch.equal(c["gamma"], ctx.values[:, 0])
    other = ctx.make_new().initialize_parameters(300, torch.Generator().manual_seed(4))
    mask = torch.rand(300, generator=torch.Generator().manual_seed(5)) < 0.4
    before = ctx.values.clone()
    ctx.exchange(other, mask)
    assert torch.equal(ctx.values[mask], other.values[mask]) and torch.equal(ctx.values[~mask], before[~mask])
    idx = torch.randint(0, 300, (300,), generator=torch.Generator().manual_seed(6))
    v = ctx.values.clone()
    ctx.resample(idx)
    assert torch.equal(ctx.values, v[idx])
    gen = torch.Generator().manual_seed(7)
    a, b, c3 = (torch.randn(300, generator=gen) for _ in range(3))
    uu = torch.rand(300, generator=gen)
    assert torch.equal(M.pmmh_accept(a, b, c3, uu), uu.log() < (a + b + c3))


@pytest.mark.parametrize("kind", ["shrinking", "nonshrinking", "liuwest", "constant"])
def test_jitter_kernels_vs_reference_golden_and_oracle(kind):
    g = _g()
    x, W, idx = g["b1024_p2_x"], g["b1024_p2_W"], g["b1024_p2_idx"]
    k = {"shrinking": NS.ShrinkingKernel(), "nonshrinking": NS.NonShrinkingKernel(), "liuwest": NS.LiuWestShrinkage(0.98),
         "constant": NS.ConstantKernel(0.1)}[kind]
    loc, sc = k.fit(x, W, idx)
    oloc, osc = S.jitter_fit(kind, x.clone(), W.clone(), idx)
    assert torch.allclose(loc, g[f"jit_{kind}_loc"], rtol=1e-5, atol=1e-6) and torch.allclose(loc, oloc, rtol=1e-5, atol=1e-6)
    # the weighted quartiles: a triangular product here, a sequential sum in the reference - a quartile may land on the neighbouring order statistic
    assert torch.allclose(torch.as_tensor(sc).float(), g[f"jit_{kind}_scale"], rtol=5e-3, atol=1e-7)
    assert torch.allclose(torch.as_tensor(sc).float(), torch.as_tensor(osc).float(), rtol=5e-3, atol=1e-7)


def test_exponential_prior_in_the_unconstrained_space():
    """``Exponential(rate)`` (examples/stochastic-volatility.ipynb:75): the prior of ``log x`` is what torch's
    ``TransformedDistribution(Exponential, biject_to(positive).inv)`` evaluates (inference/prior.py:33-44, 81-90)."""
    import torch.distributions as D

    u = torch.linspace(-6.0, 2.0, 41)
    ref = D.TransformedDistribution(D.Exponential(10.0), D.biject_to(D.constraints.positive).inv).log_prob(u)
    p = Exponential(10.0)
    assert torch.allclose(p.eval_unconstrained(u), ref, rtol=1e-5, atol=1e-5)
    x = p.sample(20_000, torch.Generator().manual_seed(1))
    assert (x > 0).all() and abs(float(x.mean()) - 0.1) < 0.005
    assert torch.allclose(p.get_constrained(p.get_unconstrained(x)), x, rtol=1e-5)


def test_smc2_move_schedule_is_the_filters():
    """``observe_every_step`` (filters/base.py:204-210): the move that consumes observation k is the one the filter's own expansion of
    the observations puts it on - SMC2 / NESS advance their resident filters by that many moves per observation."""
    import types

    from pyfilter_b200.filters.particle.base import ParticleFilter
    from pyfilter_b200.inference.smc2 import SMC2

    for every in (1, 3, 5):
        alg = SMC2.__new__(SMC2)
        alg._oes = every
        dummy = types.SimpleNamespace(_model=types.SimpleNamespace(observe_every_step=every))
        y = torch.arange(7.0)
        y_moves, observed = ParticleFilter._expand_observations(dummy, y, 0)
        assert observed == [alg._moves(k + 1) - 1 for k in range(7)]
        assert int(y_moves.shape[0]) == alg._moves(7) and alg._moves(0) == 0
        assert torch.equal(y_moves[observed, 0], y) and torch.isnan(y_moves).sum() == y_moves.shape[0] - 7
