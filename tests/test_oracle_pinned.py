"""Pin oracle/smc_oracle.py: against the golden vectors generated from the unmodified reference, against the
reference's own known-answer test (tests/test_resampling.py:31-47 of the reference) and, where /root/reference
is present (build container), against the reference package itself."""
import numpy as np
import pytest
import torch

from oracle import smc_oracle as O
from oracle.ref_loader import reference_available
from tests.golden_util import filter_cases, load_filter_case, load_residual, load_resampling, model_params, oracle_only_cases


def test_resampling_golden_systematic():
    g = load_resampling()
    names = sorted({k[4:-2] for k in g if k.startswith("sys_") and k.endswith("_W")})
    assert len(names) >= 7
    for name in names:
        W, u, idx = g[f"sys_{name}_W"], g[f"sys_{name}_u"], g[f"sys_{name}_idx"]
        got = O.systematic(torch.from_numpy(W).clone(), normalized=True, u=torch.from_numpy(u))
        assert np.array_equal(got.numpy(), idx), name
        for b in range(W.shape[1]):  # torch-free restatements must agree bit for bit too
            assert np.array_equal(O.systematic_restated(W[:, b], u[b, 0]), idx[:, b]), (name, b)
            if W.shape[0] <= 1000:
                assert np.array_equal(O.systematic_loop(W[:, b], u[b, 0]), idx[:, b]), (name, b)


def test_resampling_golden_multinomial():
    g = load_resampling()
    for name in sorted({k[4:-2] for k in g if k.startswith("mul_") and k.endswith("_W")}):
        W, U, idx = g[f"mul_{name}_W"], g[f"mul_{name}_U"], g[f"mul_{name}_idx"]
        for b in range(W.shape[1]):
            assert np.array_equal(O.multinomial_restated(W[:, b], U[:, b]), idx[:, b]), (name, b)


def test_resampling_golden_residual():
    """``pyfilter.resampling.residual`` (resampling.py:66-105; SURVEY.md 8(f) f4 - oracle only so far, no CUDA path yet): the numpy
    restatement reproduces the reference's indices from the weights and the float64 uniforms its multinomial part drew."""
    g = load_residual()
    names = sorted({k[4:-2] for k in g if k.startswith("res_") and k.endswith("_W")})
    assert len(names) >= 7
    for name in names:
        W, U, idx = g[f"res_{name}_W"], g[f"res_{name}_U"], g[f"res_{name}_idx"]
        assert np.array_equal(O.residual_restated(W, U), idx), name
        n = W.shape[0]
        counts = np.floor(np.float32(n) * W).astype(np.int64)
        assert np.array_equal(np.bincount(idx[: counts.sum()], minlength=n), counts), name   # deterministic part
    with pytest.raises(NotImplementedError):
        O.residual(torch.rand(10, 2))


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container)")
def test_residual_bitwise_vs_reference():
    from oracle.ref_loader import load_reference_resampling

    ref = load_reference_resampling()["resampling"].residual
    for seed, n, std in [(1, 500, 1.0), (2, 4096, 4.0), (3, 17, 0.1)]:
        torch.manual_seed(seed)
        lw = torch.randn(n) * std
        torch.manual_seed(100 + seed)
        a = ref(lw.clone())
        torch.manual_seed(100 + seed)
        b = O.residual(lw.clone())
        assert torch.equal(a, b), seed


def test_normalize_golden():
    g = load_resampling()
    W = O.normalize(torch.from_numpy(g["norm_in"]).clone())
    assert np.array_equal(W.numpy(), g["norm_out"], equal_nan=True)
    assert np.array_equal(O.get_ess(W, True).numpy(), g["norm_ess"], equal_nan=True)


def test_reference_kat_systematic():
    """Same construction as the reference's tests/test_resampling.py:31-47 (seed 123, fp64 weights (10,300), one
    offset per element), checked against an independent two-pointer walk instead of the filterpy loop."""
    torch.random.manual_seed(123)
    weights = O.normalize(torch.randn((10, 300), dtype=torch.float64))
    u = torch.rand(weights.shape)
    inds = O.systematic(weights.moveaxis(0, 1), u=u, normalized=True).moveaxis(0, 1).numpy()
    for i in range(weights.shape[0]):
        w = weights[i].numpy()
        positions = (u[i].numpy() + np.arange(300, dtype=np.float32)) / np.float32(300)
        c = np.cumsum(w)
        c[-1] = 1.0
        j, exp = 0, np.zeros(300, dtype=np.int64)
        for k in range(300):
            while c[j] < positions[k]:
                j += 1
            exp[k] = j
        assert (inds[i] == exp).all()


@pytest.mark.parametrize("tag", filter_cases() + ["oracleonly:" + t for t in oracle_only_cases()])
def test_teacher_forced_steps_match_reference(tag):
    """Every golden file, incl. the ``oracleonly_*`` ones (multi-dimensional LinearGaussianObservations on Lorenz-63, SURVEY.md 8(f)
    f2: the oracle is pinned, the CUDA path is not built yet)."""
    g = load_filter_case(tag.split(":")[1], prefix="oracleonly") if tag.startswith("oracleonly:") else load_filter_case(tag)
    model = O.build_model(g["model"], model_params(g))
    B = g["B"]
    for t in range(g["T"]):
        x, lw = torch.from_numpy(g["x_prev"][t]), torch.from_numpy(g["lw_prev"][t])
        inds = torch.from_numpy(g["inds_prev"][t])
        y = torch.as_tensor(g["y"][t])
        z = torch.from_numpy(g["z"][t])
        if g["proposal"].startswith("nested"):   # the inner normals and the Exp(1) values of the categorical draw
            z = (z, torch.from_numpy(g["Un"][t]))
        if g["alg"] == "gpf":                     # the propagation's draws and the draws of the sample from the Gaussian approximation
            z = (z, torch.from_numpy(g["z2"][t]))
        u = torch.from_numpy(g["u"][t])
        U = g["U"][t]
        U = (U if B else U[:, 0]) if U.size else None
        kw = dict(u=u, resampler=g["resampler"], U=U)
        out = O.STEPS[g["alg"]](model, g["proposal"], x, lw, inds, y, z, **kw)
        # The oracle is the same arithmetic on the same substrate (torch CPU).  On the machine that generated the vectors the
        # agreement is bit for bit; another host CPU vectorises exp / sums differently, so floats may move by a few ulp and,
        # where an ulp of a normalised weight decides a probe, an ancestor may flip.
        gi = torch.from_numpy(g["prev_inds"][t])
        same = out["prev_inds"] == gi
        flips = int((~same).sum())
        assert flips <= max(2, same.numel() // 200), (tag, t, flips)
        for k in ("x", "lw"):
            a, b = out[k], torch.from_numpy(g[k][t]).reshape(out[k].shape)
            m = same if a.dim() == same.dim() else same.unsqueeze(-1).expand_as(a)
            fin = torch.isfinite(b) & m
            assert torch.equal(torch.isfinite(a)[m], torch.isfinite(b)[m]), (tag, t, k)
            assert ((a[fin] - b[fin]).abs() <= 2e-6 + 2e-6 * b[fin].abs()).all(), (tag, t, k, (a - b)[fin].abs().max())
        if flips == 0:
            for k in ("ll", "mean", "var"):
                a, b = out[k].numpy(), g[k][t]
                assert np.allclose(a.reshape(b.shape), b, rtol=5e-6, atol=2e-6, equal_nan=True), (tag, t, k, np.abs(a.reshape(b.shape) - b).max())


def test_kalman_agreement_config1():
    """The reference's own accuracy criterion (tests/filters/test_particle.py:105-111) with a closed-form Kalman
    filter in place of the absent pykalman: median relative deviation < 10 %."""
    torch.manual_seed(123)
    m = O.build_model("lg_ar1")
    _, y = m.simulate(100)
    res = O.batch_filter(m, "sisr", "bootstrap", y, 1500)
    p = O.DEFAULT_PARAMS["lg_ar1"]
    km, _, kll = O.kalman_filter_1d(y.numpy(), p["alpha"], p["beta"], p["sigma"], p["a"], p["b"], p["s"],
                                    p["alpha"], p["sigma"] ** 2 / (1 - p["beta"] ** 2))
    assert abs((kll - res["loglikelihood"].item()) / kll) < 0.1
    means = res["filter_means"][1:, 0].numpy()
    assert np.median(np.abs((km - means) / km)) < 0.1


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("alg,proposal,resampler,bshape,every", [
    ("sisr", "bootstrap", "systematic", (), 1), ("sisr", "linear_gaussian", "multinomial", (3,), 1),
    ("apf", "bootstrap", "systematic", (3,), 1), ("apf", "linear_gaussian", "systematic", (), 1),
    ("sisr", "bootstrap", "systematic", (3,), 3), ("apf", "bootstrap", "systematic", (), 4),
])
def test_free_running_bitwise_vs_reference(alg, proposal, resampler, bshape, every):
    from oracle.ref_loader import load_reference
    from oracle.ref_models import build_reference_model

    load_reference()
    from pyfilter import resampling as RR
    from pyfilter.filters.particle import APF, SISR, proposals as pr

    torch.manual_seed(5)
    m = O.build_model("sine_em")
    _, y = m.simulate(40)
    ssm = build_reference_model("sine_em", O.DEFAULT_PARAMS["sine_em"], observe_every_step=every)
    cls = {"sisr": SISR, "apf": APF}[alg]
    prop = {"bootstrap": pr.Bootstrap, "linear_gaussian": pr.LinearGaussianObservations}[proposal]()
    f = cls(ssm, 300, proposal=prop, resampling=getattr(RR, resampler))
    f.set_batch_shape(torch.Size(bshape))
    torch.manual_seed(11)
    r = f.batch_filter(y, bar=False)
    torch.manual_seed(11)
    o = O.batch_filter(m, alg, proposal, y, 300, bshape, resampler, observe_every_step=every)
    assert torch.equal(r.loglikelihood, o["loglikelihood"])
    assert torch.equal(r.filter_means, o["filter_means"])
    assert torch.equal(r.filter_variance, o["filter_variance"])
    assert torch.equal(r.latest_state.timeseries_state.value, o["x"])
    assert torch.equal(r.latest_state.previous_indices, o["prev_inds"])


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("alg", ["sisr", "apf"])
def test_free_running_bitwise_vs_reference_lorenz_lgo(alg):
    """LinearGaussianObservations with a 3-D state and a 2-D observation (examples/lorenz.ipynb:214) - the oracle's
    restatement run freely against the unmodified reference on the same torch generator: identical bits."""
    from oracle.ref_loader import load_reference
    from oracle.ref_models import build_reference_model

    load_reference()
    from pyfilter import resampling as RR
    from pyfilter.filters.particle import APF, SISR, proposals as pr

    torch.manual_seed(6)
    m = O.build_model("lorenz63_em")
    _, y = m.simulate(25)
    ssm = build_reference_model("lorenz63_em", O.DEFAULT_PARAMS["lorenz63_em"])
    f = {"sisr": SISR, "apf": APF}[alg](ssm, 250, proposal=pr.LinearGaussianObservations(), resampling=RR.systematic)
    torch.manual_seed(12)
    r = f.batch_filter(y, bar=False)
    torch.manual_seed(12)
    o = O.batch_filter(m, alg, "linear_gaussian", y, 250, (), "systematic")
    assert torch.equal(r.loglikelihood, o["loglikelihood"])
    assert torch.equal(r.filter_means, o["filter_means"])
    assert torch.equal(r.latest_state.timeseries_state.value, o["x"])
    assert torch.equal(r.latest_state.previous_indices, o["prev_inds"])


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("name,alg,bshape", [("sine_em", "sisr", ()), ("lorenz63_em", "sisr", ()), ("sine_em", "apf", (3,))])
def test_fixed_lag_smoothing_bitwise_vs_reference(name, alg, bshape):
    """``smooth(states, method="fl")`` (filters/particle/base.py:130-146) - ancestral tracing over the recorded
    states - restated in the oracle and compared bit for bit with the unmodified reference on the same generator."""
    from oracle.ref_loader import load_reference
    from oracle.ref_models import build_reference_model

    load_reference()
    from pyfilter import resampling as RR
    from pyfilter.filters.particle import APF, SISR, proposals as pr

    torch.manual_seed(8)
    m = O.build_model(name)
    _, y = m.simulate(15)
    ssm = build_reference_model(name, O.DEFAULT_PARAMS[name])
    f = {"sisr": SISR, "apf": APF}[alg](ssm, 200, proposal=pr.Bootstrap(), resampling=RR.systematic, record_states=True)
    f.set_batch_shape(torch.Size(bshape))
    torch.manual_seed(13)
    r = f.batch_filter(y, bar=False)
    ref = f.smooth(r.states, method="fl")
    torch.manual_seed(13)
    o = O.batch_filter(m, alg, "bootstrap", y, 200, bshape, "systematic", record_states=True)
    got = O.smooth_fixed_lag(o["states"])
    assert ref.shape == got.shape and torch.equal(ref, got)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("name", ["sine_em", "lorenz63_em"])
def test_ffbs_smoothing_bitwise_vs_reference(name):
    """``smooth(states, method="ffbs")`` (filters/particle/base.py:105-128) restated in the oracle, same generator, same bits."""
    from oracle.ref_loader import load_reference
    from oracle.ref_models import build_reference_model

    load_reference()
    from pyfilter import resampling as RR
    from pyfilter.filters.particle import SISR, proposals as pr

    torch.manual_seed(9)
    m = O.build_model(name)
    _, y = m.simulate(10)
    ssm = build_reference_model(name, O.DEFAULT_PARAMS[name])
    f = SISR(ssm, 150, proposal=pr.Bootstrap(), resampling=RR.systematic, record_states=True)
    torch.manual_seed(14)
    r = f.batch_filter(y, bar=False)
    ref = f.smooth(r.states, method="ffbs")
    torch.manual_seed(14)
    o = O.batch_filter(m, "sisr", "bootstrap", y, 150, (), "systematic", record_states=True)
    got = O.smooth_ffbs(m, o["states"])
    assert ref.shape == got.shape and torch.equal(ref, got)
