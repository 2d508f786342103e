"""GPU parity tests of the callers and plug-ins either side of the fused move: the proposal plug-in methods,
the split predict / correct step, predict_path, batched_gather, fixed-lag and FFBS smoothing, the residual resampler and the
theta-level column operations - each against the oracle (oracle/smc_oracle.py) or the reference-generated golden vectors."""
import numpy as np
import pytest
import torch

from oracle import smc_oracle as O
from tests.golden_util import load_residual

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pf():
    import pyfilter_b200 as pf

    assert torch.cuda.is_available(), "these tests need the B200"
    return pf


def _proposal(pf, name):
    from pyfilter_b200.filters.particle import proposals

    if name.startswith("linearized"):
        parts = name.split(":")
        return proposals.Linearized(n_steps=int(parts[1]), alpha=float(parts[2]), use_second_order=bool(int(parts[3])))
    return {"bootstrap": proposals.Bootstrap, "linear_gaussian": proposals.LinearGaussianObservations}[name]()


CASES = [("lg_ar1", "bootstrap"), ("lg_ar1", "linear_gaussian"), ("sine_em", "linear_gaussian"), ("sv_ar1", "bootstrap"),
         ("lorenz63_em", "bootstrap"), ("lorenz63_em", "linear_gaussian"),
         ("sv_ar1", "linearized:1:0.0001:0"), ("sv_ar1", "linearized:4:0.0001:1"), ("sine_em", "linearized:3:0.01:0"),
         ("lg_ar1", "linearized:2:0.0001:1"), ("lorenz63_em", "linearized:2:0.0001:1"), ("lorenz63_em", "linearized:2:0.001:0")]


@pytest.mark.parametrize("B", [0, 3])
@pytest.mark.parametrize("name,prop", CASES)
def test_proposal_plugin_methods_vs_oracle(pf, name, prop, B):
    """``Proposal.pre_weight`` / ``Proposal.sample_and_weight`` (proposals/base.py:52-85, bootstrap.py:10-14, linear.py:38-86) as
    stand-alone device passes, with injected N(0,1) draws, against the oracle's restatement of the same lines."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle.state import ParticleFilterPrediction

    gen = torch.Generator().manual_seed(11)
    mo = O.build_model(name)
    N = 1777
    shape = (N, B) if B else (N,)
    ev = (mo.state_dim,) if mo.state_dim else ()
    loc, scale = mo.initial_loc_scale()
    x = loc + scale * torch.randn(shape + ev, generator=gen)
    z = torch.randn(shape + ev, generator=gen)
    _, ysim = mo.simulate(3)
    y = ysim[2].float()
    p = _proposal(pf, prop).set_model(ts.build(name))
    xs = ts.TimeseriesState(4, x.cuda(), torch.Size(ev))
    # pre_weight
    got = p.pre_weight(y, xs).cpu()
    ref = O.PROPOSALS[prop][1](mo, y, x)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, rtol=0, atol=3e-5 + 4e-6 * float(ref.abs().max())), (got - ref).abs().max()
    # sample_and_weight
    pred = ParticleFilterPrediction(xs, torch.zeros(shape).cuda(), torch.full(shape, 1.0 / N).cuda(), None)
    new, w = p.sample_and_weight(y, pred, eps=z.cuda())
    rx, rw = O.PROPOSALS[prop][0](mo, y, x, z)
    assert int(new.time_index) == 5 and new.value.shape == rx.shape
    assert torch.allclose(new.value.cpu(), rx, rtol=0, atol=2e-6 * max(1.0, float(rx.abs().max()))), (new.value.cpu() - rx).abs().max()
    assert torch.allclose(w.cpu(), rw, rtol=0, atol=3e-5 + 4e-6 * float(rw.abs().max())), (w.cpu() - rw).abs().max()
    # Philox draws when nothing is injected: same call twice on a fresh proposal differs from the injected result, is finite, reproducible
    a1, w1 = p.sample_and_weight(y, pred)
    a2, w2 = p.sample_and_weight(y, pred)
    assert torch.equal(a1.value, a2.value) and torch.isfinite(w1).all()
    assert not torch.equal(a1.value.cpu(), rx)


@pytest.mark.parametrize("name", ["lg_ar1", "lorenz63_em"])
def test_nested_proposal_stand_alone_pass(pf, name):
    """``NestedProposal.sample_and_weight`` (proposals/nested.py:27-47) as a stand-alone pass: exact against the oracle when the inner
    normals and the Exp(1) values of the categorical draw are injected; with the inner normals injected and the library's own pick, every
    new particle is one of its inner samples and the picks follow the soft-max probabilities; with the library's own draws throughout,
    the law of the result (shift and spread of the chosen normal, mean weight) is the oracle's."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import proposals
    from pyfilter_b200.filters.particle.state import ParticleFilterPrediction

    M, N = 50, 4096
    torch.manual_seed(5)
    mo = O.build_model(name)
    gen = torch.Generator().manual_seed(3)
    ev = (mo.state_dim,) if mo.state_dim else ()
    loc, scale = mo.initial_loc_scale()
    if name == "lorenz63_em":   # a tight cloud and an observation it can explain (the inner samples spread by 0.1, the noise is 0.32)
        x = loc + 0.1 * torch.randn((N,) + ev, generator=gen)
        y = mo.obs_loc(mo.mean_scale(loc)[0]).float() + 0.2
    else:
        x = loc + scale * torch.randn((N,) + ev, generator=gen)
        y = mo.simulate(3)[1][2].float()
    p = proposals.NestedProposal(M).set_model(ts.build(name))
    xs = ts.TimeseriesState(4, x.cuda(), torch.Size(ev))
    pred = ParticleFilterPrediction(xs, torch.zeros(N).cuda(), torch.full((N,), 1.0 / N).cuda(), None)
    zs = torch.randn((M, N) + ev, generator=gen)
    E = torch.empty(N, M).exponential_(1, generator=gen)
    rx, rw = O.nested_sample_and_weight(mo, y, x, (zs, E), M)
    e = p._engine_for(xs.value)
    e.set_nested_noise(zs, E)
    new, w = p.sample_and_weight(y, pred)
    xtol = 2e-6 * max(1.0, float(rx.abs().max()))
    bad = (new.value.cpu() - rx).abs() > xtol
    assert int((bad if bad.dim() == 1 else bad.any(-1)).sum()) <= 4          # near-ties of the race only
    fin = rw > -1e30
    assert ((w.cpu() - rw)[fin].abs() <= 3e-5 + 4e-6 * rw[fin].abs()).all()
    # the library's pick on injected inner normals
    e.set_nested_noise(zs, None)
    new, w2 = p.sample_and_weight(y, pred)
    mean, sc = mo.mean_scale(x)
    samples = mean + sc * (zs * mo.inc_scale)
    nv = new.value.cpu()
    dist = (samples - nv).abs() if not ev else (samples - nv).abs().amax(-1)
    best = dist.argmin(0)
    assert float(dist.gather(0, best[None]).max()) <= xtol
    assert torch.equal(w2, w)
    lp = mo.obs_log_prob(y, samples).nan_to_num(-float("inf"), -float("inf"))
    probs = lp.softmax(0)
    c = probs.double().cumsum(0)
    hi = c.gather(0, best[None])[0]
    mid = hi - probs.double().gather(0, best[None])[0] / 2                  # the uniform that picked lies around here: U(0, 1) overall
    assert abs(float(mid.mean()) - 0.5) < 0.03 and abs(float(mid.var()) - 1 / 12) < 0.01
    # the library's own draws throughout
    e.set_nested_noise(None, None)
    new, w3 = p.sample_and_weight(y, pred)
    zb, zr = (new.value.cpu() - mean) / (sc * mo.inc_scale), (rx - mean) / (sc * mo.inc_scale)
    assert (zb.mean(0) - zr.mean(0)).abs().max() < 0.1 and (zb.var(0) - zr.var(0)).abs().max() < 0.2, (zb.mean(0), zr.mean(0), zb.var(0), zr.var(0))
    f3 = (w3.cpu() > -1e30) & fin
    assert abs(float(f3.float().mean()) - float(fin.float().mean())) < 0.05
    assert abs(float(w3.cpu()[f3].mean()) - float(rw[f3].mean())) < 0.05 * max(1.0, abs(float(rw[f3].mean())))


@pytest.mark.parametrize("alg", ["sisr", "apf"])
@pytest.mark.parametrize("B", [0, 4])
def test_split_predict_correct_matches_oracle_statistics(pf, alg, B):
    """``filter.predict`` + ``filter.correct`` (filters/base.py:160-186; sisr.py:14-56, apf.py:16-46) driven like the reference's own
    loop: free running on the linear-Gaussian model, filter means against the closed-form Kalman filter (the reference's accuracy
    criterion, tests/filters/test_particle.py:105-111) and the structure of the prediction objects."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, SISR

    torch.manual_seed(123)
    m = O.build_model("lg_ar1")
    _, y = m.simulate(60)
    p = O.DEFAULT_PARAMS["lg_ar1"]
    km, _, _ = O.kalman_filter_1d(y.numpy(), p["alpha"], p["beta"], p["sigma"], p["a"], p["b"], p["s"], p["alpha"],
                                  p["sigma"] ** 2 / (1 - p["beta"] ** 2))
    N = 3000
    f = {"sisr": SISR, "apf": APF}[alg](ts.build("lg_ar1"), N, seed=5)
    if B:
        f.set_batch_shape(torch.Size([B]))
    state = f.initialize().detach_copy()
    means, resampled_any = [], False
    for t in range(60):
        pred = f.predict(state)
        shape = (N, B) if B else (N,)
        assert pred.weights.shape == shape and pred.normalized_weights.shape == shape and pred.indices.shape == shape
        assert torch.allclose(pred.normalized_weights.sum(0).cpu(), torch.ones(B if B else ()), atol=1e-4)
        if alg == "sisr":
            rs = (pred.weights == 0).all(0)
            resampled_any = resampled_any or bool(rs.any())
            # resampled columns carry uniform weights and sorted systematic ancestors (sisr.py:29-35)
            if bool(rs.any()):
                nw = pred.normalized_weights if B else pred.normalized_weights.unsqueeze(-1)
                idx = pred.indices if B else pred.indices.unsqueeze(-1)
                cols = rs.reshape(-1).nonzero().flatten()
                assert torch.allclose(nw[:, cols], torch.full_like(nw[:, cols], 1.0 / N))
                assert bool((idx[1:, cols] >= idx[:-1, cols]).all())
        else:
            assert torch.equal(pred.indices.reshape(N, -1)[:, 0].cpu(), torch.arange(N))
        state = f.correct(y[t], pred)
        assert int(state.timeseries_state.time_index) == t + 1
        means.append(state.get_mean().cpu().reshape(-1))
    got = torch.stack(means).numpy()
    dev = np.median(np.abs((got - km[:, None]) / km[:, None]), axis=0)
    assert (dev < 0.1).all(), dev
    if alg == "sisr":
        assert resampled_any


def test_split_step_equals_fused_step_on_same_noise(pf):
    """APF: ``correct(y, predict(state))`` is the same fused move as ``filter(y, state)`` from the same state - same seed and move
    index give the same Philox counters, hence the same particles."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF

    torch.manual_seed(3)
    _, y = O.build_model("sv_ar1").simulate(4)
    f = APF(ts.build("sv_ar1"), 20_000, seed=9)
    s0 = f.initialize().detach_copy()
    a = f.correct(y[0], f.predict(s0))
    b = f.filter(y[0], s0).detach_copy()
    assert torch.equal(a.timeseries_state.value, b.timeseries_state.value)
    assert torch.equal(a.weights, b.weights) and torch.equal(a.previous_indices, b.previous_indices)
    assert torch.allclose(a.get_loglikelihood(), b.get_loglikelihood())


@pytest.mark.parametrize("name,B", [("lg_ar1", 0), ("lorenz63_em", 0), ("sine_em", 3), ("sv_ar1", 2)])
def test_predict_path_shapes_and_law(pf, name, B):
    """``latest_state.predict_path(model, num_steps)`` (particle/state.py:173-174; reference test tests/filters/test_particle.py:126-135):
    shapes, and one-step-ahead moments of the simulated paths against the model's transition / observation laws."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF

    torch.manual_seed(1)
    mo = O.build_model(name)
    _, y = mo.simulate(10)
    model = ts.build(name)
    N = 40_000
    f = APF(model, N, seed=3)
    if B:
        f.set_batch_shape(torch.Size([B]))
    res = f.batch_filter(y, bar=False)
    steps = 6
    path = res.latest_state.predict_path(model, steps)
    assert len(path.get_paths()) == 2
    xp, yp = path.get_paths()
    assert xp.shape == torch.Size([steps, *f.particles, *f.ssm.hidden.event_shape])
    assert yp.shape == torch.Size([steps, *f.particles, *model.event_shape])
    x0 = res.latest_state.timeseries_state.value.cpu()
    loc, scale = mo.mean_scale(x0)
    std = scale * mo.inc_scale
    zs = ((xp[0].cpu() - loc) / std)
    assert abs(float(zs.mean())) < 0.02 and abs(float(zs.std()) - 1.0) < 0.02
    assert torch.isfinite(xp).all() and torch.isfinite(yp).all()
    # observations scatter around their location with the model's scale
    if name != "sv_ar1":
        ol = mo.obs_loc(xp[0].cpu()) if hasattr(mo, "obs_loc") else None
        if ol is not None:
            r = (yp[0].cpu() - ol)
            assert abs(float(r.mean())) < 0.02 * max(1.0, float(r.std()))


def test_batched_gather_and_fixed_lag_smoothing_vs_oracle(pf):
    """``batched_gather`` (filters/utils.py:4-21) and ``smooth(states, method="fl")`` (filters/particle/base.py:130-146): pure index
    work - bit-exact against the oracle's restatement on the device's own recorded states."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import SISR
    from pyfilter_b200.filters.utils import batched_gather

    gen = torch.Generator().manual_seed(2)
    x = torch.randn(500, 3, 2, generator=gen)
    idx = torch.randint(0, 500, (500, 3), generator=gen)
    got = batched_gather(x.cuda(), idx.cuda()).cpu()
    assert torch.equal(got, O._gather0(x, idx))
    x1 = torch.randn(1000, generator=gen)
    i1 = torch.randint(0, 1000, (1000,), generator=gen)
    assert torch.equal(batched_gather(x1.cuda(), i1.cuda()).cpu(), x1[i1])
    with pytest.raises(ValueError):
        batched_gather(x1.cuda(), (i1 + 1000).cuda())
    for name, B in (("lg_ar1", 0), ("lorenz63_em", 0), ("sine_em", 3)):
        torch.manual_seed(4)
        _, y = O.build_model(name).simulate(12)
        f = SISR(ts.build(name), 2000, seed=8, record_states=True)
        if B:
            f.set_batch_shape(torch.Size([B]))
        res = f.batch_filter(y, bar=False)
        assert len(res.states) == 13
        sm = f.smooth(res.states, method="fl").cpu()
        ref = O.smooth_fixed_lag([(s.timeseries_state.value.cpu(), s.weights.cpu(), s.previous_indices.cpu()) for s in res.states])
        assert sm.shape == ref.shape and torch.equal(sm, ref), name
    with pytest.raises(NotImplementedError):
        f.smooth(res.states, method="nope")


@pytest.mark.parametrize("name", ["lg_ar1", "lorenz63_em"])
def test_ffbs_backward_sampling_vs_oracle(pf, name):
    """``smooth(states, method="ffbs")`` (filters/particle/base.py:105-128): with injected uniforms the drawn predecessors agree with the
    inversion of the oracle's float64 cumulative probabilities except where a uniform falls within float32 rounding of a boundary; the
    free-running smoothed means agree with the oracle's own FFBS within Monte-Carlo error."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import SISR
    from pyfilter_b200.filters.particle.smoothing import ffbs

    torch.manual_seed(6)
    mo = O.build_model(name)
    _, y = mo.simulate(8)
    N = 1500
    f = SISR(ts.build(name), N, seed=4, record_states=True)
    res = f.batch_filter(y, bar=False)
    states = res.states
    cpu_states = [(s.timeseries_state.value.cpu(), s.weights.cpu(), s.previous_indices.cpu()) for s in states]
    gen = torch.Generator().manual_seed(1)
    U = [torch.rand(N, dtype=torch.float64, generator=gen) for _ in range(len(states) - 1)]
    torch.manual_seed(10)
    sm = ffbs(f, states, uniforms=U).cpu()
    assert sm.shape == torch.Size([len(states), N, *f.ssm.hidden.event_shape])
    # replay: walk backwards with the device's own smoothed particles as the conditioning values
    total, flips = 0, 0
    for k, (x_s, lw_s, _) in enumerate(reversed(cpu_states[:-1])):
        later = sm[len(states) - 1 - k]
        idx_ref = O.ffbs_backward_indices(mo, x_s, lw_s, later, U[k])
        picked = O._gather0(x_s, idx_ref)
        got = sm[len(states) - 2 - k]
        same = (picked == got).reshape(N, -1).all(-1)
        total += N
        flips += int((~same).sum())
    assert flips <= max(3, total // 500), (flips, total)
    # free running against the oracle's FFBS (its own draws): smoothed means per time step
    torch.manual_seed(12)
    ref = O.smooth_ffbs(mo, cpu_states)
    sm2 = f.smooth(states, method="ffbs").cpu()
    sd = ref.reshape(len(states), N, -1).std(1)
    err = (sm2.reshape(len(states), N, -1).mean(1) - ref.reshape(len(states), N, -1).mean(1)).abs()
    assert (err <= 6.0 * sd / np.sqrt(N) + 1e-3).all(), (err / (sd / np.sqrt(N))).max()


def test_residual_resampler_golden_and_restated(pf):
    """``pyfilter.resampling.residual`` (resampling.py:68-105): bit-exact against the reference-generated golden vectors
    (tests/golden/residual.npz: weights, the float64 uniforms of the multinomial part, indices) and against the numpy restatement."""
    g = load_residual()
    names = sorted({k[4:-2] for k in g if k.endswith("_W")})
    assert len(names) >= 5
    for name in names:
        W, U, idx = g[f"res_{name}_W"], g[f"res_{name}_U"], g[f"res_{name}_idx"]
        got = pf.resampling.residual(torch.from_numpy(W).cuda(), normalized=True, U=torch.from_numpy(U).cuda() if U.size else None)
        assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), idx), name
    gen = torch.Generator().manual_seed(5)
    n = 100_000
    W = O.normalize(torch.randn(n, generator=gen) * 2.5)
    U = torch.rand(n, dtype=torch.float64, generator=gen)
    got = pf.resampling.residual(W.cuda(), normalized=True, U=U.cuda()).cpu().numpy()
    assert np.array_equal(got, O.residual_restated(W.numpy(), U.numpy()))
    # log-weights in, Philox draws: counts of the deterministic part are honoured
    lw = torch.randn(5000, generator=gen)
    out = pf.resampling.residual(lw.cuda()).cpu()
    Wn = O.normalize(lw.clone())
    counts = torch.bincount(out, minlength=5000)
    assert bool((counts >= (5000 * Wn).floor().long()).all()) and int(counts.sum()) == 5000
    with pytest.raises(NotImplementedError):
        pf.resampling.residual(torch.zeros(8, 2).cuda())


def test_any_resampler_runs_through_the_split_step(pf):
    """``resampling=`` takes any callable in the reference (filters/particle/base.py:23,43).  systematic / multinomial live inside the
    fused move; ``residual`` or a user's function is called between the stand-alone passes of ``SISR.predict`` and the fused correction
    (the reference's own loop, filters/base.py:201-221) - same accuracy criterion as the reference's tests, NaN observations included."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, SISR

    torch.manual_seed(123)
    mo = O.build_model("lg_ar1")
    _, y = mo.simulate(60)
    y = y.float()
    y[[11, 30]] = float("nan")
    p = O.DEFAULT_PARAMS["lg_ar1"]
    _, _, kll = O.kalman_filter_1d(y.numpy(), p["alpha"], p["beta"], p["sigma"], p["a"], p["b"], p["s"], p["alpha"],
                                   p["sigma"] ** 2 / (1 - p["beta"] ** 2))
    calls = []

    def mine(w, normalized=False):
        calls.append(tuple(w.shape))
        return pf.resampling.systematic(w, normalized=normalized)

    fused = SISR(ts.build("lg_ar1"), 2000, seed=5).batch_filter(y, bar=False)
    for res in (pf.resampling.residual, mine):
        f = SISR(ts.build("lg_ar1"), 2000, resampling=res, seed=5)
        r = f.batch_filter(y, bar=False)
        assert r.filter_means.shape == fused.filter_means.shape and len(r.states) == 1
        assert abs(float(r.loglikelihood) - float(fused.loglikelihood)) < 0.05 * abs(float(fused.loglikelihood))
        assert abs(float(r.loglikelihood) - kll) < 0.1 * abs(kll)          # tests/filters/test_particle.py:105
        assert int(r.latest_state.timeseries_state.time_index) == 60
        g = f.copy()
        assert g._resampler is res
    assert calls and all(c == (2000,) for c in calls)
    # batched filters hand the callable the columns that resample
    calls.clear()
    fb = SISR(ts.build("lg_ar1"), 1000, resampling=mine, seed=6)
    fb.set_batch_shape(torch.Size([3]))
    rb = fb.batch_filter(y[:20], bar=False)
    assert rb.loglikelihood.shape == (3,) and torch.isfinite(rb.loglikelihood).all() and all(c[0] == 1000 and len(c) == 2 for c in calls)
    with pytest.raises(NotImplementedError):
        APF(ts.build("lg_ar1"), 100, resampling=pf.resampling.residual)


def test_theta_level_column_operations(pf):
    """``FilterResult.resample`` / ``.exchange`` on the resident state (filters/result.py:76-117, particle/state.py:150-168): the device
    permutation / masked copy of columns against torch indexing of the same tensors, and the filter keeps running afterwards exactly as
    a filter that was handed the permuted state."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF

    torch.manual_seed(2)
    B, N = 6, 5000
    _, y = O.build_model("sine_em").simulate(12)
    yd = y.float().reshape(12, -1).cuda().contiguous()
    gam = torch.linspace(-0.5, 0.5, B)
    def make(seed):
        f = APF(ts.build("sine_em", gamma=gam), N, seed=seed)
        f.set_batch_shape(torch.Size([B]))
        e = f._get_engine(16)
        e.initialize(); e.set_observations(yd, 0); e.run(6)
        return f, e
    f1, e1 = make(1)
    f2, e2 = make(2)
    torch.cuda.synchronize()
    snap = lambda e: dict(x=e.x_view().clone(), lw=e.logw_view().clone(), pi=e.prev_inds().clone(), ll=e.raw(6, (e.B,)).clone(),
                          mean=e.raw(3, (e.B, e.D)).clone(), hist=e.history(7)[0].clone())
    a, b = snap(e1), snap(e2)
    idx = torch.tensor([3, 3, 0, 5, 1, 1])
    e1.resample_columns(idx.cuda(), entire_history=True)
    torch.cuda.synchronize()
    c = snap(e1)
    assert torch.equal(c["x"], a["x"][:, idx]) and torch.equal(c["lw"], a["lw"][:, idx]) and torch.equal(c["pi"], a["pi"][:, idx])
    assert torch.equal(c["ll"], a["ll"][idx]) and torch.equal(c["mean"], a["mean"][idx]) and torch.equal(c["hist"], a["hist"][:, idx])
    mask = torch.tensor([True, False, False, True, True, False])
    e1.exchange_columns(e2, mask.cuda())
    torch.cuda.synchronize()
    d = snap(e1)
    for k in ("x", "lw", "pi"):
        exp = c[k].clone(); exp[:, mask] = b[k][:, mask]
        assert torch.equal(d[k], exp), k
    exp = c["ll"].clone(); exp[mask] = b["ll"][mask]
    assert torch.equal(d["ll"], exp)
    exp = c["hist"].clone(); exp[:, mask] = b["hist"][:, mask]
    assert torch.equal(d["hist"], exp)
    # the filter carries on from the permuted state: finite, and the untouched columns of handle 2 are what they were
    e1.run(4)
    torch.cuda.synchronize()
    assert torch.isfinite(e1.raw(6, (e1.B,))).all()
    assert torch.equal(snap(e2)["x"], b["x"])
    with pytest.raises(ValueError):
        e1.resample_columns(torch.tensor([0, 1, 2, 3, 4, 9]).cuda())
    # columns as records (cross-rank theta-resampling of a sharded batch): export of one handle, import into another with indices ==
    # the in-place permutation; records of SEVERAL handles concatenated (the all-gather of the ranks' shards) address by global index
    f3, e3 = make(1)
    f4, e4 = make(1)
    torch.cuda.synchronize()
    rec = e3.export_columns()
    assert rec.shape[0] == B and rec.dtype == torch.int32
    e3.resample_columns(idx.cuda(), entire_history=True)
    e4.import_columns(rec, idx.cuda())
    torch.cuda.synchronize()
    s3, s4 = snap(e3), snap(e4)
    for k in s3:
        assert torch.equal(s3[k], s4[k]), k
    both = torch.cat([rec, e2.export_columns()], 0).contiguous()          # "world size 2": records 0..5 of handle 3, 6..11 of handle 2
    gidx = torch.tensor([7, 0, 11, 6, 2, 2])
    e4.import_columns(both, gidx.cuda())
    torch.cuda.synchronize()
    s4 = snap(e4)
    a3 = a  # handle 3 before its permutation == handle 1 before its permutation (same seed)
    for k in ("x", "lw", "pi"):
        exp = torch.stack([(b[k][:, g - 6] if g >= 6 else a3[k][:, g]) for g in gidx.tolist()], 1)
        assert torch.equal(s4[k], exp), k
    e4.run(3)
    torch.cuda.synchronize()
    assert torch.isfinite(e4.raw(6, (e4.B,))).all()
    with pytest.raises(ValueError):
        e4.import_columns(both, torch.tensor([0, 1, 2, 3, 4, 12]).cuda())
