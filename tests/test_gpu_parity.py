"""GPU parity tests: the CUDA path (through the C ABI of libsmcb200.so) against the oracle and the golden vectors.

Tolerances: ancestors bit-exact for identical normalised weights and offsets; normalised weights / ESS
1e-6 relative; teacher-forced x_t 1e-6, log-weights 1e-5 (+2e-6 |lw|) for Bootstrap and 3e-5 for LinearGaussianObservations,
log-likelihood increment / mean / variance 1e-5 relative (2e-5 absolute floor)."""
import numpy as np
import pytest
import torch

from oracle import smc_oracle as O
from tests.golden_util import filter_cases, load_filter_case, load_resampling, model_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pf():
    import pyfilter_b200

    return pyfilter_b200


@pytest.fixture(params=["column", "pipeline", "twokernel"])
def smc_path(request, monkeypatch):
    """Filters of at most 4096 particles per column run in the resident column kernel (csrc/column.cuh) by default;
    ``SMCB_NO_COLUMN`` forces the per-move pipeline: the single move kernel (csrc/move.cuh, "pipeline") or, with ``SMCB_NO_MOVE``
    as well, the older resampling kernel(s) + step kernel ("twokernel").  All three are tested."""
    monkeypatch.delenv("SMCB_NO_COLUMN", raising=False)
    monkeypatch.delenv("SMCB_NO_MOVE", raising=False)
    if request.param != "column":
        monkeypatch.setenv("SMCB_NO_COLUMN", "1")
    if request.param == "twokernel":
        monkeypatch.setenv("SMCB_NO_MOVE", "1")
    return request.param


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a)) if not isinstance(a, torch.Tensor) else a
    return t.to("cuda", dtype) if dtype else t.cuda()


# ------------------------------------------------------------------------------------------------------------------ operators
def test_normalize_and_ess_golden(pf):
    g = load_resampling()
    lw = dev(g["norm_in"])
    W = pf.utils.normalize(lw.clone()).cpu().numpy()
    ref = g["norm_out"]
    ok = ~np.isnan(ref)
    assert np.allclose(W[ok], ref[ok], rtol=2e-6, atol=1e-37)
    # (all-NaN columns give NaN in the reference, i.e. they are undefined input - Appendix A-1 - and are not compared)
    ess = pf.utils.get_ess(lw.clone()).cpu().numpy()
    okc = ~np.isnan(g["norm_ess"])
    assert np.allclose(ess[okc], g["norm_ess"][okc], rtol=5e-6)


def test_systematic_golden_bit_exact(pf):
    g = load_resampling()
    names = sorted({k[4:-2] for k in g if k.startswith("sys_") and k.endswith("_W")})
    for name in names:
        W, u, idx = g[f"sys_{name}_W"], g[f"sys_{name}_u"], g[f"sys_{name}_idx"]
        got = pf.resampling.systematic(dev(W), normalized=True, u=dev(u))
        assert got.dtype == torch.int64 and got.shape == idx.shape
        assert got.stride() == (1, W.shape[0])  # transposed view like the reference's moveaxis wrapper
        assert np.array_equal(got.cpu().numpy(), idx), name


def test_reference_kat_construction(pf):
    """Same construction as the reference's tests/test_resampling.py:31-47 (seed 123, (10,300) weights, per-column offsets)."""
    torch.random.manual_seed(123)
    weights = O.normalize(torch.randn((10, 300), dtype=torch.float64)).float()  # (10 columns, 300 particles) as rows
    u = torch.rand(10, 1)
    W = weights.t().contiguous()  # (N=300, B=10)
    expect = O.systematic(W.clone(), normalized=True, u=u)
    got = pf.resampling.systematic(W.cuda(), normalized=True, u=u.cuda())
    assert torch.equal(got.cpu(), expect)
    for b in range(10):
        assert np.array_equal(got[:, b].cpu().numpy(), O.systematic_loop(W[:, b].numpy(), float(u[b])))


@pytest.mark.parametrize("n,b,std", [(1000, 3, 0.3), (4097, 2, 6.0), (1 << 20, 1, 2.0), (1_000_003, 3, 3.0),
                                     (4_000_000, 1, 0.3), (4_000_000, 1, 6.0), ((1 << 24) - 1, 1, 2.0)])
def test_systematic_random_bit_exact(pf, n, b, std):
    gen = torch.Generator().manual_seed(n + b)
    W = O.normalize(torch.randn(n, b, generator=gen) * std)
    u = torch.rand(b, 1, generator=gen)
    expect = O.systematic(W.clone(), normalized=True, u=u)
    got = pf.resampling.systematic(W.cuda(), normalized=True, u=u.cuda()).cpu()
    nbad = int((got != expect).sum())
    assert nbad == 0, f"{nbad} ancestors differ out of {n * b}"
    assert bool((got[1:] >= got[:-1]).all())  # systematic ancestors are sorted


def test_systematic_unnormalised_and_mutation(pf):
    gen = torch.Generator().manual_seed(5)
    lw = torch.randn(5000, 4, generator=gen) * 3
    lw[7, 1] = float("nan")
    lw[8, 2] = float("inf")
    u = torch.rand(4, 1, generator=gen)
    d = lw.cuda()
    got = pf.resampling.systematic(d, u=u.cuda()).cpu()
    # ancestors are exact for the weights the device normalised: re-derive them on the CPU from those weights
    Wd = pf.utils.normalize(lw.cuda()).cpu()
    assert torch.equal(got, O.systematic(Wd.clone(), normalized=True, u=u))
    ref_mut = lw.clone()
    Wr = O.normalize(ref_mut)
    assert torch.allclose(Wd, Wr, rtol=2e-6, atol=1e-37)
    assert torch.equal(d.cpu().nan_to_num(9.0), ref_mut.nan_to_num(9.0))  # the input is sanitised in place like the reference
    flips = int((got != O.systematic(Wr.clone(), normalized=True, u=u)).sum())
    assert flips <= 20  # ulp-level weight differences may move a handful of ancestors


def test_systematic_degenerate(pf):
    n = 100_000
    W = torch.zeros(n, 5)
    W[12345, 0] = 1.0
    W[:, 1] = 1.0 / n
    W[0, 2], W[n - 1, 2] = 0.5, 0.5
    W[n // 2:, 3] = 2.0 / n
    W[:3, 4] = torch.tensor([0.25, 0.5, 0.25])
    u = torch.tensor([[0.3], [0.0], [0.999], [0.5], [0.75]])
    expect = O.systematic(W.clone(), normalized=True, u=u)
    got = pf.resampling.systematic(W.cuda(), normalized=True, u=u.cuda()).cpu()
    assert torch.equal(got, expect)


def test_systematic_1d_and_random_offsets(pf):
    torch.manual_seed(3)
    W = O.normalize(torch.randn(70_000))
    got = pf.resampling.systematic(W.cuda(), normalized=True)
    assert got.shape == (70_000,) and got.dtype == torch.int64
    cnt = torch.bincount(got.cpu(), minlength=70_000).double()
    assert cnt.sum() == 70_000
    assert (cnt - 70_000 * W.double()).abs().max() <= 1.0 + 1e-3  # systematic: offspring = floor or ceil of N*W


def test_multinomial_golden_and_random(pf):
    g = load_resampling()
    for name in sorted({k[4:-2] for k in g if k.startswith("mul_") and k.endswith("_W")}):
        W, U, idx = g[f"mul_{name}_W"], g[f"mul_{name}_U"], g[f"mul_{name}_idx"]
        got = pf.resampling.multinomial(dev(W), normalized=True, U=dev(np.ascontiguousarray(U.T)))
        assert np.array_equal(got.cpu().numpy(), idx), name
    gen = torch.Generator().manual_seed(9)
    n = 300_000
    W = O.normalize(torch.randn(n, 2, generator=gen) * 2)
    U = torch.rand(2, n, dtype=torch.float64, generator=gen)
    got = pf.resampling.multinomial(W.cuda(), normalized=True, U=U.cuda()).cpu()
    for b in range(2):
        assert np.array_equal(got[:, b].numpy(), O.multinomial_restated(W[:, b].numpy(), U[b].numpy())), b


# ------------------------------------------------------------------------------------------------------- teacher-forced steps
def _make_filter(pf, g, N, B, model=None, **kw):
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, GPF, SISR, proposals

    params = {k: (torch.tensor(v) if isinstance(v, (list, tuple)) else v) for k, v in g["params"].items()}
    m = ts.build(g["model"], **params)
    cls = {"sisr": SISR, "apf": APF, "gpf": GPF}[g["alg"]]
    if g["alg"] == "gpf":                        # its default GaussianProposal
        prop = None
    elif g["proposal"].startswith("nested"):     # "nested:<num_samples>"
        prop = proposals.NestedProposal(int(g["proposal"].split(":")[1]))
    elif g["proposal"].startswith("linearized"):   # "linearized:<n_steps>:<alpha>:<second order>"
        parts = g["proposal"].split(":")
        prop = proposals.Linearized(n_steps=int(parts[1]), alpha=float(parts[2]), use_second_order=bool(int(parts[3])))
    else:
        prop = {"bootstrap": proposals.Bootstrap, "linear_gaussian": proposals.LinearGaussianObservations}[g["proposal"]]()
    res = {"systematic": pf.resampling.systematic, "multinomial": pf.resampling.multinomial}[g["resampler"]]
    f = cls(m, N, proposal=prop, resampling=res, seed=1234, **kw)
    f.set_batch_shape(torch.Size([B]) if B else torch.Size([]))
    return f


def _noise_buffers(e, z, u, U):
    """Reference-layout noise (N,[B],[d]) -> device layout (D,B,ld) / (B,) / (B,ld)."""
    D, B, ld, N = e.D, e.B, e.ld, e.N
    zt = torch.as_tensor(z).float().reshape(N, B, D)
    eps = torch.zeros(D, B, ld)
    eps[:, :, :N] = zt.permute(2, 1, 0)
    ub = torch.as_tensor(u).float().reshape(B).clone()
    Ub = None
    if U is not None and np.asarray(U).size:
        Ub = torch.zeros(B, ld, dtype=torch.float64)
        Ub[:, :N] = torch.as_tensor(U).reshape(N, B).t()
    return eps.cuda(), ub.cuda(), (Ub.cuda() if Ub is not None else None)


@pytest.mark.parametrize("exact_weights", [False, True])
@pytest.mark.parametrize("tag", filter_cases())
def test_teacher_forced_steps_vs_reference_golden(pf, tag, exact_weights, smc_path):
    """exact_weights=False (default): the handle rounds its resampling weights to multiples of 2^-52 (|dW| <= 1.1e-16) and every
    column takes the chain-free path; True: unrounded weights, columns with tiny weights take the transducer scan."""
    g = load_filter_case(tag)
    N, B, T = g["N"], g["B"], g["T"]
    f = _make_filter(pf, g, N, B, exact_weights=exact_weights)
    e = f._get_engine(2)
    lgo = g["proposal"] != "bootstrap"   # the proposals with a kernel density in the weight: three log-densities, 3e-5
    wdump = torch.zeros(e.B, e.ld, device="cuda")
    e.dump_noise(None, None, wdump)
    model = O.build_model(g["model"], model_params(g))
    total_flips = 0
    for t in range(T):
        x_prev, lw_prev = torch.from_numpy(g["x_prev"][t]), torch.from_numpy(g["lw_prev"][t])
        e.load_state(x_prev, lw_prev, torch.from_numpy(g["inds_prev"][t]), t)
        nested = g["proposal"].startswith("nested")
        if nested:   # the inner samples' normals and the Exp(1) values of torch.multinomial's single-sample draw
            e.set_nested_noise(torch.from_numpy(g["z"][t]), torch.from_numpy(g["Un"][t]))
            eps, u, U = _noise_buffers(e, np.zeros(g["x"][t].shape, np.float32), g["u"][t], g["U"][t])
        else:
            eps, u, U = _noise_buffers(e, g["z"][t], g["u"][t], g["U"][t])
        gpf = g["alg"] == "gpf"
        if gpf:      # the draws of the sample from the Gaussian approximation travel through the same hook (one "inner sample")
            e.set_nested_noise(torch.from_numpy(g["z2"][t])[None], None)
        e.set_noise(eps, u, U)
        y = torch.as_tensor(g["y"][t]).float().reshape(1, -1).cuda()
        e.set_observations(y, t)
        e.run(1)
        torch.cuda.synchronize()
        st = e.make_state()
        x, lw = st.timeseries_state.value.cpu(), st.weights.cpu()
        inds = st.previous_indices.cpu()
        gx, glw, ginds = torch.from_numpy(g["x"][t]), torch.from_numpy(g["lw"][t]), torch.from_numpy(g["prev_inds"][t])
        drew = bool(g["drew"][t].any())
        same = torch.ones_like(ginds, dtype=torch.bool)
        if drew and g["resampler"] == "systematic":
            # (1) operator-level exactness: ancestors == CPU systematic on the weights the device actually normalised
            Wd = wdump[:, :N].t().cpu()
            cols = torch.from_numpy(g["drew"][t]).reshape(-1)
            exp = O.systematic(Wd.clone(), normalized=True, u=torch.from_numpy(g["u"][t]).reshape(-1, 1))
            got2 = inds.reshape(N, -1)
            assert torch.equal(got2[:, cols], exp[:, cols]), (tag, t, "ancestors vs CPU systematic on device weights")
        same = inds == ginds
        flips = int((~same).sum())
        total_flips += flips
        assert flips <= max(2, N * max(B, 1) // 200), (tag, t, flips)  # ulp-level weight differences only
        sx = same if x.dim() == same.dim() else same.unsqueeze(-1).expand_as(x)
        xtol = (1e-5 if gpf else 2e-6) * max(1.0, float(gx.abs().max()))   # GPF: the cloud's mean / Cholesky factor come out of reductions
        picks = 0
        if nested:
            # the pick is argmax(softmax(lp) / E) in float32: where two quotients tie within an ulp of the soft-max the device may
            # take the other inner sample (same rule as for ancestors: a handful per step at most)
            bad = ((x - gx).abs() > xtol) & sx
            picks = int((bad if bad.dim() == same.dim() else bad.any(-1)).sum())
            assert picks <= max(2, N * max(B, 1) // 200), (tag, t, picks)
            sx = sx & ~bad
        assert torch.allclose(x[sx], gx[sx], rtol=0, atol=xtol), (tag, t, (x - gx)[sx].abs().max())
        tol = 3e-4 if gpf else (3e-5 if lgo else 1e-5)   # GPF: an ulp of the fitted mean moves every log-density
        fin = torch.isfinite(glw) & same
        assert ((lw[fin] - glw[fin]).abs() <= tol + 4e-6 * glw[fin].abs()).all(), (tag, t, (lw - glw)[fin].abs().max())
        if flips == 0 and picks == 0:
            for key, got in (("ll", st.get_loglikelihood()), ("mean", st.get_mean()), ("var", st.get_variance())):
                a, b = got.cpu().numpy().reshape(-1), g[key][t].reshape(-1)
                assert np.allclose(a, b, rtol=2e-5, atol=2e-5), (tag, t, key, a, b)
        # (3) EVERY quantity of the move, on every step: the oracle's step run on the device's ancestors (an ulp of a weight may move
        # an ancestor; with the ancestors forced the rest of the arithmetic is compared in full, flipped particles included)
        kw = dict(force_idx=inds.reshape(gx.shape[:ginds.dim()]))
        yt = torch.as_tensor(g["y"][t]).float()
        ut = torch.from_numpy(g["u"][t]).reshape(-1)
        zt = torch.from_numpy(g["z"][t])
        if nested:
            zt = (zt, torch.from_numpy(g["Un"][t]))
        if gpf:
            ref = O.gpf_step(model, g["proposal"], x_prev, lw_prev, torch.from_numpy(g["inds_prev"][t]), yt, (zt, torch.from_numpy(g["z2"][t])))
        elif g["alg"] == "sisr":
            ref = O.sisr_step(model, g["proposal"], x_prev, lw_prev, torch.from_numpy(g["inds_prev"][t]), yt, zt,
                              ut, resampler=g["resampler"], **kw)
        else:
            ref = O.apf_step(model, g["proposal"], x_prev, lw_prev, torch.from_numpy(g["inds_prev"][t]), yt, zt,
                             ut, resampler=g["resampler"], **kw)
        if nested:
            badr = (x - ref["x"]).abs() > xtol
            nb_ = int((badr if badr.dim() == same.dim() else badr.any(-1)).sum())
            assert nb_ <= max(2, N * max(B, 1) // 200), (tag, t, "picks vs oracle on device ancestors", nb_)
            if nb_:
                continue   # the moments below depend on the picks
        assert torch.allclose(x, ref["x"], rtol=0, atol=xtol), (tag, t, "x vs oracle on device ancestors")
        finr = torch.isfinite(ref["lw"])
        assert ((lw[finr] - ref["lw"][finr]).abs() <= tol + 4e-6 * ref["lw"][finr].abs()).all(), (tag, t, "lw vs oracle on device ancestors")
        for key, got in (("ll", st.get_loglikelihood()), ("mean", st.get_mean()), ("var", st.get_variance())):
            a, b = got.cpu().numpy().reshape(-1), ref[key].numpy().reshape(-1)
            assert np.allclose(a, b, rtol=2e-5, atol=2e-5), (tag, t, key, "vs oracle on device ancestors", a, b)
    print(tag, "ancestor flips vs golden over all steps:", total_flips)


# ------------------------------------------------------------------------------------------------------------- free running
def test_kalman_agreement_config1(pf, smc_path):
    """Reference accuracy criterion (tests/filters/test_particle.py:105-111) for the 1-D linear-Gaussian model."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, SISR, proposals

    torch.manual_seed(123)
    m = O.build_model("lg_ar1")
    _, y = m.simulate(100)
    p = O.DEFAULT_PARAMS["lg_ar1"]
    km, _, kll = O.kalman_filter_1d(y.numpy(), p["alpha"], p["beta"], p["sigma"], p["a"], p["b"], p["s"], p["alpha"],
                                    p["sigma"] ** 2 / (1 - p["beta"] ** 2))
    for cls in (SISR, APF):
        for prop in (proposals.Bootstrap, proposals.LinearGaussianObservations):
            for bshape in (torch.Size([]), torch.Size([3])):
                f = cls(ts.build("lg_ar1"), 1500, proposal=prop(), seed=7)
                f.set_batch_shape(bshape)
                res = f.batch_filter(y, bar=False)
                assert len(res.states) == 1
                ll = res.loglikelihood.cpu().numpy()
                assert (np.abs((kll - ll) / kll) < 0.1).all(), (cls.__name__, prop.__name__, ll, kll)
                means = res.filter_means[1:].cpu().numpy()
                assert means.shape == ((100,) + tuple(bshape) + (1,))
                k = km[:, None] if len(bshape) else km
                assert np.median(np.abs((k - means[..., 0]) / k)) < 0.1


@pytest.mark.parametrize("name,alg,prop", [("sv_ar1", "apf", "bootstrap"), ("sine_em", "apf", "linear_gaussian"),
                                           ("lorenz63_em", "sisr", "bootstrap"), ("lg_ar1", "sisr", "bootstrap")])
def test_free_running_statistics_vs_oracle(pf, name, alg, prop, smc_path):
    """Free-running agreement can only be statistical: total log-likelihood and mean path within a few
    Monte-Carlo standard errors of the oracle's."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, SISR, proposals

    torch.manual_seed(11)
    mo = O.build_model(name)
    _, y = mo.simulate(60)
    N = 20_000
    lls = []
    for seed in range(3):
        torch.manual_seed(100 + seed)
        lls.append(float(O.batch_filter(mo, alg, prop, y, N)["loglikelihood"]))
    ref_means = O.batch_filter(mo, alg, prop, y, N)["filter_means"]
    cls = {"sisr": SISR, "apf": APF}[alg]
    pr = {"bootstrap": proposals.Bootstrap, "linear_gaussian": proposals.LinearGaussianObservations}[prop]()
    f = cls(ts.build(name), N, proposal=pr, seed=5)
    res = f.batch_filter(y, bar=False)
    ll = float(res.loglikelihood)
    spread = max(np.std(lls), 1e-3 * abs(np.mean(lls)), 0.05)
    assert abs(ll - np.mean(lls)) < 6 * spread, (ll, lls)
    d = (res.filter_means.cpu() - ref_means).abs()
    scale = ref_means.abs().mean() + ref_means.std()
    assert float(d.mean()) < 0.05 * float(scale), (float(d.mean()), float(scale))


def test_filter_single_steps_match_batch(pf, smc_path):
    """filter() move by move == batch_filter() for the same seed (same Philox counters), incl. a missing observation."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, SISR, proposals

    torch.manual_seed(2)
    _, y = O.build_model("sine_em").simulate(12)
    y[5] = float("nan")
    for cls in (SISR, APF):
        fa = cls(ts.build("sine_em"), 3000, seed=99)
        ra = fa.batch_filter(y, bar=False)
        fb = cls(ts.build("sine_em"), 3000, seed=99, fold_lookahead=False)
        st = fb.initialize()
        rb = fb.initialize_with_result(st)
        for yt in y:
            st = fb.filter(yt, st, result=rb)
        if smc_path in ("column", "pipeline") and cls is APF:
            # fa folds the look-ahead inside the column / move kernel, fb (fold_lookahead=False) runs the pre-weight kernel:
            # the float32 normaliser differs in its last bit, a rare ancestor flips and the runs decorrelate -
            # identical up to the first flip, Monte-Carlo agreement afterwards
            assert torch.allclose(ra.filter_means[:6], rb.filter_means[:6], rtol=1e-4, atol=1e-5)
            assert float((ra.filter_means - rb.filter_means).abs().max()) < 0.2
            continue
        assert torch.allclose(ra.filter_means, rb.filter_means, rtol=1e-4, atol=1e-5)
        assert torch.allclose(ra.loglikelihood, rb.loglikelihood, rtol=1e-4, atol=1e-4)
        xa, xb = ra.latest_state.timeseries_state.value, rb.latest_state.timeseries_state.value
        if cls is SISR:
            assert torch.equal(xa, xb)


def test_batch_filter_host_entry_point(pf, smc_path):
    """The C-ABI end-to-end call on HOST buffers (what bench.py times as `e2e`)."""
    import ctypes as C
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF

    torch.manual_seed(4)
    _, y = O.build_model("sv_ar1").simulate(30)
    f = APF(ts.build("sv_ar1"), 50_000, seed=3)
    e = f._get_engine(31)
    means = torch.zeros(31, 1, 1)
    ll = torch.zeros(1)
    lib = pf._lib.load_library()
    yh = y.float().contiguous()
    pf._lib.check(lib.smcb_filter_batch_filter_host(e.handle, yh.data_ptr(), 30, means.data_ptr(), None, None, ll.data_ptr(), None))
    torch.manual_seed(8)
    ref = O.batch_filter(O.build_model("sv_ar1"), "apf", "bootstrap", y, 50_000)
    assert abs(float(ll) - float(ref["loglikelihood"])) < 0.5
    assert float((means[:, 0] - ref["filter_means"]).abs().mean()) < 0.05
    assert e.info().slow_tiles >= 0


# ----------------------------------------------------------------------------------------- BASELINE.json full sizes: properties
@pytest.mark.parametrize("exact_weights", [False, True])
def test_full_size_config3_properties(pf, exact_weights):
    """configs[2] at its full 4,000,000 particles: free-running moves, then one move checked through size-independent properties -
    ancestors sorted, in range, offspring counts add up to N, every particle with offspring count consistent with the systematic
    bound on |count - N W|, ancestors bit-exact against the oracle's CPU systematic for the dumped device weights and offset."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF

    N = 4_000_000
    torch.manual_seed(6)
    _, y = O.build_model("sv_ar1").simulate(16)
    f = APF(ts.build("sv_ar1"), N, seed=21, exact_weights=exact_weights)
    e = f._get_engine(20)
    e.initialize()
    e.set_observations(y.float().reshape(-1, 1).cuda().contiguous(), 0)
    e.run(12)
    wdump = torch.zeros(e.B, e.ld, device="cuda")
    udump = torch.zeros(e.B, device="cuda")
    e.dump_noise(None, udump, wdump)
    e.run(1)
    torch.cuda.synchronize()
    anc = e.prev_inds().cpu()
    W = wdump[0, :N].cpu()
    u = udump.cpu().reshape(1, 1)
    assert int(anc.min()) >= 0 and int(anc.max()) < N
    assert bool((anc[1:] >= anc[:-1]).all())
    counts = torch.bincount(anc, minlength=N)
    assert int(counts.sum()) == N
    # systematic bound |count - N W| < 1, loosened by the reference's own float32 probe grid (spacing 0.25 beyond i = 2^21)
    assert float((counts.double() - N * W.double()).abs().max()) < 2.0
    assert abs(float(W.double().sum()) - 1.0) < 1e-5
    if not exact_weights:  # rounded to multiples of 2^-52
        scaled = W.double() * 2.0**52
        assert bool((scaled == torch.round(scaled)).all())
    expect = O.systematic(W.clone().unsqueeze(1), normalized=True, u=u)[:, 0]
    assert torch.equal(anc, expect)
    assert e.info().slow_tiles == 0
    st = e.make_state()
    assert torch.isfinite(st.get_loglikelihood()).all() and torch.isfinite(st.get_mean()).all()


def test_full_size_config5_shard_vs_oracle(pf, smc_path):
    """configs[4], one GPU's shard: 4096 state particles x 128 theta-particles (sine diffusion, per-column gamma/sigma).  One
    teacher-forced move against the oracle on every column: ancestors exact for the device weights, log-likelihood increments and means
    within the stated tolerance."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF

    N, B = 4096, 128
    gen = torch.Generator().manual_seed(9)
    gamma, sigma = torch.randn(B, generator=gen) * 0.3, torch.exp(0.25 * torch.randn(B, generator=gen))
    mo = O.build_model("sine_em", dict(gamma=gamma, sigma=sigma))
    x0 = torch.randn(N, B, generator=gen)
    lw0 = torch.randn(N, B, generator=gen) * 0.7
    z = torch.randn(N, B, generator=gen)
    u = torch.rand(B, generator=gen)
    y = torch.tensor(0.4)
    f = APF(ts.build("sine_em", gamma=gamma, sigma=sigma), N, seed=2)
    f.set_batch_shape(torch.Size([B]))
    e = f._get_engine(2)
    e.load_state(x0, lw0, torch.arange(N).unsqueeze(1).expand(N, B), 0)
    eps = torch.zeros(e.D, e.B, e.ld)
    eps[0, :, :N] = z.t()
    e.set_noise(eps.cuda(), u.cuda(), None)
    wdump = torch.zeros(e.B, e.ld, device="cuda")
    e.dump_noise(None, None, wdump)
    e.set_observations(y.reshape(1, 1).cuda(), 0)
    e.run(1)
    torch.cuda.synchronize()
    st = e.make_state()
    ref = O.apf_step(mo, "bootstrap", x0, lw0, torch.arange(N).unsqueeze(1).expand(N, B), y, z, u=u)
    anc = st.previous_indices.cpu()
    Wd = wdump[:, :N].t().cpu()
    assert torch.equal(anc, O.systematic(Wd.clone(), normalized=True, u=u.reshape(B, 1)))
    # log-weights here reach -100 (s = 0.1): half an ulp of the float32 log-weight is 4e-6, i.e. 4e-6 relative on the weight
    assert torch.allclose(Wd, ref["resample_W"], rtol=5e-5, atol=2.3e-16)
    same = anc == ref["prev_inds"]
    assert float(same.float().mean()) > 0.98
    assert torch.allclose(st.timeseries_state.value.cpu()[same], ref["x"][same], atol=3e-6)
    lw = st.weights.cpu()
    fin = torch.isfinite(ref["lw"]) & same
    assert bool(((lw - ref["lw"])[fin].abs() <= 1e-5 + 4e-6 * ref["lw"][fin].abs()).all())
    # the weights are extremely peaked here (s = 0.1, prior N(0,1)): one flipped ancestor (an ulp of a weight) moves the statistics,
    # so they are compared on the columns whose ancestors all agree (same rule as the golden-vector test)
    clean = same.all(dim=0)
    assert int(clean.sum()) >= 4  # ~2 ulp-induced flips per column on average: roughly e^-2 of the columns are clean
    assert torch.allclose(st.get_loglikelihood().cpu()[clean], ref["ll"][clean], rtol=2e-5, atol=2e-4)
    assert torch.allclose(st.get_mean().cpu().reshape(-1)[clean], ref["mean"].reshape(-1)[clean], rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("N", [5000, 3000])
def test_fused_resampling_batched_exact(pf, smc_path, N):
    """Default path (one fused resampling kernel, Philox offsets, weights rounded to multiples of 2^-52) on a ragged batch:
    131 columns x 5000 particles (two tiles per column, the second one mostly padding), SISR so that only some columns resample in a
    move.  Ancestors of every resampling column are bit-exact against the oracle's CPU systematic for the dumped weights and offsets;
    the other columns keep their ancestors."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import SISR

    B = 131   # N = 3000: one ragged tile per column, served by the resident column kernel unless smc_path == "pipeline"
    torch.manual_seed(3)
    _, y = O.build_model("lg_ar1").simulate(12)
    gen = torch.Generator().manual_seed(4)
    f = SISR(ts.build("lg_ar1", sigma=0.05 + 0.2 * torch.rand(B, generator=gen)), N, seed=17, ess_threshold=0.5)
    f.set_batch_shape(torch.Size([B]))
    e = f._get_engine(16)
    e.initialize()
    e.set_observations(y.float().reshape(-1, 1).cuda().contiguous(), 0)
    wdump = torch.zeros(e.B, e.ld, device="cuda")
    udump = torch.zeros(e.B, device="cuda")
    e.dump_noise(None, udump, wdump)
    seen_resampled = seen_kept = 0
    for t in range(10):
        before = e.prev_inds().clone()
        _, flags = e.ess()          # resample decision taken for the coming move
        torch.cuda.synchronize()
        flags = flags.cpu() > 0
        wdump.zero_()
        e.run(1)
        torch.cuda.synchronize()
        anc = e.prev_inds().cpu()
        for b in range(B):
            if flags[b]:
                W = wdump[b, :N].cpu()
                exp = O.systematic(W.clone().unsqueeze(1), normalized=True, u=udump[b].cpu().reshape(1, 1))[:, 0]
                assert torch.equal(anc[:, b], exp), (t, b)
                seen_resampled += 1
            else:
                assert torch.equal(anc[:, b], before[:, b].cpu()), (t, b)
                seen_kept += 1
    assert seen_resampled > 0 and seen_kept > 0


@pytest.mark.parametrize("n,b", [(1, 1), (2, 3), (300, 128), (4096, 5), (4097, 5)])
def test_systematic_small_and_wide(pf, n, b):
    gen = torch.Generator().manual_seed(n * 31 + b)
    W = O.normalize(torch.randn(n, b, generator=gen) * 1.5)
    u = torch.rand(b, 1, generator=gen)
    got = pf.resampling.systematic(W.cuda(), normalized=True, u=u.cuda()).cpu()
    assert torch.equal(got, O.systematic(W.clone(), normalized=True, u=u))
    lw = torch.randn(n, b, generator=gen) * 3
    got = pf.resampling.systematic(lw.clone().cuda(), u=u.cuda()).cpu()
    Wd = pf.utils.normalize(lw.clone().cuda()).cpu()
    assert torch.equal(got, O.systematic(Wd.clone(), normalized=True, u=u))


@pytest.mark.parametrize("alg", ["sisr", "apf"])
def test_observe_every_step(pf, alg, smc_path):
    """``observe_every_step = 3`` (filters/base.py:204-210): two propagate-only moves before every observation but the first.
    The result has one moment row per OBSERVATION, the time index counts every move, and the run agrees statistically with the
    oracle's restatement (itself bit-for-bit equal to the reference: tests/test_oracle_pinned.py)."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, SISR

    torch.manual_seed(21)
    mo = O.build_model("sine_em")
    _, y = mo.simulate(30)
    N, k = 20_000, 3
    cls = {"sisr": SISR, "apf": APF}[alg]
    f = cls(ts.build("sine_em", observe_every_step=k), N, seed=5)
    res = f.batch_filter(y, bar=False)
    assert res.filter_means.shape[0] == 31
    assert int(res.latest_state.timeseries_state.time_index) == 1 + 29 * k
    lls = []
    for seed in range(3):
        torch.manual_seed(200 + seed)
        lls.append(float(O.batch_filter(mo, alg, "bootstrap", y, N, observe_every_step=k)["loglikelihood"]))
    ref = O.batch_filter(mo, alg, "bootstrap", y, N, observe_every_step=k)
    spread = max(np.std(lls), 1e-3 * abs(np.mean(lls)), 0.05)
    assert abs(float(res.loglikelihood) - np.mean(lls)) < 6 * spread, (float(res.loglikelihood), lls)
    d = (res.filter_means.cpu() - ref["filter_means"]).abs()
    scale = ref["filter_means"].abs().mean() + ref["filter_means"].std()
    assert float(d.mean()) < 0.05 * float(scale)
    # move by move through filter(): same counters, same result
    f2 = cls(ts.build("sine_em", observe_every_step=k), N, seed=5, fold_lookahead=False)
    st = f2.initialize()
    r2 = f2.initialize_with_result(st)
    for yt in y:
        st = f2.filter(yt, st, result=r2)
    if alg == "apf" and smc_path != "twokernel":
        # res folds the look-ahead inside the move kernel, f2 (fold_lookahead=False) runs the pre-weight kernel: the float32 normaliser
        # of the resampling weights is reduced in a different order, an ulp flips a rare ancestor and the runs decorrelate
        #  - identical up to the first flip, Monte-Carlo agreement afterwards
        assert torch.allclose(res.filter_means[:4], r2.filter_means[:4], rtol=1e-4, atol=1e-5)
        assert float((res.filter_means - r2.filter_means).abs().max()) < 0.1
        assert abs(float(res.loglikelihood) - float(r2.loglikelihood)) < 6 * spread
        return
    assert torch.allclose(res.filter_means, r2.filter_means, rtol=1e-4, atol=1e-5)
    assert torch.allclose(res.loglikelihood, r2.loglikelihood, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("n", [(1 << 23), (1 << 23) + 5000])
def test_filter_at_the_lean_count_limit(pf, n):
    """2^23 particles is the last size served by the fused kernel / lean probe count (exact_scan.h: K <= 2^23); above it the filter
    takes the three-kernel pipeline with the general count.  One free-running APF move each, ancestors bit-exact against the oracle's
    CPU systematic for the dumped weights and offset."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF

    torch.manual_seed(8)
    _, y = O.build_model("sv_ar1").simulate(4)
    f = APF(ts.build("sv_ar1"), n, seed=31)
    e = f._get_engine(6)
    e.initialize()
    e.set_observations(y.float().reshape(-1, 1).cuda().contiguous(), 0)
    e.run(2)
    wdump = torch.zeros(e.B, e.ld, device="cuda")
    udump = torch.zeros(e.B, device="cuda")
    e.dump_noise(None, udump, wdump)
    e.run(1)
    torch.cuda.synchronize()
    anc = e.prev_inds().cpu()
    W = wdump[0, :n].cpu()
    expect = O.systematic(W.clone().unsqueeze(1), normalized=True, u=udump.cpu().reshape(1, 1))[:, 0]
    assert torch.equal(anc, expect)
    assert e.info().slow_tiles == 0


@pytest.mark.parametrize("name,alg,prop,N,B", [("sv_ar1", "apf", "bootstrap", 4096, 7), ("lorenz63_em", "sisr", "bootstrap", 1000, 3),
                                               ("sine_em", "apf", "linear_gaussian", 4000, 5), ("lg_ar1", "sisr", "bootstrap", 37, 2)])
def test_column_kernel_matches_pipeline(pf, name, alg, prop, N, B, monkeypatch):
    """The resident column kernel (csrc/column.cuh: one block owns a column for a whole run of moves) against the multi-kernel
    pipeline on the same seed: one move - identical particles, log-weights and ancestors wherever the ancestors agree (the two paths
    reduce the normalisers in a different order, so an ulp of a weight may flip a rare ancestor); eight moves in ONE launch -
    the per-move means, variances and likelihood increments agree to Monte-Carlo accuracy and the first moves to rounding."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, SISR, proposals

    cls = APF if alg == "apf" else SISR
    P = proposals.LinearGaussianObservations if prop == "linear_gaussian" else proposals.Bootstrap
    g = torch.Generator().manual_seed(11)
    _, y = ts.build(name).sample_states(9, generator=g)
    yd = y.float().reshape(9, -1).cuda().contiguous()
    out = {}
    for path in ("pipeline", "column"):
        if path == "pipeline":
            monkeypatch.setenv("SMCB_NO_COLUMN", "1")
        else:
            monkeypatch.delenv("SMCB_NO_COLUMN", raising=False)
        f = cls(ts.build(name), N, proposal=P(), seed=5)
        f.set_batch_shape(torch.Size([B]))
        e = f._get_engine(12)
        e.initialize()
        e.set_observations(yd, 0)
        launches0 = e.info().kernel_launches
        e.run(1)
        torch.cuda.synchronize()
        one = (e.x_view().cpu().clone(), e.logw_view().cpu().clone(), e.prev_inds().cpu().clone())
        e.run(7)
        torch.cuda.synchronize()
        out[path] = one + tuple(h.cpu() for h in e.history(9)) + (e.info().kernel_launches - launches0,)
    a, b = out["pipeline"], out["column"]
    assert b[6] <= 4 < a[6]            # two launches (+ the APF's first pre-weight pass) against two per move
    same = a[2] == b[2]
    assert float(same.float().mean()) > 0.999
    xs = same if a[0].dim() == 2 else same.unsqueeze(-1).expand_as(a[0])
    assert torch.equal(a[0][xs], b[0][xs]) and torch.equal(a[1][same], b[1][same])
    scale = float(a[3].abs().mean() + a[3].std())
    assert torch.allclose(a[3][:3], b[3][:3], atol=2e-3 * scale) and torch.allclose(a[5][:3], b[5][:3], atol=2e-3, rtol=1e-3)
    assert float((a[3] - b[3]).abs().max()) < max(0.1, 3.0 / N ** 0.5) * scale
    assert float((a[4] - b[4]).abs().max()) < 0.2 * float(a[4].abs().max())


# ------------------------------------------------------------------------------- the kernels the benchmark times, pinned
def _engine(pf, name, alg, N, B=0, seed=77, rows=24, **kw):
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, SISR

    f = {"apf": APF, "sisr": SISR}[alg](ts.build(name), N, seed=seed, **kw)
    if B:
        f.set_batch_shape(torch.Size([B]))
    return f._get_engine(rows)


@pytest.mark.parametrize("name,alg,N,B", [("sv_ar1", "apf", 4_000_000, 0), ("lorenz63_em", "sisr", 300_000, 3), ("sine_em", "apf", 70_001, 2)])
def test_move_kernel_bit_identical_to_two_kernel_pipeline(pf, name, alg, N, B, monkeypatch):
    """The benchmark's kernel (move_kernel, csrc/move.cuh: resampling + propagation in one pass, Philox offsets, its compile-time
    steady-state loop) against the two-kernel pipeline whose generic loop the golden vectors pin: from the SAME state, one move gives
    bit-identical particles, log-weights, folded resampling weights and ancestors - with the steady-state loop (no dump) and with the
    generic loop (noise dumped) alike; then the move is replayed on the CPU oracle with the dumped noise, offset and the device's
    ancestors, so log-likelihood increment, mean and variance of the benchmark path itself are compared with the reference arithmetic."""
    from pyfilter_b200 import _lib

    monkeypatch.delenv("SMCB_NO_COLUMN", raising=False)
    monkeypatch.delenv("SMCB_NO_MOVE", raising=False)
    torch.manual_seed(5)
    mo = O.build_model(name)
    _, y = mo.simulate(12)
    yd = y.float().reshape(12, -1).cuda().contiguous()
    e0 = _engine(pf, name, alg, N, B)
    e0.initialize()
    e0.set_observations(yd, 0)
    e0.run(8)                      # a realistic cloud: weights of a running filter, not of the prior
    torch.cuda.synchronize()
    x0, lw0, pi0 = e0.x_view().clone(), e0.logw_view().clone(), e0.prev_inds().clone()
    out = {}
    for path in ("move", "move_dump", "twokernel", "twokernel_dump"):
        if path.startswith("twokernel"):
            monkeypatch.setenv("SMCB_NO_MOVE", "1")
        else:
            monkeypatch.delenv("SMCB_NO_MOVE", raising=False)
        e = _engine(pf, name, alg, N, B)
        e.load_state(x0, lw0, pi0, 8)
        e.set_observations(yd, 0)
        eps = torch.zeros(e.D, e.B, e.ld, device="cuda")
        ud = torch.zeros(e.B, device="cuda")
        wd = torch.zeros(e.B, e.ld, device="cuda")
        if path.endswith("_dump"):
            e.dump_noise(eps, ud, wd)
        launches = e.info().kernel_launches
        e.run(1)
        torch.cuda.synchronize()
        st = e.make_state()
        rw = e.raw(_lib.PTR_RESAMPLE_LOGW, (e.B, e.ld))[:, : e.N].clone()
        out[path] = dict(x=e.x_view().clone(), lw=e.logw_view().clone(), pi=e.prev_inds().clone(), rw=rw, ll=st.get_loglikelihood().clone(),
                         mean=st.get_mean().clone(), var=st.get_variance().clone(), eps=eps, u=ud, w=wd,
                         launches=e.info().kernel_launches - launches)
    a = out["move"]
    for other in ("move_dump", "twokernel", "twokernel_dump"):
        b = out[other]
        assert torch.equal(a["pi"], b["pi"]), (other, "ancestors", int((a["pi"] != b["pi"]).sum()))
        assert torch.equal(a["x"], b["x"]), (other, "x")
        assert torch.equal(a["lw"], b["lw"]), (other, "lw")
        if alg == "apf":
            assert torch.equal(a["rw"], b["rw"]), (other, "folded resampling log-weights")
        assert torch.allclose(a["ll"], b["ll"], rtol=1e-5, atol=1e-5) and torch.allclose(a["mean"], b["mean"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(a["var"], b["var"], rtol=1e-4, atol=1e-7)
    assert torch.equal(out["move_dump"]["eps"], out["twokernel_dump"]["eps"]) and torch.equal(out["move_dump"]["u"], out["twokernel_dump"]["u"])
    assert torch.equal(out["move_dump"]["w"], out["twokernel_dump"]["w"])
    # replay on the CPU oracle with the dumped noise and the device's ancestors
    d = out["move_dump"]
    z = d["eps"][:, :, : e.N].permute(2, 1, 0).cpu()           # (N, B, D)
    z = z[..., 0] if e.D == 1 and len(mo_state_shape(mo)) == 0 else z
    if not B:
        z = z[:, 0]
    step = O.apf_step if alg == "apf" else O.sisr_step
    u = d["u"].cpu()
    ref = step(mo, "bootstrap", x0.cpu(), lw0.cpu(), pi0.cpu(), y[8].float(), z, u, force_idx=a["pi"].cpu())
    xs = float(ref["x"].abs().max())
    assert torch.allclose(a["x"].cpu(), ref["x"], rtol=0, atol=2e-6 * max(1.0, xs))
    fin = torch.isfinite(ref["lw"])
    assert bool(((a["lw"].cpu() - ref["lw"])[fin].abs() <= 1e-5 + 4e-6 * ref["lw"][fin].abs()).all())
    assert torch.allclose(a["ll"].cpu(), ref["ll"].float(), rtol=2e-5, atol=2e-4 if name == "sine_em" else 2e-5)  # (peaked weights: c5 tolerance)
    assert torch.allclose(a["mean"].cpu().reshape(-1), ref["mean"].reshape(-1), rtol=2e-5, atol=2e-5)
    assert torch.allclose(a["var"].cpu().reshape(-1), ref["var"].reshape(-1), rtol=1e-4, atol=2e-5)
    # and the ancestors are the CPU systematic ancestors for the device's weights and offset (bit-exact)
    Wd = d["w"][:, : e.N].t().cpu()
    cols = Wd.sum(0) > 0.5                      # the columns that resampled dumped their weights
    assert bool(cols.all()) or alg == "sisr"
    if bool(cols.any()):
        exp = O.systematic(Wd[:, cols].clone(), normalized=True, u=u[cols].reshape(-1, 1))
        assert torch.equal(a["pi"].cpu().reshape(e.N, -1)[:, cols], exp)


def mo_state_shape(mo):
    return (mo.state_dim,) if mo.state_dim else ()


def test_move_kernel_reproducible_and_many_windows(pf, monkeypatch):
    """(1) Same seed, same bits: tile tickets are drawn dynamically, the statistics are reduced per tile - two runs of 30 moves
    agree bit for bit.  (2) Degenerate weights: one particle owns (almost) every slot, i.e. one tile has hundreds of windows of
    offspring - still exact against the CPU systematic."""
    monkeypatch.delenv("SMCB_NO_MOVE", raising=False)
    torch.manual_seed(2)
    mo = O.build_model("sv_ar1")
    _, y = mo.simulate(32)
    yd = y.float().reshape(-1, 1).cuda().contiguous()
    res = []
    for rep in range(2):
        e = _engine(pf, "sv_ar1", "apf", 1_000_000, rows=40)
        e.initialize()
        e.set_observations(yd, 0)
        e.run(30)
        torch.cuda.synchronize()
        res.append((e.x_view().clone(), e.logw_view().clone(), e.prev_inds().clone(), e.history(31)))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    for h0, h1 in zip(res[0][3], res[1][3]):
        assert torch.equal(h0, h1)
    # degenerate cloud
    N = 300_000
    e = _engine(pf, "sv_ar1", "sisr", N, seed=3)
    x0 = torch.randn(N) * 0.1 - 1.0
    lw0 = torch.full((N,), -60.0)
    lw0[123_457] = 0.0
    lw0[5] = -3.0
    e.load_state(x0, lw0, torch.arange(N), 0)
    e.set_observations(yd, 0)
    wd = torch.zeros(e.B, e.ld, device="cuda")
    ud = torch.zeros(e.B, device="cuda")
    e.dump_noise(None, ud, wd)
    e.run(1)
    torch.cuda.synchronize()
    anc = e.prev_inds().cpu()
    exp = O.systematic(wd[0, :N].cpu().clone().unsqueeze(1), normalized=True, u=ud.cpu().reshape(1, 1))[:, 0]
    assert torch.equal(anc, exp)
    assert int((anc == 123_457).sum()) > N * 0.9
    assert torch.equal(e.x_view().cpu() != 0, torch.ones(N, dtype=torch.bool))


# ------------------------------------------------------------------------------- BASELINE configs at full size, teacher-forced
def test_full_size_config2_lgo_vs_oracle(pf):
    """configs[1]: sine diffusion, APF + LinearGaussianObservations, systematic, 1,000,000 particles - one teacher-forced move."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, proposals

    N = 1_000_000
    gen = torch.Generator().manual_seed(12)
    mo = O.build_model("sine_em")
    x0 = torch.randn(N, generator=gen) * 1.5
    lw0 = torch.randn(N, generator=gen) * 0.8
    z = torch.randn(N, generator=gen)
    u = torch.rand(1, generator=gen)
    y = torch.tensor(0.7)
    f = APF(ts.build("sine_em"), N, proposal=proposals.LinearGaussianObservations(), seed=4)
    e = f._get_engine(2)
    e.load_state(x0, lw0, torch.arange(N), 0)
    eps = torch.zeros(e.D, e.B, e.ld)
    eps[0, 0, :N] = z
    e.set_noise(eps.cuda(), u.cuda(), None)
    wdump = torch.zeros(e.B, e.ld, device="cuda")
    e.dump_noise(None, None, wdump)
    e.set_observations(y.reshape(1, 1).cuda(), 0)
    e.run(1)
    torch.cuda.synchronize()
    st = e.make_state()
    anc = st.previous_indices.cpu()
    Wd = wdump[0, :N].cpu()
    assert torch.equal(anc, O.systematic(Wd.clone().unsqueeze(1), normalized=True, u=u.reshape(1, 1))[:, 0])
    ref0 = O.apf_step(mo, "linear_gaussian", x0, lw0, torch.arange(N), y, z, u=u)
    assert torch.allclose(Wd, ref0["resample_W"], rtol=2e-5, atol=2.3e-16)
    # s = 0.1 makes the weights peaked: a relative 1e-6 on a heavy weight moves the cumulative sum by a whole probe spacing (1e-6), so
    # against the oracle's OWN (ulp-different) weights most ancestors sit one or two places away - close, not equal
    assert float((anc - ref0["prev_inds"]).abs().float().mean()) < 50.0
    ref = O.apf_step(mo, "linear_gaussian", x0, lw0, torch.arange(N), y, z, u=u, force_idx=anc)
    assert torch.allclose(st.timeseries_state.value.cpu(), ref["x"], rtol=0, atol=2e-6 * max(1.0, float(ref["x"].abs().max())))
    fin = torch.isfinite(ref["lw"])
    assert bool(((st.weights.cpu() - ref["lw"])[fin].abs() <= 3e-5 + 4e-6 * ref["lw"][fin].abs()).all())
    assert abs(float(st.get_loglikelihood()) - float(ref["ll"])) <= 2e-5 + 2e-5 * abs(float(ref["ll"]))
    assert torch.allclose(st.get_mean().cpu().reshape(-1), ref["mean"].reshape(-1), rtol=2e-5, atol=2e-5)
    assert torch.allclose(st.get_variance().cpu().reshape(-1), ref["var"].reshape(-1), rtol=1e-4, atol=2e-5)


def test_full_size_config4_multinomial_vs_oracle(pf):
    """configs[3]: Lorenz-63, SISR + Bootstrap, multinomial resampling, 2,000,000 particles - one teacher-forced move with injected
    float64 uniforms: ancestors bit-exact against the restated torch.multinomial for the device's weights."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import SISR

    N = 2_000_000
    gen = torch.Generator().manual_seed(14)
    mo = O.build_model("lorenz63_em")
    loc, scale = mo.initial_loc_scale()
    x0 = loc + scale * torch.randn(N, 3, generator=gen)
    y = torch.tensor([0.8 * -5.9 + 0.3, 0.8 * 24.5 - 0.2])
    lw0 = mo.obs_log_prob(y, x0).float() * 0.25          # uneven enough for the ESS test to fire
    z = torch.randn(N, 3, generator=gen)
    U = torch.rand(N, dtype=torch.float64, generator=gen)
    f = SISR(ts.build("lorenz63_em"), N, resampling=pf.resampling.multinomial, seed=4)
    e = f._get_engine(2)
    e.load_state(x0, lw0, torch.arange(N), 0)
    eps = torch.zeros(e.D, e.B, e.ld)
    eps[:, 0, :N] = z.t()
    Ub = torch.zeros(e.B, e.ld, dtype=torch.float64)
    Ub[0, :N] = U
    e.set_noise(eps.cuda(), None, Ub.cuda())
    wdump = torch.zeros(e.B, e.ld, device="cuda")
    e.dump_noise(None, None, wdump)
    _, flags = e.ess()
    torch.cuda.synchronize()
    assert bool(flags.cpu() > 0), "the test state must trigger resampling"
    e.set_observations(y.reshape(1, 2).cuda(), 0)
    e.run(1)
    torch.cuda.synchronize()
    st = e.make_state()
    anc = st.previous_indices.cpu()
    Wd = wdump[0, :N].cpu()
    exp = torch.from_numpy(O.multinomial_restated(Wd.numpy(), U.numpy()))
    assert torch.equal(anc, exp)
    ref = O.sisr_step(mo, "bootstrap", x0, lw0, torch.arange(N), y, z, None, resampler="multinomial", U=U.numpy(), force_idx=anc)
    assert bool(ref["resampled"])
    assert torch.allclose(st.timeseries_state.value.cpu(), ref["x"], rtol=0, atol=2e-6 * float(ref["x"].abs().max()))
    fin = torch.isfinite(ref["lw"])
    assert bool(((st.weights.cpu() - ref["lw"])[fin].abs() <= 1e-5 + 4e-6 * ref["lw"][fin].abs()).all())
    assert abs(float(st.get_loglikelihood()) - float(ref["ll"])) <= 2e-5 + 2e-5 * abs(float(ref["ll"]))
    # At 2 M particles the reference's OWN float32 reductions are 2.3e-5 (relative) away from the float64 value of the same sums
    # (measured with the oracle on this cloud: the soft-max normaliser), so the comparison with it is made at 1e-4 and the device's
    # moments are also held against the float64 moments of the device's own particles and log-weights
    assert torch.allclose(st.get_mean().cpu().reshape(-1), ref["mean"].reshape(-1), rtol=1e-4, atol=2e-4)
    assert torch.allclose(st.get_variance().cpu().reshape(-1), ref["var"].reshape(-1), rtol=3e-4, atol=2e-4)
    Wd64 = torch.softmax(st.weights.cpu().double(), 0).unsqueeze(-1)
    xd64 = st.timeseries_state.value.cpu().double()
    m64 = (Wd64 * xd64).sum(0)
    v64 = (Wd64 * (xd64 - m64) ** 2).sum(0)
    assert torch.allclose(st.get_mean().cpu().reshape(-1).double(), m64, rtol=3e-5, atol=1e-5)
    assert torch.allclose(st.get_variance().cpu().reshape(-1).double(), v64, rtol=1e-4, atol=1e-5)
    assert e.info().slow_tiles >= 0


@pytest.mark.parametrize("name,B", [("lg_ar1", 0), ("sv_ar1", 3), ("lorenz63_em", 2)])
def test_initialize(pf, name, B):
    """``ParticleFilter.initialize`` (filters/particle/base.py:87-103): x_0 = loc + scale * z for the dumped unit normals (bit for
    bit), z ~ N(0, 1), log w = 0, log-likelihood 0, previous indices = arange, history row 0 = mean / variance of x_0."""
    N = 200_000
    e = _engine(pf, name, "sisr", N, B, seed=9)
    eps = torch.zeros(e.D, e.B, e.ld, device="cuda")
    e.dump_noise(eps, None, None)
    e.initialize()
    torch.cuda.synchronize()
    st = e.make_state()
    x = st.timeseries_state.value.cpu()
    mo = O.build_model(name)
    loc, scale = mo.initial_loc_scale()
    z = eps[:, :, :N].permute(2, 1, 0).cpu()                      # (N, B, D)
    zz = z.reshape(-1).double()
    assert abs(float(zz.mean())) < 5.0 / zz.numel() ** 0.5 and abs(float(zz.std()) - 1.0) < 5.0 / zz.numel() ** 0.5
    assert abs(float((zz ** 3).mean())) < 0.05 and abs(float((zz ** 4).mean()) - 3.0) < 0.1
    xe = (loc.float() + scale.float() * z).reshape(N, e.B, e.D)
    got = x.reshape(N, e.B, e.D)
    assert torch.allclose(got, xe, rtol=1e-6, atol=1e-6)   # (the device derives the stationary scale in double precision)
    assert int(st.timeseries_state.time_index) == 0
    assert torch.equal(st.weights.cpu(), torch.zeros(st.weights.shape))
    assert torch.equal(st.get_loglikelihood().cpu(), torch.zeros(st.get_loglikelihood().shape))
    pi = st.previous_indices.cpu()
    ar = torch.arange(N)
    assert torch.equal(pi, ar if not B else ar.unsqueeze(1).expand(N, B))
    m = got.double().mean(0)
    v = got.double().var(0, unbiased=False)
    assert torch.allclose(st.get_mean().cpu().reshape(e.B, e.D).double(), m, rtol=2e-5, atol=2e-5)
    assert torch.allclose(st.get_variance().cpu().reshape(e.B, e.D).double(), v, rtol=1e-4, atol=1e-6)
    hm, hv, hl = e.history(1)
    assert torch.allclose(hm.reshape(-1).cpu(), st.get_mean().reshape(-1).cpu()) and float(hl.abs().max()) == 0.0


def test_get_ess_normalized(pf):
    """``get_ess(W, normalized=True)`` (utils.py:8-20) against the oracle."""
    gen = torch.Generator().manual_seed(3)
    lw = torch.randn(50_000, 4, generator=gen) * 2
    W = O.normalize(lw.clone())
    ref = O.get_ess(W, normalized=True)
    got = pf.utils.get_ess(W.cuda(), normalized=True).cpu()
    assert torch.allclose(got, ref, rtol=5e-6)
    got1 = pf.utils.get_ess(W[:, 0].contiguous().cuda(), normalized=True).cpu()
    assert torch.allclose(got1, ref[0], rtol=5e-6)


# ------------------------------------------------------------------------------------------------ recorded states (result objects)
@pytest.mark.parametrize("alg", ["sisr", "apf"])
def test_record_states_hold_their_own_move(pf, alg, smc_path):
    """``record_states=True`` (filters/result.py:39,119-133): every recorded state shows the particles, log-weights and ancestors of ITS
    move, not of a later one (the device buffers are rewritten by every move: the result copies a state when it is appended).  The
    recorded states are compared with a second filter of the same seed that is stepped by hand and copied after every move, and a
    stale view of the device buffers refuses to be read."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, SISR

    torch.manual_seed(4)
    _, y = O.build_model("sine_em").simulate(9)
    y[4] = float("nan")
    cls = {"sisr": SISR, "apf": APF}[alg]
    N = 6000
    fa = cls(ts.build("sine_em"), N, seed=31, record_states=True)
    ra = fa.batch_filter(y, bar=False)
    assert len(ra.states) == 10
    fb = cls(ts.build("sine_em"), N, seed=31)
    eb = fb._get_engine(12)
    eb.initialize()
    eb.set_observations(y.float().reshape(-1, 1).cuda().contiguous(), 0)
    live = eb.make_state()
    copies = [live.detach_copy()]
    for _ in range(9):
        eb.run(1)
        copies.append(eb.make_state().detach_copy())
    for i, (sa, sb) in enumerate(zip(ra.states, copies)):
        assert int(sa.timeseries_state.time_index) == i
        assert torch.equal(sa.timeseries_state.value, sb.timeseries_state.value), i
        assert torch.equal(sa.weights, sb.weights), i
        assert torch.equal(sa.previous_indices, sb.previous_indices), i
        assert torch.equal(sa["_prev_inds"], sb["_prev_inds"])
    # consecutive recorded states differ (they are not all views of the last one)
    assert not torch.equal(ra.states[3].timeseries_state.value, ra.states[9].timeseries_state.value)
    with pytest.raises(RuntimeError):
        live.weights   # the engine has moved nine times since this view was made


def test_record_intermediary_moments(pf):
    """``record_intermediary_states=True`` with ``observe_every_step = 3`` (filters/base.py:207-208): the result also carries the moments
    of the propagate-only moves - one row per move - whether or not the states themselves are recorded."""
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF

    torch.manual_seed(6)
    _, y = O.build_model("sine_em").simulate(7)
    k = 3
    fa = APF(ts.build("sine_em", observe_every_step=k), 5000, seed=3, record_intermediary_states=True)
    ra = fa.batch_filter(y, bar=False)
    fb = APF(ts.build("sine_em", observe_every_step=k), 5000, seed=3, record_intermediary_states=True, record_states=True)
    rb = fb.batch_filter(y, bar=False)
    moves = 1 + 6 * k
    assert ra.filter_means.shape[0] == moves + 1 == rb.filter_means.shape[0]
    assert len(rb.states) == moves + 1
    assert torch.allclose(ra.loglikelihood, rb.loglikelihood, rtol=1e-4, atol=1e-3)
    fc = APF(ts.build("sine_em", observe_every_step=k), 5000, seed=3)
    assert fc.batch_filter(y, bar=False).filter_means.shape[0] == 8


# ------------------------------------------------------------------------------------------------ theta shards: peer-memory exchange
@pytest.mark.parametrize("N,online", [(4096, True), (4096, False), (20_000, True)])
def test_peer_exchange_emulated_ranks(pf, N, online):
    """The exchange of the theta-sharded loop (include/smcb200.h: smcb_filter_attach_exchange): the finalising kernel stores (value, tag)
    pairs into every rank's buffer.  Two "ranks" are emulated on one GPU - two shards of one batch of 10 columns, two handles, two
    buffers - and every rank must read the log-likelihood increments and totals of ALL columns exactly as the shards computed them,
    move after move (N = 4096: the resident column kernel, online and whole runs; N = 20000: move kernel + finalize kernel).  The
    shards pass their global column offset, so the run also equals ONE unsharded filter on the same seed, bit for bit."""
    from pyfilter_b200 import _lib, timeseries as ts
    from pyfilter_b200.filters.particle import APF
    from pyfilter_b200.sharding import PeerExchange, column_shard

    THETA, world = 10, 2
    gen = torch.Generator().manual_seed(5)
    gamma, sigma = torch.randn(THETA, generator=gen) * 0.3, torch.exp(0.3 * torch.randn(THETA, generator=gen))
    torch.manual_seed(3)
    _, y = O.build_model("sine_em").simulate(13)
    yd = y.float().reshape(-1, 1).cuda().contiguous()
    bufs = [torch.zeros(2 * 2 * THETA, dtype=torch.int64, device="cuda") for _ in range(world)]
    engines, xs = [], []
    for r in range(world):
        lo, hi = column_shard(THETA, r, world)
        f = APF(ts.build("sine_em", gamma=gamma[lo:hi], sigma=sigma[lo:hi]), N, seed=77, column_offset=lo)
        f.set_batch_shape(torch.Size([hi - lo]))
        e = f._get_engine(16)
        e.initialize()
        e.set_observations(yd, 0)
        engines.append((e, lo, hi))
        xs.append(PeerExchange(e, THETA, lo, buffers=bufs, rank=r))
    whole = APF(ts.build("sine_em", gamma=gamma, sigma=sigma), N, seed=77)
    whole.set_batch_shape(torch.Size([THETA]))
    ew = whole._get_engine(16)
    ew.initialize()
    ew.set_observations(yd, 0)
    for steps in ([1] * 6 if online else [4, 5, 3]):
        for e, _, _ in engines:
            e.run(steps)
        ew.run(steps)
        got = [x.wait() for x in xs]
        torch.cuda.synchronize()
        inc = torch.cat([e.raw(_lib.PTR_LL, (e.B,)).clone() for e, _, _ in engines])
        tot = torch.cat([e.raw(_lib.PTR_LL_TOTAL, (e.B,)).clone() for e, _, _ in engines])
        for a, b in got:
            assert torch.equal(a, inc) and torch.equal(b, tot)
        assert torch.equal(tot, ew.raw(_lib.PTR_LL_TOTAL, (THETA,))), "sharded run differs from the unsharded one"
    assert torch.isfinite(tot).all()
