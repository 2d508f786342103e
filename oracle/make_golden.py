"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference (build container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/*.npz

Teacher-forced protocol: the reference filter is advanced one ``filter()`` call at a
time; the random numbers it consumed are recovered exactly by re-winding torch's CPU generator and repeating
the reference's draw sequence (``u`` of shape (k,1) when k columns resample - resampling.py:40-41 - then the
transition noise, Appendix A-15).  Each file stores, per step, the inputs ``x_prev, lw_prev, inds_prev, y, u|U, z``
and the reference's outputs ``x, lw, ll, mean, var, prev_inds``.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference, load_reference_resampling  # noqa: E402
from oracle.ref_models import build_reference_model  # noqa: E402
from oracle import smc_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy().copy()


def filter_case(tag, model_name, params, alg, proposal, resampler, N, B, T, seed, nan_steps=(), prefix="filter"):
    pf = load_reference()
    from pyfilter.filters.particle import APF, GPF, SISR, proposals as pr
    from pyfilter import resampling as RR
    from pyfilter.utils import get_ess, normalize

    torch.manual_seed(seed)
    sim = O.build_model(model_name, {k: (v[0] if isinstance(v, (list, tuple)) else v) for k, v in params.items()})
    _, y = sim.simulate(T)
    for s in nan_steps:
        y[s] = float("nan")
    ref_params = {k: (torch.tensor(v) if isinstance(v, (list, tuple)) else v) for k, v in params.items()}
    ssm = build_reference_model(model_name, {**O.DEFAULT_PARAMS[model_name], **ref_params})
    cls = {"sisr": SISR, "apf": APF, "gpf": GPF}[alg]
    nested_m = int(proposal.split(":")[1]) if proposal.startswith("nested") else 0
    if alg == "gpf":                        # gpf.py:24: the default GaussianProposal
        prop = None
    elif nested_m:                          # "nested:<num_samples>" (proposals/nested.py:17)
        prop = pr.NestedProposal(nested_m)
    elif proposal.startswith("linearized"):   # "linearized:<n_steps>:<alpha>:<second order>" (proposals/linearized.py:22)
        parts = proposal.split(":")
        prop = pr.Linearized(n_steps=int(parts[1]), alpha=float(parts[2]), use_second_order=bool(int(parts[3])))
    else:
        prop = {"bootstrap": pr.Bootstrap, "linear_gaussian": pr.LinearGaussianObservations}[proposal]()
    filt = cls(ssm, N, proposal=prop, resampling={"systematic": RR.systematic, "multinomial": RR.multinomial}[resampler])
    bshape = torch.Size([B]) if B else torch.Size([])
    filt.set_batch_shape(bshape)
    state = filt.initialize()
    d = ssm.hidden.event_shape
    nb = B if B else 1
    rec = {k: [] for k in ("x_prev", "lw_prev", "inds_prev", "u", "U", "z", "x", "lw", "ll", "mean", "var", "prev_inds", "drew")}
    if nested_m:
        rec["Un"] = []
    if alg == "gpf":
        rec["z2"] = []
    x0 = _np(state.timeseries_state.value)
    for t in range(T):
        x_prev, lw_prev = state.timeseries_state.value.clone(), state.weights.clone()
        inds_prev = state.previous_indices.clone()
        rng = torch.get_rng_state()
        new_state = filt.filter(y[t], state)
        new_state.timeseries_state.value  # materialise lazy samples before touching the generator
        after = torch.get_rng_state()
        # replay the draws
        torch.set_rng_state(rng)
        isnan = bool(torch.isnan(y[t]).all())
        if alg == "sisr":
            mask = (get_ess(normalize(lw_prev.clone()), True) < 0.9 * N).reshape(-1)
        elif alg == "gpf":
            mask = torch.zeros(nb, dtype=torch.bool)   # nothing resamples
        else:
            mask = torch.full((nb,), not isnan)
        u = np.zeros(nb, dtype=np.float32)
        U = np.zeros((N, nb) if resampler == "multinomial" else (0,), dtype=np.float64)
        if bool(mask.any()):
            k = int(mask.sum())
            if resampler == "systematic":
                u[mask.numpy()] = torch.empty((k, 1)).uniform_().reshape(-1).numpy()
            else:
                Uf = np.zeros((N, nb))
                Uf[:, mask.numpy()] = torch.empty((k, N), dtype=torch.float64).uniform_().numpy().T
                U = Uf
        if nested_m:   # hidden_density.sample(num_samples), then Categorical.sample: torch.multinomial's Exp(1) race (nested.py:29,40)
            z = torch.empty((nested_m,) + tuple(filt.particles) + tuple(d)).normal_()
            # (Categorical normalises the moved-axis view, which keeps its strides: the race's Exp(1) tensor is laid out sample-major)
            Un = torch.empty(nested_m, int(np.prod(tuple(filt.particles)))).exponential_(1).t().reshape(tuple(filt.particles) + (nested_m,))
            rec["Un"].append(_np(Un))
        else:
            z = torch.empty(tuple(filt.particles) + tuple(d)).normal_()
        if alg == "gpf":   # the sample from the Gaussian approximation (approximate.py:26), none on a propagate-only move
            z2 = torch.zeros(tuple(filt.particles) + tuple(d)) if isnan else torch.empty(tuple(filt.particles) + tuple(d)).normal_()
            rec["z2"].append(_np(z2))
        assert torch.equal(torch.get_rng_state(), after), f"{tag}: draw replay out of sync at step {t}"
        for k_, v in (("x_prev", x_prev), ("lw_prev", lw_prev), ("inds_prev", inds_prev), ("z", z),
                      ("x", new_state.timeseries_state.value), ("lw", new_state.weights),
                      ("ll", new_state.get_loglikelihood()), ("mean", new_state.get_mean()),
                      ("var", new_state.get_variance()), ("prev_inds", new_state.previous_indices)):
            rec[k_].append(_np(v))
        rec["u"].append(u)
        rec["U"].append(U)
        rec["drew"].append(mask.numpy().copy())
        state = new_state
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update(y=_np(y), x0=x0, N=N, B=B, T=T, seed=seed, model=model_name, alg=alg, proposal=proposal,
               resampler=resampler, params_json=np.array(repr(params)))
    np.savez_compressed(os.path.join(OUT, f"{prefix}_{tag}.npz"), **out)
    print("wrote", tag, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim > 0 and k in ("x", "lw", "U")})


def resampling_cases():
    R = load_reference_resampling()
    systematic, multinomial, normalize = R["resampling"].systematic, R["resampling"].multinomial, R["utils"].normalize
    torch.manual_seed(123)
    cases = {}
    specs = [("n300_b10", 300, 10, 1.0), ("n1000_b3_heavy", 1000, 3, 6.0), ("n4096_b2", 4096, 2, 2.0),
             ("n65536_b1", 65536, 1, 3.0), ("n1_b2", 1, 2, 1.0), ("n2_b2", 2, 2, 1.0), ("n33_b4", 33, 4, 12.0)]
    for name, n, b, std in specs:
        lw = torch.randn(n, b) * std
        W = normalize(lw.clone())
        u = torch.rand(b, 1)
        idx = systematic(W.clone(), normalized=True, u=u)
        cases[f"sys_{name}_W"] = _np(W)
        cases[f"sys_{name}_u"] = _np(u)
        cases[f"sys_{name}_idx"] = _np(idx.contiguous())
    # degenerate columns: one-hot, exact zeros, uniform
    n = 513
    W = torch.zeros(n, 4)
    W[100, 0] = 1.0
    W[:, 1] = 1.0 / n
    W[::2, 2] = 2.0 / (n + 1)
    W[0, 3], W[n - 1, 3] = 0.5, 0.5
    u = torch.tensor([[0.25], [0.999999], [0.0], [0.5]])
    cases["sys_degenerate_W"], cases["sys_degenerate_u"] = _np(W), _np(u)
    cases["sys_degenerate_idx"] = _np(systematic(W.clone(), normalized=True, u=u).contiguous())
    # normalize / ess incl. NaN / inf hygiene (utils.py:57-62)
    lw = torch.randn(257, 5) * 3
    lw[3, 1] = float("nan")
    lw[4, 1] = float("inf")
    lw[5, 2] = -float("inf")
    lw[:, 3] = -float("inf")
    lw[:, 4] = -1e30
    cases["norm_in"] = _np(lw)
    W = normalize(lw.clone())
    cases["norm_out"] = _np(W)
    cases["norm_ess"] = _np(R["utils"].get_ess(W, normalized=True))
    # multinomial with replayed float64 uniforms
    for name, n, b, std in [("n300_b3", 300, 3, 1.0), ("n5000_b2", 5000, 2, 4.0)]:
        W = normalize(torch.randn(n, b) * std)
        rng = torch.get_rng_state()
        idx = multinomial(W.clone(), normalized=True)
        after = torch.get_rng_state()
        torch.set_rng_state(rng)
        U = torch.empty((b, n), dtype=torch.float64).uniform_()
        assert torch.equal(torch.get_rng_state(), after)
        cases[f"mul_{name}_W"], cases[f"mul_{name}_U"] = _np(W), _np(U.T)
        cases[f"mul_{name}_idx"] = _np(idx.contiguous())
    np.savez_compressed(os.path.join(OUT, "resampling.npz"), **cases)
    print("wrote resampling.npz with", len(cases), "arrays")


def residual_cases():
    """``pyfilter.resampling.residual`` (resampling.py:66-105, 1-D only): weights, the float64 uniforms its multinomial part drew
    (recovered by re-winding the generator) and the indices it returned.  Own file and own seed so that resampling.npz is untouched."""
    R = load_reference_resampling()
    residual, normalize = R["resampling"].residual, R["utils"].normalize
    torch.manual_seed(321)
    cases = {}
    specs = [("n300", 300, 1.0), ("n1000_heavy", 1000, 6.0), ("n4096", 4096, 2.0), ("n33", 33, 0.3), ("n65536", 65536, 3.0)]
    for name, n, std in specs:
        W = normalize(torch.randn(n) * std)
        rng = torch.get_rng_state()
        idx = residual(W.clone(), normalized=True)
        after = torch.get_rng_state()
        m = n - int((n * W).floor().sum())
        torch.set_rng_state(rng)
        U = torch.empty((1, m), dtype=torch.float64).uniform_() if m else torch.empty((1, 0), dtype=torch.float64)
        assert torch.equal(torch.get_rng_state(), after), name
        cases[f"res_{name}_W"], cases[f"res_{name}_U"], cases[f"res_{name}_idx"] = _np(W), _np(U[0]), _np(idx)
    # nothing left for the multinomial part: uniform weights (every floor is 1) and a one-hot column (floor = N)
    n = 64
    for name, W in (("uniform", torch.full((n,), 1.0 / n)), ("onehot", torch.nn.functional.one_hot(torch.tensor(5), n).float())):
        cases[f"res_{name}_W"], cases[f"res_{name}_U"] = _np(W), np.zeros(0)
        cases[f"res_{name}_idx"] = _np(residual(W.clone(), normalized=True))
    np.savez_compressed(os.path.join(OUT, "residual.npz"), **cases)
    print("wrote residual.npz with", len(cases), "arrays")


def oracle_only_cases():
    """f2: multi-dimensional LinearGaussianObservations on Lorenz-63 (examples/lorenz.ipynb:214).  Written as ``oracleonly_*.npz``
    while the oracle existed before the CUDA path; the kernel exists now (csrc/step.cuh, ProposalLGO<.., true>) and the files are
    part of the GPU parity set (``filter_*.npz``).  A ``prefix="oracleonly"`` keeps a future row out of the GPU tests the same way."""
    filter_case("c4_apf_lgo_sys", "lorenz63_em", {}, "apf", "linear_gaussian", "systematic", 400, 0, 8, 131)
    filter_case("c4_sisr_lgo_sys", "lorenz63_em", {}, "sisr", "linear_gaussian", "systematic", 400, 0, 8, 132)


def gpf_cases():
    """f4: the Gaussian particle filter (filters/particle/gpf.py) - scalar and vector state, batched and not, a NaN observation."""
    filter_case("c1_gpf", "lg_ar1", {}, "gpf", "bootstrap", "systematic", 500, 0, 8, 161, nan_steps=(3,))
    filter_case("c3_gpf_b3", "sv_ar1", {}, "gpf", "bootstrap", "systematic", 400, 3, 6, 162)
    filter_case("c4_gpf", "lorenz63_em", {}, "gpf", "bootstrap", "systematic", 600, 0, 6, 163)
    filter_case("c4_gpf_b2", "lorenz63_em", {}, "gpf", "bootstrap", "systematic", 300, 2, 5, 164, nan_steps=(2,))


def nested_cases():
    """f2: ``NestedProposal`` (proposals/nested.py) - scalar and vector state, batched and not, SISR and APF."""
    filter_case("c3_sisr_nested", "sv_ar1", {}, "sisr", "nested:20", "systematic", 400, 0, 6, 151)
    filter_case("c1_apf_nested_b3", "lg_ar1", {}, "apf", "nested:50", "systematic", 300, 3, 6, 152)
    filter_case("c4_sisr_nested", "lorenz63_em", {}, "sisr", "nested:12", "systematic", 300, 0, 6, 153)
    filter_case("c2_apf_nested", "sine_em", {}, "apf", "nested:8", "systematic", 400, 0, 6, 154)


def linearized_cases():
    """f2: the ``Linearized`` proposal (proposals/linearized.py, proposals/utils.py:30-146) - first order (the default) and second order,
    scalar and vector state, batched and not."""
    filter_case("c3_sisr_lin1", "sv_ar1", {}, "sisr", "linearized:1:0.0001:0", "systematic", 500, 0, 6, 141)
    filter_case("c3_apf_lin2nd", "sv_ar1", {}, "apf", "linearized:3:0.0001:1", "systematic", 500, 0, 6, 142)
    filter_case("c1_sisr_lin2nd_b3", "lg_ar1", {}, "sisr", "linearized:2:0.0001:1", "systematic", 400, 3, 6, 143)
    filter_case("c2_apf_lin_alpha", "sine_em", {}, "apf", "linearized:2:0.01:0", "systematic", 400, 0, 6, 144)
    filter_case("c4_sisr_lin2nd", "lorenz63_em", {}, "sisr", "linearized:2:0.0001:1", "systematic", 300, 0, 6, 145)
    filter_case("c4_apf_lin1", "lorenz63_em", {}, "apf", "linearized:1:0.0001:0", "systematic", 300, 0, 6, 146)


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "linearized":   # added later: leaves the other files as they are
        linearized_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "gpf":
        gpf_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "nested":
        nested_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "residual":   # added later: leaves the other files as they are
        residual_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "oracle_only":
        oracle_only_cases()
        return
    resampling_cases()
    residual_cases()
    filter_case("c1_sisr_boot", "lg_ar1", {}, "sisr", "bootstrap", "systematic", 1000, 0, 12, 123)
    filter_case("c1_sisr_boot_b3_nan", "lg_ar1", {}, "sisr", "bootstrap", "systematic", 193, 3, 10, 124, nan_steps=(4,))
    filter_case("c1_sisr_lgo", "lg_ar1", {}, "sisr", "linear_gaussian", "systematic", 257, 0, 8, 125)
    filter_case("c1_apf_boot", "lg_ar1", {}, "apf", "bootstrap", "systematic", 500, 0, 8, 126, nan_steps=(3,))
    filter_case("c2_apf_lgo", "sine_em", {}, "apf", "linear_gaussian", "systematic", 1000, 0, 10, 127)
    filter_case("c3_apf_boot", "sv_ar1", {}, "apf", "bootstrap", "systematic", 1000, 0, 10, 128)
    filter_case("c4_sisr_boot_mult", "lorenz63_em", {}, "sisr", "bootstrap", "multinomial", 400, 0, 8, 129)
    filter_case("c4_sisr_boot_sys", "lorenz63_em", {}, "sisr", "bootstrap", "systematic", 400, 0, 8, 130)
    filter_case("c5_apf_boot_theta", "sine_em", {"gamma": [0.0, 0.3, -0.4, 1.0], "sigma": [1.0, 0.7, 1.3, 0.9]},
                "apf", "bootstrap", "systematic", 256, 4, 8, 131)
    filter_case("c5_sisr_boot_theta", "sine_em", {"gamma": [0.0, 0.3, -0.4, 1.0], "sigma": [1.0, 0.7, 1.3, 0.9]},
                "sisr", "bootstrap", "systematic", 256, 4, 10, 132)


if __name__ == "__main__":
    main()
