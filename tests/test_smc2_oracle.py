"""CPU: the oracle of SMC2's theta-level arithmetic (oracle/smc2_oracle.py) against the golden vectors the UNMODIFIED reference produced
(tests/golden/smc2_theta.npz, oracle/make_golden_smc2.py) and - when /root/reference is present - against the live reference."""
import os

import numpy as np
import pytest
import torch

from oracle import smc2_oracle as S
from oracle import smc_oracle as O
from oracle.ref_loader import load_reference, reference_available

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "smc2_theta.npz")
CASES = ("b64_p2", "b1024_p2", "b256_p3")


def _g():
    z = np.load(GOLDEN)
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.mark.parametrize("name", CASES)
def test_smc2_oracle_vs_golden(name):
    g = _g()
    x, lw, W = g[f"{name}_x"], g[f"{name}_lw"], g[f"{name}_W"]
    assert torch.allclose(S.normalize(lw), W, rtol=1e-6, atol=1e-9)
    assert torch.allclose(S.get_ess(lw), g[f"{name}_ess"], rtol=1e-5)
    mvn = S.symmetric_proposal(x, W)
    assert torch.allclose(mvn.loc, g[f"{name}_mean"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(mvn.scale_tril, g[f"{name}_tril"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(mvn.log_prob(g[f"{name}_pts"]), g[f"{name}_lp"], rtol=1e-5, atol=1e-4)
    idx = O.systematic(W.clone().unsqueeze(-1), normalized=True, u=g[f"{name}_u"])[:, 0]
    assert torch.equal(idx, g[f"{name}_idx"])


def test_smc2_oracle_priors_vs_golden():
    g = _g()
    u = g["prior_u"]
    assert torch.allclose(S.normal_unconstrained_log_prob(u, 0.0, 1.0), g["prior_normal"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(S.normal_unconstrained_log_prob(u, 0.0, 0.5), g["prior_lognormal"], rtol=1e-5, atol=1e-5)


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container)")
def test_smc2_oracle_vs_live_reference():
    load_reference()
    from pyfilter.inference.utils import construct_mvn
    from pyfilter.utils import get_ess

    torch.manual_seed(5)
    x, lw = torch.randn(300, 2) * 2.0, torch.randn(300) * 2.0
    W = S.normalize(lw)
    ref = construct_mvn(x, W, scale=1.1)
    mine = S.symmetric_proposal(x, W)
    assert torch.equal(ref.loc, mine.loc) and torch.equal(ref.scale_tril, mine.scale_tril)
    assert torch.allclose(get_ess(lw.clone()), S.get_ess(lw), rtol=1e-6)
    # acceptance rule: the reference's expression on the same numbers
    ll_old, ll_new, pr_old, pr_new = torch.randn(300), torch.randn(300), torch.randn(300), torch.randn(300)
    new_x = ref.sample((300,))
    new_kernel = construct_mvn(new_x, torch.full((300,), 1.0 / 300), scale=1.1)
    u = torch.rand(300)
    acc, log_acc = S.run_pmmh_acceptance(x, new_x, mine, S.symmetric_proposal(new_x, torch.full((300,), 1.0 / 300)), pr_old, pr_new, ll_old, ll_new, u)
    ref_log = (new_kernel.log_prob(x) - ref.log_prob(new_x)) + (pr_new - pr_old) + (ll_new - ll_old)
    assert torch.allclose(log_acc, ref_log, rtol=1e-6, atol=1e-6) and torch.equal(acc, u.log() < ref_log)
    assert S.smc2_needs_rejuvenation(lw, 0.9) and not S.smc2_needs_rejuvenation(torch.zeros(50), 0.2)


@pytest.mark.parametrize("kind", ["shrinking", "nonshrinking", "liuwest", "constant"])
def test_jitter_kernels_oracle_vs_golden(kind):
    """NESS: ``fit`` of the jittering kernels (inference/sequential/kernels/jittering.py:140-225) as the unmodified reference computed it."""
    g = _g()
    x, W, idx = g["b1024_p2_x"], g["b1024_p2_W"], g["b1024_p2_idx"]
    loc, sc = S.jitter_fit(kind, x.clone(), W.clone(), idx)
    assert torch.allclose(loc, g[f"jit_{kind}_loc"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(torch.as_tensor(sc).float(), g[f"jit_{kind}_scale"], rtol=1e-5, atol=1e-8)
