"""Golden vectors of the theta-level arithmetic of SMC2 from the UNMODIFIED reference (run in the build container):
``python -m oracle.make_golden_smc2`` -> tests/golden/smc2_theta.npz.  Inputs are seeded clouds of unconstrained parameters with
log-weights; outputs are what the reference's own functions return: ``construct_mvn`` (inference/utils.py:61-76) mean and scale_tril,
``MultivariateNormal.log_prob`` of given points, ``PriorMixin.eval_prior(constrained=False)`` (inference/prior.py:81-90) for the two
priors of BASELINE configs[4], ``get_ess`` / ``normalize`` on the log-weights, ``systematic`` theta-indices for a given offset."""
import os

import numpy as np
import torch

from oracle.ref_loader import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    load_reference()
    from pyfilter.inference.utils import construct_mvn
    from pyfilter.inference.prior import PriorMixin  # noqa: F401  (patches torch distributions with the prior methods)
    from pyfilter.resampling import systematic
    from pyfilter.utils import get_ess, normalize
    from pyro.distributions import LogNormal, Normal

    torch.manual_seed(77)
    cases = {}
    for name, B, p, spread in (("b64_p2", 64, 2, 1.0), ("b1024_p2", 1024, 2, 3.0), ("b256_p3", 256, 3, 0.5)):
        x = torch.randn(B, p) * torch.tensor([1.0, 0.5, 2.0][:p]) + torch.tensor([0.3, -0.2, 1.0][:p])
        lw = torch.randn(B) * spread
        W = normalize(lw.clone())
        mvn = construct_mvn(x, W, scale=1.1)
        pts = torch.randn(B, p)
        u = torch.rand(1, 1)
        idx = systematic(W.clone().unsqueeze(-1), normalized=True, u=u)[:, 0]   # (the wrapper honours `u=` for 2-D input only)
        cases[f"{name}_x"], cases[f"{name}_lw"], cases[f"{name}_W"] = x.numpy(), lw.numpy(), W.numpy()
        cases[f"{name}_mean"], cases[f"{name}_tril"] = mvn.loc.numpy(), mvn.scale_tril.numpy()
        cases[f"{name}_pts"], cases[f"{name}_lp"] = pts.numpy(), mvn.log_prob(pts).numpy()
        cases[f"{name}_ess"] = get_ess(lw.clone()).numpy()
        cases[f"{name}_u"], cases[f"{name}_idx"] = u.numpy(), idx.numpy()
    # NESS: the jittering kernels' fit (location, scale) on a weighted cloud and theta-indices (jittering.py:140-225)
    from pyfilter.inference.sequential.kernels.jittering import ConstantKernel, LiuWestShrinkage, NonShrinkingKernel, ShrinkingKernel

    x = cases["b1024_p2_x"]; x = torch.from_numpy(x)
    W = torch.from_numpy(cases["b1024_p2_W"]); idx = torch.from_numpy(cases["b1024_p2_idx"])
    for kname, k in (("shrinking", ShrinkingKernel()), ("nonshrinking", NonShrinkingKernel()), ("liuwest", LiuWestShrinkage(0.98)),
                     ("constant", ConstantKernel(0.1))):
        loc, sc = k.fit(x.clone(), W.clone(), idx)
        cases[f"jit_{kname}_loc"], cases[f"jit_{kname}_scale"] = loc.numpy(), torch.as_tensor(sc).float().numpy()
    uvals = torch.randn(500) * 2.0
    pn, pl = Normal(0.0, 1.0), LogNormal(0.0, 0.5)
    cases["prior_u"] = uvals.numpy()
    cases["prior_normal"] = pn.eval_prior(pn.get_constrained(uvals), constrained=False).numpy()
    cases["prior_lognormal"] = pl.eval_prior(pl.get_constrained(uvals), constrained=False).numpy()
    np.savez_compressed(os.path.join(OUT, "smc2_theta.npz"), **cases)
    print("wrote smc2_theta.npz with", len(cases), "arrays")


if __name__ == "__main__":
    main()
