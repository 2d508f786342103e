import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import smc_oracle as O
from pyfilter_b200 import _lib, timeseries as ts
from pyfilter_b200.inference import SMC2, LogNormal, Normal
from pyfilter_b200.filters.particle import proposals
from pyfilter_b200.inference import smc2 as M
def builder(p): return ts.build("sine_em", gamma=p["gamma"], sigma=p["sigma"])
pri = {"gamma": Normal(0.0, 1.0), "sigma": LogNormal(0.0, 0.5)}
torch.manual_seed(2)
truth = O.build_model("sine_em", dict(gamma=0.5, sigma=2.0))
_, y = truth.simulate(60)
orig = SMC2._run_pmmh
def traced(self, ctx, state, kernel, pe, sub, T):
    B = int(self.particles[0])
    # noise of the likelihood estimate: re-filter the CURRENT parameters with the proposal engine
    pe.set_params(self._builder(ctx.constrained())); pe.set_seed(self._draw_seed()); pe.initialize(); pe.set_observations(self._y_dev[:T], 0); pe.run(T)
    same = pe.raw(_lib.PTR_LL_TOTAL, (B,)) - state.loglikelihood
    acc = orig(self, ctx, state, kernel, pe, sub, T)
    print(f"T={T} N={pe.N} refilter-same-theta: mean {float(same.mean()):.3f} std {float(same.std()):.3f} | accepted {float(acc.float().mean()):.3f} "
          f"| unique theta {len(torch.unique(ctx.values[:,0]))} tril {kernel[1].cpu().tolist()}")
    return acc
SMC2._run_pmmh = traced
alg = SMC2(builder, pri, particles=128, state_particles=256, proposal=proposals.LinearGaussianObservations(), threshold=0.5, seed=5, max_observations=64)
state = alg.initialize()
try:
    for t, yt in enumerate(y):
        state = alg.step(yt, state)
        print(t, "ess %.1f" % state.ess[-1], "N", state.engine.N, "post", {k: round(float(v), 3) for k, v in alg.posterior_mean(state).items()})
except Exception as e:
    print("EXC", type(e).__name__, e)
