"""The reference's examples/stochastic-volatility.ipynb on this library: a Verhulst volatility process observed through a sinh-arcsinh
transformed normal every fifth Euler step, days without a price change masked to NaN, fitted with SMC2 over APF(model, 400) - the
notebook's configuration.  The model is not in the compiled zoo: tests/user_models/verhulst_sas_user.h states it as device code and
the library is compiled once more with it (cached).  Synthetic returns stand in for the notebook's AAPL download.

    python examples/stochastic_volatility.py [theta particles, default 1000] [observations, default 300]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from pyfilter_b200 import timeseries as ts
from pyfilter_b200.filters.particle import APF
from pyfilter_b200.inference import SMC2, Exponential, LogNormal, Normal

thetas = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
DT = 0.2
EVERY = int(1.0 / DT)
make = ts.compile_user_model(open(os.path.join(ROOT, "tests", "user_models", "verhulst_sas_user.h")).read(), state_dim=1, obs_dim=1)

# synthetic "daily returns in percent" from the model itself (kappa, gamma, sigma, mu, nu, tau), on the filter's schedule
true = dict(kappa=0.05, gamma=1.2, sigma=0.12, mu=0.05, nu=-0.1, tau=1.1)
torch.manual_seed(7)
v, ys = torch.tensor(true["gamma"]), []
for k in range(T):
    for _ in range(1 if k == 0 else EVERY):
        v = v + true["kappa"] * v * (true["gamma"] - v) * DT + true["sigma"] * v * math.sqrt(DT) * torch.randn(())
    ys.append(true["mu"] + v * torch.sinh((torch.asinh(torch.randn(())) + true["nu"]) * true["tau"]))
y = torch.stack(ys)
y[torch.rand(T) < 0.03] = float("nan")                                   # "days where the price has not changed"


def build_model(p):                                                      # the notebook's build_model(context)
    return make(p["kappa"], p["gamma"], p["sigma"], p["mu"], p["nu"], p["tau"], DT, observe_every_step=EVERY)


priors = {"kappa": Exponential(10.0), "gamma": LogNormal(0.0, 1.0), "sigma": LogNormal(math.log(0.05), 1.0),
          "mu": Normal(0.0, 0.5), "nu": Normal(0.0, 0.15), "tau": LogNormal(0.0, 0.1)}
alg = SMC2(build_model, priors, particles=thetas, state_particles=400, filter_cls=APF, threshold=0.2, num_steps=5, seed=1,
           max_observations=T, max_increases=8)
state = alg.fit(y)
print("posterior means", {k: round(float(val), 3) for k, val in alg.posterior_mean(state).items()})
print("data generated with", true)
print("rejuvenations", state.rejuvenations, "acceptance", [round(a, 2) for a in state.acceptance], "state particles", state.engine.N,
      "final ESS", round(state.ess[-1], 1))
