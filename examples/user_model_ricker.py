"""A model that is not in the compiled zoo: the user writes mean_scale / the observation density as CUDA device functions
(tests/user_models/ricker_user.h documents the contract) and the library is compiled once more with them - about 40 s the first time,
cached afterwards.  python examples/user_model_ricker.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from pyfilter_b200 import timeseries as ts
from pyfilter_b200.filters.particle import APF

make = ts.compile_user_model(open(os.path.join(ROOT, "tests", "user_models", "ricker_user.h")).read(), state_dim=1, obs_dim=1)
model = make(0.8, 20.0, 0.15, 0.3)                       # r, K, sigma, tau
torch.manual_seed(1)
x, ys = torch.log(torch.tensor(20.0)), []
for _ in range(100):                                     # simulate on the CPU with the same equations
    x = x + 0.8 * (1.0 - x.exp() / 20.0) + 0.15 * torch.randn(())
    ys.append(x.exp() + 0.3 * (0.5 * x).exp() * torch.randn(()))
result = APF(model, 500_000).batch_filter(torch.stack(ys))
print("log-likelihood", float(result.loglikelihood), "last filter mean of log-population", float(result.filter_means[-1]))
