"""The reference's README example (README.md:44-79: sine diffusion observed with noise, APF with the LinearGaussianObservations proposal)
on pyfilter_b200: same classes, same calls, the model comes from the compiled zoo.  Run on a B200:  python examples/readme_sine_apf.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pyfilter_b200 import timeseries as ts
from pyfilter_b200.filters.particle import APF, proposals

torch.manual_seed(123)
model = ts.build("sine_em", gamma=0.0, sigma=1.0, dt=0.1, a=1.0, b=0.0, s=0.1)   # dx = sin(x - gamma) dt + sigma dW ;  y = x + 0.1 nu
x, y = model.sample_states(250)                                                  # synthetic data (torch CPU)

filt = APF(model, 1_000_000, proposal=proposals.LinearGaussianObservations())    # BASELINE.json configs[1]
result = filt.batch_filter(y)                                                    # the whole time loop on the device

means = result.filter_means[1:, 0].cpu()
print("log-likelihood", float(result.loglikelihood))
print("RMSE of the filter mean against the simulated state", float((means - x).pow(2).mean().sqrt()))
path = result.latest_state.predict_path(model, 10)                               # particle/state.py:173-174
print("10-step predictive paths", tuple(path.get_paths()[0].shape))
