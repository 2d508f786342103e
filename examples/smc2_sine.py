"""SMC2 (Chopin et al.) for the parameters of the sine diffusion - the analogue of the reference's examples/*.ipynb inference cells -
with the theta-particles as the columns of ONE resident batch of filters.  One GPU:  python examples/smc2_sine.py
Several GPUs (theta-particles sharded, bit-identical result):  torchrun --nproc-per-node 8 examples/smc2_sine.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pyfilter_b200 import timeseries as ts
from pyfilter_b200.filters.particle import proposals
from pyfilter_b200.inference import SMC2, LogNormal, Normal, ShardedSMC2

world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    import torch.distributed as dist

    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))

torch.manual_seed(2)
_, y = ts.build("sine_em", gamma=0.5, sigma=2.0).sample_states(200)

cls = ShardedSMC2 if world > 1 else SMC2
alg = cls(lambda p: ts.build("sine_em", gamma=p["gamma"], sigma=p["sigma"]),        # the model builder the reference calls per theta
          {"gamma": Normal(0.0, 1.0), "sigma": LogNormal(0.0, 0.5)},                # priors
          particles=1024, state_particles=4096,                                     # BASELINE.json configs[4]
          proposal=proposals.LinearGaussianObservations(), threshold=0.2, seed=1, max_observations=256)
state = alg.fit(y)
if int(os.environ.get("RANK", "0")) == 0:
    print("posterior means", {k: round(float(v), 3) for k, v in alg.posterior_mean(state).items()}, "(data: gamma 0.5, sigma 2.0)")
    print("rejuvenations", state.rejuvenations, "acceptance", [round(a, 2) for a in state.acceptance], "final ESS", round(state.ess[-1], 1))
if world > 1:
    dist.destroy_process_group()
