"""ShardedSMC2 on N ranks (torchrun) against SMC2 in one process, same seed: the theta log-weights, the parameter cloud, the ESS history and
the running log-likelihoods must be IDENTICAL bit for bit (the filters' random streams are keyed by the global column index, the
theta-level state is replicated, column migration copies records).  Prints one JSON line on rank 0 with the verdict and the wall times.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/smc2_sharded_check.py [theta] [state particles] [T]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pyfilter_b200 import _lib, timeseries as ts
from pyfilter_b200.filters.particle import proposals
from pyfilter_b200.inference import SMC2, ShardedSMC2, LogNormal, Normal
from oracle import smc_oracle as O   # (data simulation only)

THETA = int(sys.argv[1]) if len(sys.argv) > 1 else 256
NSTATE = int(sys.argv[2]) if len(sys.argv) > 2 else 512
T = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(2)
_, y = O.build_model("sine_em", dict(gamma=0.5, sigma=2.0)).simulate(T)
builder = lambda p: ts.build("sine_em", gamma=p["gamma"], sigma=p["sigma"])
priors = lambda: {"gamma": Normal(0.0, 1.0), "sigma": LogNormal(0.0, 0.5)}
kw = dict(particles=THETA, state_particles=NSTATE, proposal=proposals.LinearGaussianObservations(), threshold=0.5, seed=5, max_observations=T + 2)


def run(cls):
    alg = cls(builder, priors(), **kw)
    state = alg.initialize()
    warm = 4
    for yt in y[:warm]:                 # warm-up: first launches and ONE forced rejuvenation (first-use allocations of both filters,
        state = alg.step(yt, state)     # the migration buffers, cuSOLVER); the same sequence in both runs, so they stay comparable
    state = alg.rejuvenate(state)
    torch.cuda.synchronize()
    if dist.is_initialized() and cls is ShardedSMC2:
        dist.barrier()
    t0 = time.perf_counter()
    for yt in y[warm:]:
        state = alg.step(yt, state)
    torch.cuda.synchronize()
    return alg, state, (time.perf_counter() - t0) * (T - 1) / (T - warm)


# one-time library initialisation outside the timed runs: cuSOLVER / cuBLAS handles (the p x p Cholesky factor of the proposal), NCCL
_w = torch.eye(2, device="cuda")
torch.linalg.cholesky_ex(_w); torch.linalg.solve_triangular(_w, _w, upper=False)
_g = torch.empty(world * 4, device="cuda"); dist.all_gather_into_tensor(_g, torch.ones(4, device="cuda")); torch.cuda.synchronize()
alg, state, dt = run(ShardedSMC2)
dist.barrier()
res = None
if rank == 0:
    alg1, state1, dt1 = run(SMC2)
    lo, hi = alg._lo, alg._hi
    same = {"w": bool(torch.equal(state.w, state1.w)), "theta": bool(torch.equal(alg.context.values, alg1.context.values)),
            "ess": state.ess == state1.ess, "rejuvenations": state.rejuvenations == state1.rejuvenations,
            "loglikelihood": bool(torch.equal(state.loglikelihood, state1.loglikelihood[lo:hi])),
            "particles": bool(torch.equal(state.engine.x_view(), state1.engine.x_view()[:, lo:hi])),
            "state_particles": state.engine.N == state1.engine.N}
    if alg.phases.on:
        print("phases sharded:", {k: round(v * 1e3, 2) for k, v in alg.phases.t.items()}, "single:", {k: round(v * 1e3, 2) for k, v in alg1.phases.t.items()}, file=sys.stderr)
    res = {"world": world, "theta": THETA, "state_particles": NSTATE, "observations": T, "identical": same, "all_identical": all(same.values()),
           "rejuvenations": state.rejuvenations, "acceptance": [round(a, 3) for a in state.acceptance], "state_particles_final": state.engine.N,
           "sharded_ms_per_observation": dt / (T - 1) * 1e3, "single_process_ms_per_observation": dt1 / (T - 1) * 1e3,
           "posterior_mean": {k: float(v) for k, v in alg.posterior_mean(state).items()}}
    print(json.dumps(res), flush=True)
dist.barrier()
dist.destroy_process_group()
if rank == 0 and not res["all_identical"]:
    sys.exit(1)
