#!/usr/bin/env python
"""bench.py - particle-steps/sec of the SMC inner loop (BASELINE.json metric) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--particles P]

Headline workload (BASELINE.json configs[2], the configuration the metric "at 4M particles (1D SSM)" is quoted on): stochastic-
volatility model, APF + Bootstrap proposal, systematic resampling every move, 4,000,000 particles, one filter per GPU (replicas: the
particle dimension of one filter does not shard without an exchange step).  A "step" is one filter move (predict +
resample + propagate + weight + moments) over all particles.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes through the C-ABI call on HOST
buffers (smcb_filter_batch_filter_host: H2D of the observations, all moves, D2H of moments and likelihoods inside the timed region);
`roofline` is for the dominant kernel, from algorithmic bytes over its CUDA-event duration; `cpu_baseline` times the
oracle port of the reference's torch-CPU path on this box's host cores on a bounded sample of the same workload.  The same line also
carries (so that the driver's records hold every workload BASELINE.json names): `configs` - the other BASELINE configurations
device-timed on this rank's GPU, incl. the 3-D Lorenz-63 filter; `smc2_shard` - configs[4], 1024 theta x 4096 particles with the theta
columns sharded over the ranks (strong scaling; "batch" = all moves in one launch, "online" = one move per launch with the exchange of
the log-likelihood increments after every move); `exact_weights_true` and `quantised_vs_exact_flip_rate` - the cost and the effect of
the rounding-free weight mode; `torch_cuda_baseline` - the reference arithmetic with its tensors on cuda:0 (torch's generic ATen
kernels), the GPU path pyfilter users have today.

`--impl reference` runs the UNMODIFIED reference package (installed by __graft_entry__.build() into the git-ignored baseline/_ref with
`pip install --no-deps --target`, imported with oracle/standins for the absent stochproc / pyro) on the host cores; if that package is
not there it times the pinned oracle port and says so (`kind`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec at 4M particles (1D SSM)"
UNIT = "particle-steps/s"
WORKLOAD = "sv_ar1 APF bootstrap systematic, 4M particles (BASELINE.json configs[2])"


def simulate(name, T, seed=123):
    import torch
    from pyfilter_b200 import timeseries as ts

    g = torch.Generator().manual_seed(seed)
    _, y = ts.build(name).sample_states(T, generator=g)
    return y.float().contiguous()


def simulate_sv(T, seed=123):
    return simulate("sv_ar1", T, seed)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons WHILE the timed region runs: NVML polled from a thread every millisecond (the timed
    region is tens of milliseconds, too short for `nvidia-smi -lms`); falls back to one nvidia-smi query per sample."""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.sm, self.mask, self.max_mhz, self.stop_flag, self.thread, self.h = index, [], 0, None, False, None, None
        self.errors = 0
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _poll(self):
        while not self.stop_flag:
            if self.nv:
                try:
                    self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                except Exception:
                    self.errors += 1
                try:
                    self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    try:
                        self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    except Exception:
                        self.errors += 1
                time.sleep(0.001)
            else:
                try:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    a, b = [float(v) for v in out.strip().split(",")[:2]]
                    self.sm.append(a); self.max_mhz = b
                except Exception:
                    self.errors += 1
                    time.sleep(0.005)

    def start(self):
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(2)
        sm = sorted(self.sm)
        reasons = sorted(k for k, bit in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------- CPU arms
def cpu_baseline(particles, budget_s=20.0):
    """The oracle port (torch CPU, every host thread) on a bounded sample of the same workload: fewer moves, same particles."""
    import torch
    from oracle import smc_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = O.build_model("sv_ar1")
    y = simulate_sv(64)
    torch.manual_seed(123)
    t0 = time.perf_counter()
    O.batch_filter(model, "apf", "bootstrap", y[:2], particles)
    per = (time.perf_counter() - t0) / 2
    steps = int(max(3, min(40, budget_s / max(per, 1e-3))))
    t0 = time.perf_counter()
    O.batch_filter(model, "apf", "bootstrap", y[:steps], particles)
    dt = time.perf_counter() - t0
    return {"value": particles * steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"oracle/smc_oracle.py batch_filter (torch {torch.__version__} CPU, {cores} threads), sv_ar1 APF bootstrap systematic, "
                      f"{particles} particles x {steps} moves, {dt:.1f} s"}


def _load_installed_reference():
    """The unmodified reference package from baseline/_ref (pip --target install made by __graft_entry__.build()), imported with the
    stand-ins for its absent dependencies (oracle/standins: stochproc 0.3.0, pyro, matplotlib, statsmodels - SURVEY.md Appendix C)."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "pyfilter")):
        return None
    for p in (os.path.join(ROOT, "oracle", "standins"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import importlib

        mod = importlib.import_module("pyfilter")
        if not os.path.abspath(mod.__file__).startswith(os.path.abspath(ref)):
            return None
        return mod
    except Exception:
        return None


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, same metric / config: the
    unmodified package when baseline/_ref holds it (kind "reference"), else the pinned oracle port (kind "port")."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    K, W = args.steps, args.warmup
    y = simulate_sv(W + K)
    n = args.particles
    ref = _load_installed_reference()
    from oracle import smc_oracle as O

    if ref is not None:
        from oracle.ref_models import build_reference_model
        from pyfilter.filters.particle import APF as RefAPF   # the reference's class (baseline/_ref)

        def run(yy, particles):
            f = RefAPF(build_reference_model("sv_ar1", O.DEFAULT_PARAMS["sv_ar1"]), particles)
            return f.batch_filter(yy, bar=False)

        kind, what = "reference", f"unmodified pyfilter {getattr(ref, '__version__', '?')} APF.batch_filter (baseline/_ref, stand-ins for stochproc/pyro)"
    else:
        model = O.build_model("sv_ar1")

        def run(yy, particles):
            return O.batch_filter(model, "apf", "bootstrap", yy, particles)

        kind, what = "port", "oracle port of pyfilter APF.batch_filter"
    # bounded sample: shrink the particle count until K + W moves fit in ~2 minutes (throughput is flat in N at this size)
    torch.manual_seed(123)
    t0 = time.perf_counter()
    run(y[:1], min(n, 1_000_000))
    per_particle = (time.perf_counter() - t0) / min(n, 1_000_000)
    while n > 250_000 and per_particle * n * (K + W) > 120.0:
        n //= 2
    if W:
        run(y[:W], n)
    t0 = time.perf_counter()
    run(y[W:W + K], n)
    dt = time.perf_counter() - t0
    value = n * K / dt
    sample = f"{what} on torch {torch.__version__} CPU, {cores} threads, sv_ar1, {n} particles x {K} moves per run"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "particles": n, "particles_requested": args.particles},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------- device legs
def _timed_moves(e, moves, stream, dist=None):
    """Device time (ms) of `moves` filter moves, CUDA events on the launching stream, max over ranks."""
    import torch

    torch.cuda.synchronize()
    if dist:
        dist.barrier()
        torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    e.run(moves)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if dist:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


def bench_configs(world, dist, moves=200):
    """The other BASELINE.json configurations, device-timed on every rank's GPU (replicas; value = all ranks / max time)."""
    import torch
    import pyfilter_b200 as pf
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, SISR, proposals

    cfgs = [
        ("c1 lg_ar1 SISR bootstrap systematic 1k", "lg_ar1", SISR, proposals.Bootstrap, pf.resampling.systematic, 1_000, 0, 24),
        ("c2 sine_em APF LinearGaussianObservations systematic 1M", "sine_em", APF, proposals.LinearGaussianObservations, pf.resampling.systematic, 1_000_000, 0, 24),
        ("c4 lorenz63_em (3-D) SISR bootstrap multinomial 2M", "lorenz63_em", SISR, proposals.Bootstrap, pf.resampling.multinomial, 2_000_000, 0, 40),
        ("c4s lorenz63_em (3-D) SISR bootstrap systematic 2M", "lorenz63_em", SISR, proposals.Bootstrap, pf.resampling.systematic, 2_000_000, 0, 40),
        ("c4l lorenz63_em (3-D) APF LinearGaussianObservations systematic 2M (examples/lorenz.ipynb:214)", "lorenz63_em", APF, proposals.LinearGaussianObservations, pf.resampling.systematic, 2_000_000, 0, 40),
        ("c5 sine_em APF bootstrap systematic 4096 x 128 theta (one GPU's shard of configs[4])", "sine_em", APF, proposals.Bootstrap, pf.resampling.systematic, 4096, 128, 24),
    ]
    peak, _ = load_peaks()
    out = {}
    stream = torch.cuda.current_stream()
    for name, model, cls, prop, res, N, B, bytes_pp in cfgs:
        try:
            torch.manual_seed(123)
            kw = dict(gamma=torch.randn(B), sigma=torch.exp(0.5 * torch.randn(B))) if B else {}
            T = moves if "multinomial" not in name else max(20, moves // 4)
            y = simulate(model, T + 24)
            f = cls(ts.build(model, **kw), N, proposal=prop(), resampling=res, seed=7)
            if B:
                f.set_batch_shape(torch.Size([B]))
            e = f._get_engine(T + 30)
            e.initialize()
            e.set_observations(y.reshape(T + 24, -1).cuda().contiguous(), 0)
            e.run(20)
            ms = _timed_moves(e, T, stream, dist)
            nb = B if B else 1
            rate = world * N * nb * T / (ms * 1e-3)
            out[name.split()[0]] = {"workload": name, "value": rate, "unit": UNIT, "us_per_move": ms * 1e3 / T, "moves": T,
                                    "algorithmic_bytes_per_particle": bytes_pp,
                                    "roofline_frac": bytes_pp * N * nb / (ms * 1e-3 / T) / 1e9 / peak, "scaling": "weak (replicas)"}
            del e, f
        except Exception as ex:  # a leg that fails must not take the headline down with it
            out[name.split()[0]] = {"workload": name, "error": repr(ex)[:200]}
    return out


def bench_smc2(world, rank, dist, K=250, W=10):
    """BASELINE.json configs[4]: 1024 theta x 4096 state particles, sine diffusion, APF + Bootstrap, systematic - STRONG scaling: the
    theta columns are block distributed over the ranks (pyfilter_b200.sharding.column_shard), every rank runs the resident column
    kernel on its shard.  "batch" = K moves in one launch + one exchange, "online" = one move per launch + the exchange of the
    (B_local,) log-likelihood increments after every move (what SMC2's theta-level ESS test needs)."""
    import torch
    from pyfilter_b200 import _lib, timeseries as ts
    from pyfilter_b200.filters.particle import APF
    from pyfilter_b200.sharding import LogLikelihoodGather, column_shard

    THETA, N = 1024, 4096
    lo, hi = column_shard(THETA, rank, world)
    torch.manual_seed(123)
    gamma, sigma = torch.randn(THETA), torch.exp(0.5 * torch.randn(THETA))   # theta ~ prior
    y = simulate("sine_em", 5 * (W + K) + 16)
    y_dev = y.reshape(-1, 1).cuda().contiguous()
    f = APF(ts.build("sine_em", gamma=gamma[lo:hi], sigma=sigma[lo:hi]), N, seed=123, column_offset=lo)
    f.set_batch_shape(torch.Size([hi - lo]))
    e = f._get_engine(5 * (W + K) + 20)
    stream = torch.cuda.current_stream()
    e.initialize()
    e.set_observations(y_dev, 0)
    ll_view, ll_tot_view = e.raw(_lib.PTR_LL, (e.B,)), e.raw(_lib.PTR_LL_TOTAL, (e.B,))
    gather = LogLikelihoodGather(THETA, "cuda") if dist else None

    def timed(fn):
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream); fn(); ev1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # (1) the exchange as NCCL collective (torch.distributed): the baseline
    def online_nccl(moves):
        if not dist:
            e.run_stepwise(moves)   # (the loop in C: no exchange on one GPU)
            return
        for _ in range(moves):
            e.run(1)
            gather(ll_view)

    def batch_nccl(moves):
        e.run(moves)
        if dist:
            gather(ll_tot_view)

    online_nccl(W)
    ms_online_nccl = timed(lambda: online_nccl(K))
    batch_nccl(W)
    ms_batch_nccl = timed(lambda: batch_nccl(K))
    res = {"workload": "sine_em APF bootstrap systematic, 1024 theta x 4096 particles (BASELINE.json configs[4]), theta columns sharded over the ranks",
           "scaling": "strong", "theta_per_gpu": hi - lo, "moves": K}
    ms_online, ms_batch, how = ms_online_nccl, ms_batch_nccl, "none (1 GPU)"
    if dist:
        res["nccl_exchange"] = {"batch_us_per_move": ms_batch_nccl * 1e3 / K, "online_us_per_move": ms_online_nccl * 1e3 / K,
                                "what": "torch.distributed all_gather_into_tensor of the (B_local,) values after every launch"}
        how = "NCCL all_gather_into_tensor"
        try:
            # (2) the exchange inside the finalising kernel: (value, tag) stores into every rank's buffer over NVLink peer memory
            from pyfilter_b200.sharding import PeerExchange

            px = PeerExchange(e, THETA, lo)

            def online_peer(moves):
                e.run_stepwise(moves)   # one launch per move + the reader of the exchange after every move, the loop in C

            def batch_peer(moves):
                e.run(moves)
                px.wait()

            online_peer(W)
            ms_online = timed(lambda: online_peer(K))
            batch_peer(W)
            ms_batch = timed(lambda: batch_peer(K))
            how = "peer-memory stores from the finalising kernel (no collective launch) + one reader kernel per exchange"
            # every rank must hold the same totals
            e.run(1)
            _, tot = px.wait()
            chk = tot.clone()
            dist.all_reduce(chk, op=dist.ReduceOp.MAX)
            res["peer_exchange_consistent"] = bool(torch.equal(chk, tot))
        except Exception as ex:
            res["peer_exchange_error"] = repr(ex)[:300]
    res["batch"] = {"us_per_move": ms_batch * 1e3 / K, "value": THETA * N * K / (ms_batch * 1e-3), "unit": UNIT}
    res["online"] = {"us_per_move": ms_online * 1e3 / K, "value": THETA * N * K / (ms_online * 1e-3), "unit": UNIT}
    res["exchange"] = how
    # speed-up against this box's own 1-GPU run (the driver runs N = 1, 2, 4, 8 back to back on one box)
    memo = os.path.join("/tmp", "smcb_smc2_n1.json")
    if rank == 0:
        try:
            if world == 1:
                json.dump({"batch": ms_batch / K, "online": ms_online / K}, open(memo, "w"))
            elif os.path.exists(memo):
                one = json.load(open(memo))
                res["speedup_vs_own_1gpu"] = {"batch": one["batch"] / (ms_batch / K), "online": one["online"] / (ms_online / K)}
        except Exception:
            pass
    return res


def operators_leg():
    """The Level-1 drop-ins on their own: stand-alone `systematic` / `multinomial` / `normalize` on a caller's
    tensor, device-timed; algorithmic bytes = weights in (4 B) + int64 ancestors out (8 B) per particle (normalize: 4 + 4)."""
    import torch
    import pyfilter_b200 as pf

    peak, _ = load_peaks()
    out = {}
    gen = torch.Generator().manual_seed(1)
    for name, n, fn, bpp in (("systematic_4M", 4_000_000, lambda w: pf.resampling.systematic(w, normalized=True), 12),
                             ("multinomial_2M", 2_000_000, lambda w: pf.resampling.multinomial(w, normalized=True), 12),
                             ("residual_2M", 2_000_000, lambda w: pf.resampling.residual(w, normalized=True), 12),
                             ("normalize_4M", 4_000_000, lambda w: pf.utils.normalize(w), 8)):
        lw = (torch.randn(n, generator=gen) * 2.0).cuda()
        w = pf.utils.normalize(lw.clone()) if name != "normalize_4M" else lw
        fn(w); fn(w)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        ev0.record()
        for _ in range(reps):
            fn(w)
        ev1.record()
        torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) * 1e3 / reps
        out[name] = {"us_per_call": us, "GB/s": bpp * n / (us * 1e-6) / 1e9, "roofline_frac": bpp * n / (us * 1e-6) / 1e9 / peak,
                     "algorithmic_bytes_per_particle": bpp}
    return out


def smc2_full_leg(T=40):
    """The whole SMC2 algorithm (pyfilter_b200.inference.SMC2: filter moves, theta-level ESS test with its host synchronisation, PMMH
    rejuvenation with re-filtering, accept / exchange) on BASELINE.json configs[4]'s sizes on ONE GPU: wall time per observation."""
    import torch
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import proposals
    from pyfilter_b200.inference import SMC2, LogNormal, Normal

    THETA, N = 1024, 4096
    y = simulate("sine_em", T)
    alg = SMC2(lambda p: ts.build("sine_em", gamma=p["gamma"], sigma=p["sigma"]), {"gamma": Normal(0.0, 1.0), "sigma": LogNormal(0.0, 0.5)},
               particles=THETA, state_particles=N, proposal=proposals.LinearGaussianObservations(), threshold=0.2, seed=123, max_observations=T + 2)
    state = alg.initialize()
    warm = 5
    for yt in y[:warm]:               # warm-up: library load, first launches, and ONE forced rejuvenation (the first use of cuSOLVER /
        state = alg.step(yt, state)   # cuBLAS for the p x p Cholesky factor of the proposal costs about a second on its own)
    state = alg.rejuvenate(state)
    r0 = state.rejuvenations
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for yt in y[warm:]:
        state = alg.step(yt, state)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    T = T - warm + 1
    post = alg.posterior_mean(state)
    return {"workload": f"SMC2 {THETA} theta x {N} state particles (grows when the acceptance rate falls), sine_em APF LinearGaussianObservations, "
                        f"threshold 0.2, {T - 1} observations, one GPU", "ms_per_observation": dt / (T - 1) * 1e3,
            "rejuvenations": state.rejuvenations - r0, "acceptance": [round(a, 3) for a in state.acceptance], "state_particles_final": state.engine.N,
            "posterior_mean": {k: float(v) for k, v in post.items()}, "ess_final": state.ess[-1]}


def torch_cuda_baseline(N, moves=10):
    """The reference arithmetic (the same torch calls as pyfilter's APF + Bootstrap + systematic, as restated in the oracle) with every
    tensor on cuda:0 - torch's generic ATen kernels, dozens of launches per move: the GPU path pyfilter users have today."""
    import torch

    dev = torch.device("cuda")
    mu, phi, sv = -1.0, 0.97, 0.2
    y = simulate_sv(moves + 4).to(dev)
    g = torch.Generator(device=dev).manual_seed(123)
    x = mu + sv / (1.0 - phi * phi) ** 0.5 * torch.randn(N, device=dev, generator=g)
    lw = torch.zeros(N, device=dev)

    def obs_lp(yt, xx):  # Normal(0, exp(x / 2)).log_prob(y)
        return torch.distributions.Normal(0.0, (xx / 2.0).exp()).log_prob(yt)

    def move(x, lw, yt):
        z = torch.randn(N, device=dev, generator=g)
        u = torch.rand(1, device=dev, generator=g)
        W = torch.softmax(lw - lw.max(), 0)                       # normalize (utils.py:49-64)
        gpre = obs_lp(yt, mu + phi * (x - mu))                    # pre-weight at the transition mean (proposals/base.py:69-85)
        rw = gpre + lw
        Wr = torch.softmax(rw - rw.max(), 0)                      # resampling.py:10-11
        probs = (torch.arange(N, device=dev, dtype=torch.float32) + u) / N      # resampling.py:44-50
        cs = Wr.cumsum(0)
        cs[-1] = 1.0
        idx = torch.searchsorted(cs, probs).clamp_max(N - 1)
        xr = x[idx]                                               # apf.py:34
        xn = mu + phi * (xr - mu) + sv * z
        inc = obs_lp(yt, xn)
        lwn = inc - gpre[idx]                                     # apf.py:43
        ll = torch.logsumexp(lwn, 0) - torch.log(torch.tensor(float(N), device=dev)) + (W * gpre.exp()).sum().log()   # apf.py:44
        Wn = torch.softmax(lwn - lwn.max(), 0)                    # particle/state.py:95
        mean = (Wn * xn).sum()
        var = (Wn * (xn - mean) ** 2).sum()
        return xn, lwn, ll, mean, var

    for t in range(3):
        x, lw, *_ = move(x, lw, y[t])
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for t in range(moves):
        x, lw, *_ = move(x, lw, y[3 + t])
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    return {"value": N * moves / (ms * 1e-3), "unit": UNIT, "us_per_move": ms * 1e3 / moves, "moves": moves,
            "what": f"torch {torch.__version__} ATen kernels on cuda:0, APF + Bootstrap + systematic arithmetic of the reference (sv_ar1), {N} particles"}


def exact_weights_leg(N, y_dev, stream, K):
    """The rounding-free weight mode against the unrounded one: throughput with exact_weights=True, and how many ancestors of ONE
    move differ between the two modes from the same state (both are bit-exact for the weights they use, smcb200.h)."""
    import torch
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF

    fe = APF(ts.build("sv_ar1"), N, seed=123, exact_weights=True)
    ee = fe._get_engine(K + 40)
    ee.initialize()
    ee.set_observations(y_dev, 0)
    ee.run(10)
    ms = _timed_moves(ee, K, stream)
    out = {"exact_weights_true": {"value": N * K / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / K}}
    # one move from the same state in both modes
    x0, lw0, pi0 = ee.x_view().clone(), ee.logw_view().clone(), ee.prev_inds().clone()
    t0 = ee.t
    anc = []
    for exact in (True, False):
        f = APF(ts.build("sv_ar1"), N, seed=123, exact_weights=exact)
        e = f._get_engine(4)
        e.load_state(x0, lw0, pi0, t0)
        e.set_observations(y_dev, 0)
        e.run(1)
        torch.cuda.synchronize()
        anc.append(e.prev_inds().clone())
    out["quantised_vs_exact_flip_rate"] = float((anc[0] != anc[1]).float().mean())
    return out


def run_b200(args):
    import ctypes as C
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import pyfilter_b200 as pf  # noqa: F401
    from pyfilter_b200 import _lib, timeseries as ts
    from pyfilter_b200.filters.particle import APF

    lib = _lib.load_library()
    N, K, W = args.particles, args.steps, args.warmup
    P = min(K, 40)  # profiled moves
    y = simulate_sv(W + K + P + 2)
    f = APF(ts.build("sv_ar1"), N, seed=123, column_offset=rank)   # replicas: every rank its own random streams
    f._exact_weights = args.exact_weights
    e = f._get_engine(W + K + P + 4)
    y_dev = y.reshape(-1, 1).cuda()
    stream = torch.cuda.current_stream()
    e.initialize()
    e.set_observations(y_dev, 0)
    e.run(W)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    flush.fill_(1.0)  # 256 MB > 126 MB L2
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = e.info().kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    e.run(K)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    launches = e.info().kernel_launches - l0
    if dist:
        # replicas: no data-path collective; the marginal log-likelihoods of all replicas are gathered once (outside the kernels)
        from pyfilter_b200.sharding import gather_loglikelihood

        ll = e.raw(_lib.PTR_LL_TOTAL, (e.B,)).clone()
        gather_loglikelihood(ll, world * e.B)
        tmax = torch.tensor([ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax)
    clocks = sampler.stop()
    value = world * N * K / (ms * 1e-3)

    # ---- per-kernel device time (CUDA events on the launching stream) -> roofline of the dominant kernel
    prof = (C.c_float * 5)()
    _lib.check(lib.smcb_filter_profile(e.handle, P, prof, stream.cuda_stream))
    e.t += P
    fused = (not args.exact_weights) and N <= (1 << 23) and not os.environ.get("SMCB_NO_FUSED")  # run_one's own conditions
    move = (not args.exact_weights) and N <= (1 << 23) and not os.environ.get("SMCB_NO_MOVE")
    if move:     # ONE kernel per move (csrc/move.cuh) + its one-block finalize; the other slots only hold the bracketing events' overhead
        names = ["apf_preweight", "move_kernel", "-", "-", "-"]
    else:
        names = ["apf_preweight", "resample_fused_kernel" if fused else "normalize_kernel", "describe_kernel", "expand_kernel", "step_kernel"]
    per = {n_: prof[i] / P for i, n_ in enumerate(names) if n_ != "-"}
    bracketed = dict(per)
    if fused and not move:
        per.pop("describe_kernel"); per.pop("expand_kernel")
    if move:
        # one move = ONE launch of the dominant kernel (plus its one-block finalize, chained with programmatic dependent launch).  The
        # events of smcb_filter_profile sit between the kernels and break that overlap, so the kernel's average launch duration is
        # taken from the timed region itself: K back-to-back moves on the launching stream (the bracketed figure is kept next to it)
        per["move_kernel"] = ms / K
    # algorithmic bytes per launch: whole move = 16 + 8 d per particle; resampling = load log-weight 4 + store ancestor 4;
    # step = ancestor 4 + gather 4 + x 4 + log-weight 4
    alg_bytes = {"move_kernel": 24.0 * N, "resample_fused_kernel": 8.0 * N, "normalize_kernel": 8.0 * N, "describe_kernel": 4.0 * N,
                 "expand_kernel": 8.0 * N, "step_kernel": 16.0 * N, "apf_preweight": 12.0 * N}
    dom = max(per, key=per.get)
    peak, peak_src = load_peaks()
    achieved = alg_bytes[dom] / (per[dom] * 1e-3) / 1e9 if per[dom] > 0 else 0.0
    step_ms = sum(per.values())
    whole = 24.0 * N / (ms / K * 1e-3) / 1e9
    traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/traffic.json)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if int(tj["particles"]) == N:
            traffic = tj["bytes_per_launch"].get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel_ms": per, "kernel_ms_event_bracketed": bracketed,
                "how": "algorithmic bytes (24 B x particles: SURVEY.md 8(d)) / average launch duration of the dominant kernel over the timed region "
                       "(CUDA events on the launching stream around K back-to-back moves; one launch per move)" if move else
                       "algorithmic bytes / CUDA-event duration of the dominant kernel (smcb_filter_profile)",
                "whole_step": {"algorithmic_bytes_per_particle": 24, "achieved": whole, "frac": whole / peak,
                               "ms_per_step_profiled": step_ms}}
    try:  # every kernel of the move against the same peak
        roofline["per_kernel"] = {k: {"ms": per[k], "algorithmic_bytes": alg_bytes[k],
                                      "achieved": alg_bytes[k] / (per[k] * 1e-3) / 1e9, "frac": alg_bytes[k] / (per[k] * 1e-3) / 1e9 / peak}
                                  for k in per if k != "apf_preweight" and per[k] > 0.005}  # skip slots that only hold event overhead
    except Exception:
        pass

    # ---- the other workloads BASELINE.json names (every rank takes part: replicas / theta shards)
    extra = {}
    if not args.headline_only:
        del flush
        try:
            extra["configs"] = bench_configs(world, dist, moves=200)
        except Exception as ex:
            extra["configs"] = {"error": repr(ex)[:200]}
        try:
            extra["smc2_shard"] = bench_smc2(world, rank, dist)
        except Exception as ex:
            extra["smc2_shard"] = {"error": repr(ex)[:200]}

    # ---- end to end through the C ABI on host buffers (pinned), incl. H2D of y and D2H of the results: every rank runs its replica,
    #      barrier on both sides, the slowest rank's wall time counts, the value is the whole job's
    yk = y[:K].contiguous().pin_memory()
    means = torch.empty(K + 1, 1, 1).pin_memory()
    varis = torch.empty(K + 1, 1, 1).pin_memory()
    lls = torch.empty(K + 1, 1).pin_memory()
    tot = torch.empty(1).pin_memory()
    call = lambda T_: _lib.check(lib.smcb_filter_batch_filter_host(e.handle, yk.data_ptr(), T_, means.data_ptr(), varis.data_ptr(),
                                                                   lls.data_ptr(), tot.data_ptr(), stream.cuda_stream))
    call(min(K, 3))
    l1 = e.info().kernel_launches
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    call(K)
    dt = time.perf_counter() - t0
    if dist:
        tmax = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dt = float(tmax)
    if rank == 0:
        e2e = {"value": world * N * K / dt, "unit": UNIT, "h2d_bytes_per_step": 4 * world, "d2h_bytes_per_step": (12 + 4.0 / K) * world,
               "api": "smcb_filter_batch_filter_host (C ABI, host buffers), one call per rank, max over ranks", "loglikelihood": float(tot[0]),
               "gpu_launches": int(e.info().kernel_launches - l1)}
        if world == 1 and not args.headline_only:
            try:
                extra.update(exact_weights_leg(N, y_dev, stream, min(K, 200)))
            except Exception as ex:
                extra["exact_weights_true"] = {"error": repr(ex)[:200]}
            try:
                extra["operators"] = operators_leg()
            except Exception as ex:
                extra["operators"] = {"error": repr(ex)[:200]}
            try:
                extra["smc2_full"] = smc2_full_leg()
            except Exception as ex:
                extra["smc2_full"] = {"error": repr(ex)[:200]}
            try:
                extra["torch_cuda_baseline"] = torch_cuda_baseline(N)
            except Exception as ex:
                extra["torch_cuda_baseline"] = {"error": repr(ex)[:200]}
        cpu = cpu_baseline(N) if world == 1 and not args.no_cpu else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "particles": N, "filters_per_gpu": 1, "parallelism": f"replicas x{world}",
                           "exact_weights": args.exact_weights,
                           "l2": "256 MB flush before the timed loop; the moves of one filter are sequential, so its ~64 MB working "
                                 "set (two state and two weight rows) stays within reach of the 126 MB L2 between moves by construction"},
                "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "slow_tiles": int(e.info().slow_tiles)}
        line.update(extra)
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)  # T of BASELINE.json configs[2]
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--particles", type=int, default=4_000_000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--headline-only", action="store_true", help="only the headline workload (profiling runs)")
    ap.add_argument("--workload", default="config3", choices=["config3", "smc2"],
                    help="config3 (default): the full line; smc2: only the theta-sharded batch of configs[4] as its own line")
    ap.add_argument("--exact-weights", action="store_true", help="do not round the resampling weights to multiples of 2^-52 (smcb_config.exact_weights)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.warmup < 3:
        args.warmup = 3
    if args.workload == "smc2":
        import torch

        rank = int(os.environ.get("RANK", "0"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(local)
        dist = None
        if world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        res = bench_smc2(world, rank, dist, K=250 if args.steps == 2000 else args.steps, W=args.warmup)
        if rank == 0:
            print(json.dumps({"metric": "particle-steps/sec, SMC2 theta batch 1024 x 4096 (BASELINE.json configs[4])", "n_gpus": world, **res}), flush=True)
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return
    run_b200(args)


if __name__ == "__main__":
    main()
