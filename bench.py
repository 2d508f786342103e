#!/usr/bin/env python
"""bench.py - particle-steps/sec of the SMC inner loop (BASELINE.json metric) on N GPUs of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--particles P]

Workload (BASELINE.json configs[2], the configuration the metric "at 4M particles (1D SSM)" is quoted on): stochastic-volatility
model, APF + Bootstrap proposal, systematic resampling every move, 4,000,000 particles, one filter per GPU (replicas: the particle
dimension of one filter does not shard without an exchange step; the theta-batch does, and is weak-scaled the same
way).  A "step" is one filter move (predict + resample + propagate + weight + moments) over all particles.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes through the C-ABI call on HOST
buffers (smcb_filter_batch_filter_host: H2D of the observations, all moves, D2H of moments and likelihoods inside the timed region);
`roofline` is for the dominant kernel, from algorithmic bytes over its CUDA-event duration; `cpu_baseline` times
the oracle port of the reference's torch-CPU path on this box's host cores on a bounded sample of the same workload.
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


def simulate_sv(T, seed=123):
    import torch
    from pyfilter_b200 import timeseries as ts

    g = torch.Generator().manual_seed(seed)
    _, y = ts.build("sv_ar1").sample_states(T, generator=g)
    return y.float().contiguous()


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


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port on torch CPU; the reference itself is pure
    Python on torch and cannot travel to the GPU box) on this box's host cores, same metric/config."""
    import torch
    from oracle import smc_oracle as O

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = O.build_model("sv_ar1")
    K, W = args.steps, args.warmup
    y = simulate_sv(W + K)
    # bounded sample: shrink the particle count until K + W moves fit in ~2 minutes (throughput is flat in N at this size)
    n = args.particles
    torch.manual_seed(123)
    t0 = time.perf_counter()
    O.batch_filter(model, "apf", "bootstrap", y[:1], min(n, 1_000_000))
    per_particle = (time.perf_counter() - t0) / min(n, 1_000_000)
    while n > 250_000 and per_particle * n * (K + W) > 120.0:
        n //= 2
    st = O.batch_filter(model, "apf", "bootstrap", y[:W], n) if W else None
    x0 = st["x"] if st else None
    t0 = time.perf_counter()
    O.batch_filter(model, "apf", "bootstrap", y[W:W + K], n, x0=x0)
    dt = time.perf_counter() - t0
    value = n * K / dt
    sample = f"oracle port of pyfilter APF.batch_filter on torch CPU, {cores} threads, sv_ar1, {n} particles x {K} moves per run"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": "sv_ar1 APF bootstrap systematic (BASELINE configs[2])", "particles": n,
                                            "particles_requested": args.particles},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
    import pyfilter_b200 as pf
    from pyfilter_b200 import _lib, timeseries as ts
    from pyfilter_b200.filters.particle import APF

    lib = _lib.load_library()
    N, K, W = args.particles, args.steps, args.warmup
    P = min(K, 40)  # profiled moves
    y = simulate_sv(W + K + P + 2)
    f = APF(ts.build("sv_ar1"), N, seed=123 + rank)
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
        # the one collective of the theta-sharded loop: marginal log-likelihoods of all replicas (outside the kernels' data path)
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
    if move:     # ONE kernel per move (csrc/move.cuh); the other slots only hold the bracketing events' own overhead
        names = ["apf_preweight", "move_kernel", "-", "-", "-"]
    else:
        names = ["apf_preweight", "resample_fused_kernel" if fused else "normalize_kernel", "describe_kernel", "expand_kernel", "step_kernel"]
    per = {n_: prof[i] / P for i, n_ in enumerate(names) if n_ != "-"}
    if fused and not move:
        per.pop("describe_kernel"); per.pop("expand_kernel")
    # algorithmic bytes per launch: resampling = load log-weight 4 + store ancestor 4; step = ancestor 4 + gather 4 + x 4 + log-weight 4
    alg_bytes = {"move_kernel": 24.0 * N, "resample_fused_kernel": 8.0 * N, "normalize_kernel": 8.0 * N, "describe_kernel": 4.0 * N, "expand_kernel": 8.0 * N,
                 "step_kernel": 16.0 * N, "apf_preweight": 12.0 * N}
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
                "traffic": traffic, "peak_source": peak_src, "kernel_ms": per,
                "whole_step": {"algorithmic_bytes_per_particle": 24, "achieved": whole, "frac": whole / peak,
                               "ms_per_step_profiled": step_ms}}

    try:  # every kernel of the move against the same peak (the two hot kernels are within 2 % of each other: "dominant" can flip)
        roofline["per_kernel"] = {k: {"ms": per[k], "algorithmic_bytes": alg_bytes[k],
                                      "achieved": alg_bytes[k] / (per[k] * 1e-3) / 1e9, "frac": alg_bytes[k] / (per[k] * 1e-3) / 1e9 / peak}
                                  for k in per if k != "apf_preweight" and per[k] > 0.005}  # skip slots that only hold event overhead
    except Exception:
        pass

    line = None
    if rank == 0:
        # ---- end to end through the C ABI on host buffers (pinned), incl. H2D of y and D2H of the results
        yk = y[:K].contiguous().pin_memory()
        means = torch.empty(K + 1, 1, 1).pin_memory()
        varis = torch.empty(K + 1, 1, 1).pin_memory()
        lls = torch.empty(K + 1, 1).pin_memory()
        tot = torch.empty(1).pin_memory()
        call = lambda T_: _lib.check(lib.smcb_filter_batch_filter_host(e.handle, yk.data_ptr(), T_, means.data_ptr(), varis.data_ptr(),
                                                                       lls.data_ptr(), tot.data_ptr(), stream.cuda_stream))
        call(min(K, 3))
        l1 = e.info().kernel_launches
        t0 = time.perf_counter()
        call(K)
        dt = time.perf_counter() - t0
        e2e = {"value": N * K / dt, "unit": UNIT, "h2d_bytes_per_step": 4, "d2h_bytes_per_step": 12 + 4.0 / K,
               "api": "smcb_filter_batch_filter_host (C ABI, host buffers)", "loglikelihood": float(tot[0]),
               "gpu_launches": int(e.info().kernel_launches - l1)}
        cpu = cpu_baseline(N) if world == 1 and not args.no_cpu else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "sv_ar1 APF bootstrap systematic, 4M particles (BASELINE.json configs[2])", "particles": N,
                           "filters_per_gpu": 1, "parallelism": f"replicas x{world}", "exact_weights": args.exact_weights,
                           "l2": "256 MB flush before the timed loop; the ~80 MB working set of one filter is L2-resident across "
                                 "moves by construction (the moves of one filter are sequential)"},
                "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "slow_tiles": int(e.info().slow_tiles), "lb_fail": int(e.info().lb_fail), "lb_windows": int(e.info().lb_windows)}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def run_smc2_shard(args):
    """Secondary workload (not the driver's default line): BASELINE.json configs[4], the SMC2 / NESS batch of independent filters -
    1024 theta x 4096 state particles, sine diffusion, APF + Bootstrap, systematic - STRONG scaling: the theta columns are block
    distributed over the ranks (pyfilter_b200.sharding.column_shard), every rank runs the resident column kernel on its shard, and the
    per-move exchange is the all-gather of the (B_local,) log-likelihood increments the theta-level ESS test needs.
    Two timings: "filter" = one move per launch + the NCCL all-gather after every move (what SMC2's online loop does), "batch" = all K
    moves in one launch + one all-gather (batch_filter inside PMMH / the initial SMC2 sweep)."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pyfilter_b200 import _lib, timeseries as ts
    from pyfilter_b200.filters.particle import APF
    from pyfilter_b200.sharding import LogLikelihoodGather, column_shard

    THETA, N, K, W = 1024, 4096, args.steps, args.warmup
    lo, hi = column_shard(THETA, rank, world)
    Bl = hi - lo
    torch.manual_seed(123)
    gamma, sigma = torch.randn(THETA), torch.exp(0.5 * torch.randn(THETA))   # theta ~ prior
    g = torch.Generator().manual_seed(123)
    _, y = ts.build("sine_em").sample_states(2 * (W + K) + 4, generator=g)
    y_dev = y.float().reshape(-1, 1).cuda().contiguous()
    f = APF(ts.build("sine_em", gamma=gamma[lo:hi], sigma=sigma[lo:hi]), N, seed=123)
    f.set_batch_shape(torch.Size([Bl]))
    e = f._get_engine(2 * (W + K) + 8)
    stream = torch.cuda.current_stream()
    e.initialize()
    e.set_observations(y_dev, 0)

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn):
        sync_all()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream); fn(); ev1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    ll_view, ll_tot_view = e.raw(_lib.PTR_LL, (e.B,)), e.raw(_lib.PTR_LL_TOTAL, (e.B,))
    gather = LogLikelihoodGather(THETA, "cuda") if dist else None

    def online(moves):
        for _ in range(moves):
            e.run(1)
            if dist:
                gather(ll_view)

    def batch(moves):
        e.run(moves)
        if dist:
            gather(ll_tot_view)

    online(W)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = e.info().kernel_launches
    ms_online = timed(lambda: online(K))
    launches = e.info().kernel_launches - l0
    batch(W)
    ms_batch = timed(lambda: batch(K))
    clocks = sampler.stop()
    if rank == 0:
        line = {"metric": "particle-steps/sec, SMC2 theta batch 1024 x 4096 (BASELINE.json configs[4])", "value": THETA * N * K / (ms_batch * 1e-3),
                "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_batch / K, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "sine_em APF bootstrap systematic, 1024 theta x 4096 particles, theta columns sharded over the ranks",
                           "theta_per_gpu": Bl, "parallelism": f"theta-shard x{world}", "kernel": "column_kernel (resident column)",
                           "l2": "working set per GPU (<= 50 MB) is on chip by design: the columns stay in shared memory between moves"},
                "online": {"value": THETA * N * K / (ms_online * 1e-3), "ms_per_step": ms_online / K,
                           "what": "one move per launch + all-gather of the log-likelihood increments after every move"},
                "gpu_launches": int(launches), "clocks": clocks}
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
    ap.add_argument("--workload", default="config3", choices=["config3", "smc2"],
                    help="config3 (default, the driver's line): 4M-particle SV APF; smc2: the theta-sharded batch of configs[4], strong scaling")
    ap.add_argument("--exact-weights", action="store_true", help="do not round the resampling weights to multiples of 2^-52 (smcb_config.exact_weights)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        if args.workload == "smc2":
            if args.steps == 2000:
                args.steps = 250   # T of configs[4]
            run_smc2_shard(args)
        else:
            run_b200(args)


if __name__ == "__main__":
    main()
