"""Throughput of every BASELINE.json configuration on one GPU (device-timed, inputs resident): particle-steps/s.
usage: configs_bench.py [moves]   (the full T of the configs is hours of CPU baseline; the device rate is flat in T)"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyfilter_b200 as pf
from pyfilter_b200 import timeseries as ts
from pyfilter_b200.filters.particle import APF, SISR, proposals

T = int(sys.argv[1]) if len(sys.argv) > 1 else 200
CONFIGS = [
    ("c1 lg_ar1 SISR bootstrap systematic 1k", "lg_ar1", SISR, proposals.Bootstrap, pf.resampling.systematic, 1_000, ()),
    ("c2 sine_em APF LinearGaussianObservations systematic 1M", "sine_em", APF, proposals.LinearGaussianObservations, pf.resampling.systematic, 1_000_000, ()),
    ("c3 sv_ar1 APF bootstrap systematic 4M", "sv_ar1", APF, proposals.Bootstrap, pf.resampling.systematic, 4_000_000, ()),
    ("c4 lorenz63_em SISR bootstrap multinomial 2M", "lorenz63_em", SISR, proposals.Bootstrap, pf.resampling.multinomial, 2_000_000, ()),
    ("c4' lorenz63_em SISR bootstrap systematic 2M", "lorenz63_em", SISR, proposals.Bootstrap, pf.resampling.systematic, 2_000_000, ()),
    ("c5 sine_em APF bootstrap systematic 4096 x 128 theta (one GPU's shard of 1024)", "sine_em", APF, proposals.Bootstrap, pf.resampling.systematic, 4096, (128,)),
    ("c5 whole batch on one GPU 4096 x 1024 theta", "sine_em", APF, proposals.Bootstrap, pf.resampling.systematic, 4096, (1024,)),
]
for name, model, cls, prop, res, N, batch in CONFIGS:
    torch.manual_seed(123)
    kw = {}
    if batch:
        B = batch[0]
        kw = dict(gamma=torch.randn(B), sigma=torch.exp(0.5 * torch.randn(B)))
    m = ts.build(model, **kw)
    g = torch.Generator().manual_seed(123)
    _, y = ts.build(model).sample_states(T + 24, generator=g)
    f = cls(m, N, proposal=prop(), resampling=res, seed=7)
    if batch:
        f.set_batch_shape(torch.Size(batch))
    e = f._get_engine(T + 30)
    e.initialize()
    e.set_observations(y.float().reshape(T + 24, -1).cuda().contiguous(), 0)
    e.run(20)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); e.run(T); ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    nb = batch[0] if batch else 1
    info = e.info()
    print(json.dumps({"config": name, "particle_steps_per_s": N * nb * T / (ms * 1e-3), "us_per_move": ms * 1e3 / T, "moves": T,
                      "slow_tiles": info.slow_tiles, "loglik_mean": float(e.raw(6, (e.B,)).mean())}), flush=True)
