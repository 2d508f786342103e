"""Diagnostics: device time per move of one configuration for several tile geometries of move_kernel (SMCB_MV_GEOM="items1,items2,t1";
"" = the library's own choice).  usage: geom_sweep.py c3|c2|c4s [moves] [geom ...]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyfilter_b200 as pf
from pyfilter_b200 import timeseries as ts
from pyfilter_b200.filters.particle import APF, SISR, proposals
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
geoms = sys.argv[3:] or [""]
CFG = {"c3": ("sv_ar1", APF, proposals.Bootstrap, 4_000_000),
       "c2": ("sine_em", APF, proposals.LinearGaussianObservations, 1_000_000),
       "c4s": ("lorenz63_em", SISR, proposals.Bootstrap, 2_000_000),
       "c3h": ("sv_ar1", APF, proposals.Bootstrap, 2_000_000),
       "c3q": ("sv_ar1", APF, proposals.Bootstrap, 500_000)}[which]
model, cls, prop, N = CFG
g = torch.Generator().manual_seed(123)
_, y = ts.build(model).sample_states(T + 30, generator=g)
yd = y.float().reshape(T + 30, -1).cuda().contiguous()
for geom in geoms:
    if geom: os.environ["SMCB_MV_GEOM"] = geom
    else: os.environ.pop("SMCB_MV_GEOM", None)
    f = cls(ts.build(model), N, proposal=prop(), seed=1)
    e = f._get_engine(T + 40)
    e.initialize(); e.set_observations(yd, 0); e.run(20); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e.initialize(); e.set_observations(yd, 0); e.run(10)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); e.run(T); ev1.record(); torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1) * 1e3 / T)
    print(json.dumps({"config": which, "geom": geom or "auto", "us_per_move": round(best, 2), "loglik": float(e.raw(6, (e.B,)).mean())}), flush=True)
    del e, f
