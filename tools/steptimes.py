"""Diagnostics: per-move device time of every kernel group (CUDA events around the groups, so a few us of event overhead each).
usage: steptimes.py [c3|c4|c4s|c2|c5] [moves]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import statistics as st
import pyfilter_b200 as pf
from pyfilter_b200 import _lib, timeseries as ts
from pyfilter_b200.filters.particle import APF, SISR, proposals
which = sys.argv[1] if len(sys.argv) > 1 else "c3"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
CFG = {"c3": ("sv_ar1", APF, proposals.Bootstrap, pf.resampling.systematic, 4_000_000, ()),
       "c2": ("sine_em", APF, proposals.LinearGaussianObservations, pf.resampling.systematic, 1_000_000, ()),
       "c4": ("lorenz63_em", SISR, proposals.Bootstrap, pf.resampling.multinomial, 2_000_000, ()),
       "c4s": ("lorenz63_em", SISR, proposals.Bootstrap, pf.resampling.systematic, 2_000_000, ()),
       "c5": ("sine_em", APF, proposals.Bootstrap, pf.resampling.systematic, 4096, (128,))}[which]
model, cls, prop, res, N, batch = CFG
g = torch.Generator().manual_seed(123)
_, y = ts.build(model).sample_states(T + 30, generator=g)
f = cls(ts.build(model), N, proposal=prop(), resampling=res, seed=1)
if batch:
    f.set_batch_shape(torch.Size(batch))
e = f._get_engine(T + 40)
lib = _lib.load_library()
e.initialize(); e.set_observations(y.float().reshape(T + 30, -1).cuda().contiguous(), 0); e.run(20); torch.cuda.synchronize()
stream = torch.cuda.current_stream().cuda_stream
rows = []
for t in range(T):
    prof = (C.c_float * 5)()
    _lib.check(lib.smcb_filter_profile(e.handle, 1, prof, stream)); e.t += 1
    rows.append([prof[i] * 1e3 for i in range(5)])
names = ["preweight", "normalize/fused", "describe", "expand(+draw)", "step"]
print(which, "moves", T, " ".join(f"{n} {st.mean(r[i] for r in rows):.1f} (max {max(r[i] for r in rows):.1f})" for i, n in enumerate(names)),
      "total %.1f us/move" % st.mean(sum(r) for r in rows), "slow tiles", e.info().slow_tiles)
