"""Diagnostics: per-move device time of every kernel group next to the column verdict (benign / general exact scan)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pyfilter_b200 import _lib, timeseries as ts
from pyfilter_b200.filters.particle import APF
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 120
y = bench.simulate_sv(T + 30)
f = APF(ts.build("sv_ar1"), N, seed=1)
e = f._get_engine(T + 40)
lib = _lib.load_library()
e.initialize(); e.set_observations(y.reshape(-1, 1).cuda(), 0); e.run(20); torch.cuda.synchronize()
stream = torch.cuda.current_stream().cuda_stream
rows = []
for t in range(T):
    prof = (C.c_float * 5)()
    _lib.check(lib.smcb_filter_profile(e.handle, 1, prof, stream)); e.t += 1
    v = int(e.raw(21, (1,), "<i4")[0])
    rows.append((v, [prof[i] * 1e3 for i in range(5)]))
import statistics as st
for tag, sel in (("benign", 1), ("general", 0)):
    r = [x[1] for x in rows if (x[0] & 1) == sel]
    if r:
        print(f"{tag:8s} moves {len(r):4d}  normalize {st.mean(a[1] for a in r):6.1f}  describe {st.mean(a[2] for a in r):6.1f} (max {max(a[2] for a in r):.1f})  "
              f"expand {st.mean(a[3] for a in r):6.1f} (max {max(a[3] for a in r):.1f})  step {st.mean(a[4] for a in r):6.1f} us")
print("all      moves %4d  total %.1f us/move" % (len(rows), st.mean(sum(a[1]) for a in rows)))
