"""Diagnostics: per-tile timeline of the scan kernel (SMCB_DEBUG_TIMELINE=1)."""
import os, sys
os.environ["SMCB_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pyfilter_b200 import _lib, timeseries as ts
from pyfilter_b200.filters.particle import APF
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
y = bench.simulate_sv(40)
f = APF(ts.build("sv_ar1"), N, seed=1)
e = f._get_engine(64)
e.initialize(); e.set_observations(y.reshape(-1, 1).cuda(), 0); e.run(20); torch.cuda.synchronize()
tiles = (N + 4095) // 4096
d = e.raw(20, (tiles, 8), "<i8").clone().cpu().double()
t0 = d[:, 0].min()
names = ["start", "phaseB_done", "spec_done", "bar3_passed", "end_spec_ok", "lb_start", "lb_end", "end_redo"]
print("kernel span us:", (d[:, [4, 7]].max() - t0).item() / 1e3)
for a, b, lab in [(0, 1, "load+phaseA+B"), (1, 2, "speculative expansion"), (2, 3, "wait for exact state"), (3, 4, "verify+writeout"), (5, 6, "lookback warp"), (0, 4, "tile residency")]:
    x = (d[:, b] - d[:, a]) / 1e3
    x = x[(d[:, b] > 0) & (d[:, a] > 0)]
    print(f"{lab:24s} mean {x.mean():8.2f} us  p50 {x.median():8.2f}  max {x.max():8.2f}  n={len(x)}")
redo = (d[:, 7] > d[:, 0]).sum().item()
print("tiles redone exactly:", redo, "of", tiles)
st = (d[:, 0] - t0) / 1e3
print("tile start times us: p10 %.1f p50 %.1f p90 %.1f max %.1f" % tuple(st.quantile(torch.tensor([0.1, 0.5, 0.9, 1.0], dtype=torch.double)).tolist()))
