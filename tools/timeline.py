"""Diagnostics (SMCB_DEBUG_TIMELINE=1): where describe_kernel spends its time on general (non-benign) moves."""
import os, sys
os.environ["SMCB_DEBUG_TIMELINE"] = "1"
os.environ["SMCB_NO_FUSED"] = "1"   # the stamps live in the three-kernel pipeline (normalize / describe / expand) and the step kernel
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pyfilter_b200 import _lib, timeseries as ts
from pyfilter_b200.filters.particle import APF
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
T = 80
y = bench.simulate_sv(T + 30)
f = APF(ts.build("sv_ar1"), N, seed=1)
e = f._get_engine(T + 40)
e.initialize(); e.set_observations(y.reshape(-1, 1).cuda(), 0); e.run(20); torch.cuda.synchronize()
dbg = e.raw(20, (16,), "<i8")
for t in range(T):
    dbg.zero_(); dbg[0] = 2**62; dbg[8] = 2**62; dbg[11] = 2**62
    e.run(1); torch.cuda.synchronize()
    v = int(e.raw(21, (1,), "<i4")[0])
    d = dbg.cpu().tolist()
    if t < 6:
        print(f"move {t}: normalize main {(d[9]-d[8])/1e3:5.1f} tail {(d[10]-d[9])/1e3:5.1f} | step prologue {(d[14]-d[11])/1e3:5.1f} loop-end(max) {(d[15]-d[11])/1e3:5.1f} ticket {(d[12]-d[11])/1e3:5.1f} finalize {(d[13]-d[12])/1e3:5.1f} us")
    if not (v & 1):
        nt = (N + 4095) // 4096
        print(f"move {t}: describe first block start -> chain start {(d[1]-d[0])/1e3:6.1f} us; last block started at {(d[7]-d[0])/1e3:5.1f}; mean block: scan {d[3]/nt/1e3:5.2f} us, total {d[2]/nt/1e3:5.2f} us; chain {(d[6]-d[1])/1e3:6.1f} us; slowest block {(d[5]>>20)/1e3:5.1f} us (tile {d[5] & 0xfffff})")
