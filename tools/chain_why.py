import os, sys
os.environ["SMCB_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pyfilter_b200 as pf
from pyfilter_b200 import timeseries as ts
from pyfilter_b200.filters.particle import SISR, proposals
g = torch.Generator().manual_seed(123)
_, y = ts.build("lorenz63_em").sample_states(40, generator=g)
f = SISR(ts.build("lorenz63_em"), 2_000_000, proposal=proposals.Bootstrap(), resampling=pf.resampling.multinomial, seed=1)
e = f._get_engine(50)
e.initialize(); e.set_observations(y.float().reshape(40, -1).cuda().contiguous(), 0); e.run(20); torch.cuda.synchronize()
dbg = e.raw(20, (32,), "<i8")
for t in range(6):
    dbg.zero_(); dbg[0] = 2**62; e.run(1); torch.cuda.synchronize(); d = dbg.cpu().tolist()
    print(f"chain {(d[6]-d[1])/1e3:6.1f} us: P2 {d[27]/1e3:6.1f} us of which raw walks {d[26]/1e3:6.1f} us ({d[28]} walks); segments: {d[30]} runs, {d[29]} singles; why {d[16:22]}")
