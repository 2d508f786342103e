"""Where a move's time goes inside move_kernel (SMCB_DEBUG_TIMELINE=1): globaltimer stamps of thread 0 of every tile at the phase
boundaries: 0 start, 1 dependency awaited, 2 tile sum published, 3 sum of the predecessors known, 4 marks done, 5 ancestors emitted,
6 propagation done, 7 end (partial record / finalize)."""
import os, sys
os.environ["SMCB_DEBUG_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pyfilter_b200 import timeseries as ts
from pyfilter_b200.filters.particle import APF
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
y = bench.simulate_sv(60)
f = APF(ts.build("sv_ar1"), N, seed=1)
e = f._get_engine(64)
e.initialize(); e.set_observations(y.reshape(-1, 1).cuda(), 0); e.run(20); torch.cuda.synchronize()
T = (N + 4095) // 4096
dbg = e.raw(20, (32 + 16 * T,), "<i8")
wdg = e.raw(22, (4,), "<i8")
for rep in range(3):
    dbg.zero_(); wdg.zero_(); e.run(3); torch.cuda.synchronize()
    w = wdg.cpu().tolist()
    print(f"look-back over 3 moves: tiles that waited {w[0]}, re-polls {w[2]}, mean time in look-back {w[3] / (3 * T) / 1e3:.2f} us, tiles with >1 window {w[1]}")   # the stamps of the last of three back-to-back moves remain
    raw16 = dbg[32:].cpu().numpy().reshape(T, 16).astype(np.float64)
    tl = raw16[:, :8] / 1e3
    lb = raw16[:, 8:13] / 1e3
    lbw = raw16[:, 13]
    t0 = tl[:, 0].min()
    tl -= t0
    order = np.argsort(tl[:, 0])
    names = ["wait dep", "load+weights+scan+publish", "look-back", "counts+marks", "emit", "propagate", "partial/finalize"]
    print(f"move: first start 0, last start {tl[:,0].max():.1f} us, last end {tl[:,7].max():.1f} us; blocks started within 2 us: {(tl[:,0] < 2).sum()}")
    for lab, sel in (("first wave", tl[:, 0] < 2.0), ("later", tl[:, 0] >= 2.0)):
        if sel.sum() == 0: continue
        d = np.diff(tl[sel], axis=1)
        print(f"  {lab} ({sel.sum()} tiles): " + ", ".join(f"{n} {d[:, i].mean():.2f}" for i, n in enumerate(names)) + f"  | total {(tl[sel][:, 7] - tl[sel][:, 0]).mean():.2f} us")
    print("  end-time percentiles of propagate (stamp 6):", np.percentile(tl[:, 6], [10, 50, 90, 99, 100]).round(1), " finalize end:", tl[:, 7].max().round(1))
    w1 = tl[:, 0] < 2.0
    print("  first wave: dependency released (stamp 1) pct:", np.percentile(tl[w1, 1], [0, 10, 50, 90, 100]).round(2))
    print("  first wave: published (stamp 2) pct:", np.percentile(tl[w1, 2], [0, 10, 50, 90, 99, 100]).round(2))
    print("  first wave: predecessors known (stamp 3) pct:", np.percentile(tl[w1, 3], [0, 10, 50, 90, 99, 100]).round(2))
    idx = np.arange(T)[w1]
    for lo in range(0, 600, 100):
        s = (idx >= lo) & (idx < lo + 100)
        if s.sum(): print(f"    tiles {lo}-{lo+99}: publish {tl[w1][s, 2].mean():.2f}  known {tl[w1][s, 3].mean():.2f}  marks {tl[w1][s, 4].mean():.2f} end {tl[w1][s, 6].mean():.2f}")
    order = np.argsort(np.arange(T))
    maxpub = np.maximum.accumulate(tl[:, 2])
    delay = tl[:, 3] - maxpub
    print("  known minus latest publication among the predecessors (and itself), pct:", np.percentile(delay[w1], [10, 50, 90, 99, 100]).round(2),
          " later wave:", np.percentile(delay[~w1], [10, 50, 90, 100]).round(2))
    print("  own publication -> known, pct:", np.percentile((tl[:, 3] - tl[:, 2])[w1], [10, 50, 90, 100]).round(2))
    lb -= t0
    sel = w1 & (np.arange(T) >= 64)
    d = np.stack([lb[:, 0] - tl[:, 2], lb[:, 1] - lb[:, 0], lb[:, 2] - lb[:, 1], lb[:, 3] - lb[:, 2], lb[:, 4] - lb[:, 3], tl[:, 3] - lb[:, 4]], 1)
    print("  inside the look-back (first wave, tiles >= 64): publish+count issued %.2f, counter back %.2f, group closed %.2f, group totals in %.2f, mates in %.2f, barrier %.2f; re-polls of lane 0: %.1f"
          % (tuple(d[sel].mean(0)) + (lbw[sel].mean(),)))
    sel = ~w1
    print("  (later wave) %.2f %.2f %.2f %.2f %.2f %.2f; re-polls of lane 0: %.1f" % (tuple(d[sel].mean(0)) + (lbw[sel].mean(),)))
