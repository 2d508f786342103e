"""Diagnostics: the resident column kernel's block shapes on one GPU's theta-shard (4096 particles x B columns):
SMCB_COLUMN_MINB = 0 (1024 threads x 4 particles), 1 (512 x 8, one block per SM), 2 (512 x 8, two per SM).  usage: column_variants.py [B ...]"""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "one":
    import torch
    import pyfilter_b200 as pf
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF
    B, T = int(sys.argv[2]), 300
    torch.manual_seed(123)
    g = torch.Generator().manual_seed(123)
    _, y = ts.build("sine_em").sample_states(T + 24, generator=g)
    f = APF(ts.build("sine_em", gamma=torch.randn(B), sigma=torch.exp(0.5 * torch.randn(B))), 4096, seed=7)
    f.set_batch_shape(torch.Size([B]))
    e = f._get_engine(T + 30)
    yd = y.float().reshape(-1, 1).cuda().contiguous()
    best = 1e9
    for rep in range(3):
        e.initialize(); e.set_observations(yd, 0); e.run(20)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); ev0.record(); e.run(T); ev1.record(); torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1) * 1e3 / T)
    print(json.dumps({"B": B, "minb": os.environ.get("SMCB_COLUMN_MINB", "auto"), "us_per_move": round(best, 3)}))
else:
    for B in [int(a) for a in sys.argv[1:]] or [128]:
        for v in ("auto", "0", "1", "2"):
            env = dict(os.environ)
            if v != "auto": env["SMCB_COLUMN_MINB"] = v
            r = subprocess.run([sys.executable, __file__, "one", str(B)], env=env, capture_output=True, text=True)
            print(r.stdout.strip()[-200:], r.stderr.strip()[-300:] if r.returncode else "", flush=True)
