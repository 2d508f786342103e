"""Diagnostics: build variants of libsmcb200.so with extra -D flags (into build/, git-ignored but shipped to the GPU box) and time
configs[2] with each.   build here:  variants.py build name1:-DX=1,-DY=2 name2:...      run on the box:  variants.py run [moves]"""
import os, subprocess, sys, glob, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BUILD = os.path.join(ROOT, "build")


def build(specs):
    from pyfilter_b200 import _lib
    os.makedirs(BUILD, exist_ok=True)
    for f in glob.glob(os.path.join(BUILD, "var_*.so")):
        os.remove(f)
    procs = []
    for spec in specs:
        name, _, flags = spec.partition(":")
        cmd = ["/usr/local/cuda/bin/nvcc"] + _lib.NVCC_FLAGS + [x for x in flags.split(",") if x] + \
              ["-o", os.path.join(BUILD, f"var_{name}.so"), os.path.join(ROOT, "pyfilter_b200", "csrc", "smcb_api.cu")]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out, _ = p.communicate()
        print(name, "rc", p.returncode, out[-2000:] if p.returncode else "")


def run(moves, cfg):
    for so in sorted(glob.glob(os.path.join(BUILD, "var_*.so"))):
        env = dict(os.environ, SMCB_LIB_PATH=so)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "variants.py"), "one", str(moves), cfg], env=env, capture_output=True, text=True)
        print(os.path.basename(so), r.stdout.strip()[-300:], r.stderr.strip()[-300:] if r.returncode else "", flush=True)


def one(T, cfg):
    import torch
    import pyfilter_b200 as pf
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, SISR, proposals
    model, cls, prop, res, N, *batch = {"c3": ("sv_ar1", APF, proposals.Bootstrap, pf.resampling.systematic, 4_000_000),
                                "c1": ("lg_ar1", SISR, proposals.Bootstrap, pf.resampling.systematic, 1_000),
                                "c5s": ("sine_em", APF, proposals.Bootstrap, pf.resampling.systematic, 4096, 128),
                                "c5": ("sine_em", APF, proposals.Bootstrap, pf.resampling.systematic, 4096, 1024),
                                "c2": ("sine_em", APF, proposals.LinearGaussianObservations, pf.resampling.systematic, 1_000_000),
                                "c4": ("lorenz63_em", SISR, proposals.Bootstrap, pf.resampling.multinomial, 2_000_000),
                                "c4s": ("lorenz63_em", SISR, proposals.Bootstrap, pf.resampling.systematic, 2_000_000)}[cfg]
    g = torch.Generator().manual_seed(123)
    _, y = ts.build(model).sample_states(T + 24, generator=g)
    kw = {}
    if batch:
        torch.manual_seed(123)
        kw = dict(gamma=torch.randn(batch[0]), sigma=torch.exp(0.5 * torch.randn(batch[0])))
    f = cls(ts.build(model, **kw), N, proposal=prop(), resampling=res, seed=7)
    if batch:
        f.set_batch_shape(torch.Size(batch))
    e = f._get_engine(T + 30)
    yd = y.float().reshape(T + 24, -1).cuda().contiguous()
    e.initialize(); e.set_observations(yd, 0); e.run(20); torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        if rep:
            e.initialize(); e.set_observations(yd, 0); e.run(20)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); ev0.record(); e.run(T); ev1.record(); torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1) * 1e3 / T)
    print(json.dumps({"cfg": cfg, "us_per_move": round(best, 2), "loglik": float(e.raw(6, (e.B,)).mean()), "slow": e.info().slow_tiles}))


if __name__ == "__main__":
    if sys.argv[1] == "build": build(sys.argv[2:])
    elif sys.argv[1] == "run": run(int(sys.argv[2]) if len(sys.argv) > 2 else 300, sys.argv[3] if len(sys.argv) > 3 else "c3")
    else: one(int(sys.argv[2]), sys.argv[3])
