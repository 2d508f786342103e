"""debug: the four variants of one move from the same state"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import smc_oracle as O
import pyfilter_b200 as pf
from pyfilter_b200 import _lib, timeseries as ts
from pyfilter_b200.filters.particle import APF, SISR
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
def eng():
    f = APF(ts.build("sv_ar1"), N, seed=77)
    return f._get_engine(24)
torch.manual_seed(5)
mo = O.build_model("sv_ar1")
_, y = mo.simulate(12)
yd = y.float().reshape(12, -1).cuda().contiguous()
e0 = eng(); e0.initialize(); e0.set_observations(yd, 0); e0.run(8); torch.cuda.synchronize()
x0, lw0, pi0 = e0.x_view().clone(), e0.logw_view().clone(), e0.prev_inds().clone()
out = {}
for path in ("move", "move_dump", "twokernel", "twokernel_dump", "move2"):
    if path.startswith("twokernel"): os.environ["SMCB_NO_MOVE"] = "1"
    else: os.environ.pop("SMCB_NO_MOVE", None)
    e = eng(); e.load_state(x0, lw0, pi0, 8); e.set_observations(yd, 0)
    eps = torch.zeros(e.D, e.B, e.ld, device="cuda"); ud = torch.zeros(e.B, device="cuda"); wd = torch.zeros(e.B, e.ld, device="cuda")
    if path.endswith("_dump"): e.dump_noise(eps, ud, wd)
    e.run(1); torch.cuda.synchronize()
    out[path] = dict(x=e.x_view().clone(), lw=e.logw_view().clone(), pi=e.prev_inds().clone(), u=ud, w=wd,
                     rw=e.raw(_lib.PTR_RESAMPLE_LOGW, (e.B, e.ld))[:, :N].clone())
names = list(out)
for i in range(len(names)):
    for j in range(i + 1, len(names)):
        a, b = out[names[i]], out[names[j]]
        print(names[i], names[j], "pi", int((a["pi"] != b["pi"]).sum()), "x", int((a["x"] != b["x"]).sum()), "lw", int((a["lw"] != b["lw"]).sum()),
              "rw", int((a["rw"] != b["rw"]).sum()))
d = out["move_dump"]
exp = O.systematic(d["w"][0, :N].cpu().clone().unsqueeze(1), normalized=True, u=d["u"].cpu().reshape(1, 1))[:, 0]
for k in names:
    print(k, "vs oracle systematic on dumped weights:", int((out[k]["pi"].cpu() != exp).sum()), "first diff", (out[k]["pi"].cpu() != exp).nonzero()[:3].reshape(-1).tolist())
print("w dumps equal:", torch.equal(out["move_dump"]["w"], out["twokernel_dump"]["w"]), "sum", float(d["w"].double().sum()))
