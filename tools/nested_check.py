"""NestedProposal, diagnostic of the library's own draws: (A) inner normals injected, pick by the library's uniform; (B) exponentials
injected, inner normals by the library."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import smc_oracle as O
import pyfilter_b200 as pf
from pyfilter_b200 import timeseries as ts
from pyfilter_b200.filters.particle import APF, SISR, proposals
from pyfilter_b200.filters.particle.state import ParticleFilterPrediction

M = 50
torch.manual_seed(5)
name = "lg_ar1"
mo = O.build_model(name)
gen = torch.Generator().manual_seed(3)
N = 4096
loc, scale = mo.initial_loc_scale()
x = loc + scale * torch.randn((N,), generator=gen)
_, ysim = mo.simulate(3)
y = ysim[2].float()
p = proposals.NestedProposal(M).set_model(ts.build(name))
xs = ts.TimeseriesState(4, x.cuda(), torch.Size(()))
pred = ParticleFilterPrediction(xs, torch.zeros(N).cuda(), torch.full((N,), 1.0 / N).cuda(), None)
zs = torch.randn((M, N), generator=gen)
E = torch.empty(N, M).exponential_(1, generator=gen)
mean, sc = mo.mean_scale(x)
samples = mean + sc * zs
lp = mo.obs_log_prob(y, samples)
probs = lp.softmax(0)
e = p._engine_for(xs.value)
# reference: everything injected
e.set_nested_noise(zs, E)
new, w = p.sample_and_weight(y, pred)
rx, rw = O.nested_sample_and_weight(mo, y, x, (zs, E), M)
print("both injected: x mismatch", int(((new.value.cpu() - rx).abs() > 1e-6).sum()), "w maxdiff", float((w.cpu() - rw).abs().max()))
# A
e.set_nested_noise(zs, None)
new, w = p.sample_and_weight(y, pred)
nv = new.value.cpu()
best = (samples - nv).abs().argmin(0)
print("A: residual", float((samples.gather(0, best[None])[0] - nv).abs().max()), "w maxdiff", float((w.cpu() - rw).abs().max()))
print("A: best[:32]", best[:32].tolist())
print("A: frac best==0", float((best == 0).float().mean()), "best==M-1", float((best == M - 1).float().mean()))
c = probs.cumsum(0)
hi = c.gather(0, best[None])[0]; lo = hi - probs.gather(0, best[None])[0]
mid = (lo + hi) / 2
print("A: implied U mid: mean %.3f var %.3f (uniform: 0.5, 0.083)" % (float(mid.mean()), float(mid.var())), "shift of x", float((nv - mean).mean()), "oracle", float((rx - mean).mean()))
# B
e.set_nested_noise(None, E)
new, w = p.sample_and_weight(y, pred)
zb = (new.value.cpu() - mean) / sc
print("B: z_best mean %.3f var %.3f" % (float(zb.mean()), float(zb.var())), "oracle", float(((rx - mean) / sc).mean()), float(((rx - mean) / sc).var()))
print("B: logmeanexp w", float(torch.logsumexp(w.cpu(), 0)) - np.log(N), "oracle", float(torch.logsumexp(rw, 0)) - np.log(N), "w mean", float(w.mean()), float(rw.mean()))
# none injected
e.set_nested_noise(None, None)
new, w = p.sample_and_weight(y, pred)
zb = (new.value.cpu() - mean) / sc
print("C: z_best mean %.3f var %.3f" % (float(zb.mean()), float(zb.var())), "w mean", float(w.mean()))
