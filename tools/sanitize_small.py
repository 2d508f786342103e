"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): fused default path, three-kernel general
path (unrounded weights), multinomial, batched SISR with conditional resampling, resident column kernel, stand-alone operators."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pyfilter_b200 as pf
from pyfilter_b200 import timeseries as ts
from pyfilter_b200.filters.particle import APF, SISR, proposals

torch.manual_seed(0)
for name, cls, prop, res, N, B, kw in [
    ("sv_ar1", APF, proposals.Bootstrap, pf.resampling.systematic, 9000, (), {}),
    ("sv_ar1", APF, proposals.Bootstrap, pf.resampling.systematic, 9000, (), {"exact_weights": True}),
    ("sine_em", APF, proposals.LinearGaussianObservations, pf.resampling.systematic, 5000, (3,), {}),
    ("lorenz63_em", SISR, proposals.Bootstrap, pf.resampling.multinomial, 6000, (), {}),
    ("lg_ar1", SISR, proposals.Bootstrap, pf.resampling.systematic, 4500, (5,), {"ess_threshold": 0.5}),
    # resident column kernel (n <= 4096): full and ragged columns, APF fold / SISR conditional resampling, 3-D state, > 148 columns
    ("sine_em", APF, proposals.Bootstrap, pf.resampling.systematic, 4096, (3,), {}),
    ("sv_ar1", APF, proposals.Bootstrap, pf.resampling.systematic, 1001, (), {}),
    ("lorenz63_em", SISR, proposals.Bootstrap, pf.resampling.systematic, 3000, (2,), {"ess_threshold": 0.5}),
    ("lg_ar1", SISR, proposals.Bootstrap, pf.resampling.systematic, 130, (150,), {"ess_threshold": 0.7}),
]:
    g = torch.Generator().manual_seed(1)
    _, y = ts.build(name).sample_states(8, generator=g)
    f = cls(ts.build(name), N, proposal=prop(), resampling=res, seed=3, **kw)
    f.set_batch_shape(torch.Size(B))
    r = f.batch_filter(y, bar=False)
    print(name, cls.__name__, res.__name__, B, kw, "ll", r.loglikelihood.flatten()[:2].tolist())
w = torch.randn(7000, 2, device="cuda") * 4
print(pf.resampling.systematic(w.clone()).shape, pf.resampling.multinomial(w.clone()).shape, pf.utils.normalize(w.clone()).sum(0).tolist())
torch.cuda.synchronize()
