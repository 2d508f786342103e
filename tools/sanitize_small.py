"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): fused default path, three-kernel general
path (unrounded weights), multinomial, batched SISR with conditional resampling, resident column kernel, stand-alone operators; round
2: move_kernel (several tiles and groups, both tile classes, SISR identity moves), the online column path with the in-kernel APF
pre-weight, the plug-in passes, predict_path, smoothing (fixed-lag, FFBS), residual, column resample / exchange / export / import, one
SMC2 rejuvenation."""
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
# ---- round 2
from pyfilter_b200.inference import SMC2, LogNormal, Normal
g = torch.Generator().manual_seed(2)
_, y = ts.build("sine_em").sample_states(10, generator=g)
f = APF(ts.build("sv_ar1"), 150_000, seed=5)                                   # move_kernel: 37 tiles, two groups of 32
print("move", f.batch_filter(ts.build("sv_ar1").sample_states(6, generator=g)[1], bar=False).loglikelihood.tolist())
import os as _os
_os.environ["SMCB_MV_GEOM"] = "16,4,20"                                        # both tile classes in one column
f = SISR(ts.build("lorenz63_em"), 100_000, seed=5, ess_threshold=0.5)
print("move 3-D mixed classes", f.batch_filter(ts.build("lorenz63_em").sample_states(6, generator=g)[1], bar=False).loglikelihood.tolist())
_os.environ.pop("SMCB_MV_GEOM")
f = APF(ts.build("sine_em"), 2000, seed=5, record_states=True)                  # online: one launch per move, in-kernel pre-weight
st = f.initialize()
res = f.initialize_with_result(st)
for yt in y:
    st = f.filter(yt, st, result=res)
print("online", float(res.loglikelihood), f.smooth(res.states, method="fl").shape, f.smooth(res.states, method="ffbs").shape)
st2 = st.detach_copy()
pred = f.predict(st2)
print("split", float(f.correct(y[0], pred).get_loglikelihood()), st2.predict_path(f.ssm, 3).get_paths()[0].shape)
p = proposals.LinearGaussianObservations().set_model(ts.build("lorenz63_em"))
xs = ts.TimeseriesState(0, torch.randn(3000, 2, 3).cuda(), torch.Size([3]))
print("plugin", p.pre_weight(torch.zeros(2), xs).shape)
print("residual", pf.resampling.residual(torch.randn(9000, device="cuda")).shape)
alg = SMC2(lambda q: ts.build("sine_em", gamma=q["gamma"], sigma=q["sigma"]), {"gamma": Normal(0.0, 1.0), "sigma": LogNormal(0.0, 0.5)},
           particles=16, state_particles=300, proposal=proposals.LinearGaussianObservations(), threshold=0.5, seed=1, max_observations=16)
state = alg.initialize()
for yt in y[:4]:
    state = alg.step(yt, state)
state = alg.rejuvenate(state)
rec = state.engine.export_columns()
state.engine.import_columns(rec, torch.arange(15, -1, -1).cuda())
print("smc2", state.rejuvenations, state.acceptance, rec.shape)
w = torch.randn(7000, 2, device="cuda") * 4
print(pf.resampling.systematic(w.clone()).shape, pf.resampling.multinomial(w.clone()).shape, pf.utils.normalize(w.clone()).sum(0).tolist())
torch.cuda.synchronize()
