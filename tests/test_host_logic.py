"""Host-side logic that needs no GPU: the observation schedule of ``observe_every_step`` (reference filters/base.py:204-210),
constructor argument checks, the model factory."""
import math

import pytest
import torch

import pyfilter_b200 as pf
from pyfilter_b200 import timeseries as ts
from pyfilter_b200.filters.particle import APF, SISR, proposals


def test_observation_schedule_matches_reference_loop():
    """Reference: ``while time_index % observe_every_step != 0: propagate`` before every observation (time index of the state)."""
    for k in (1, 2, 3, 5):
        f = APF(ts.build("sine_em", observe_every_step=k), 100)
        for t0 in (0, 1, 4):
            y = torch.arange(1.0, 6.0)
            moves, observed = f._expand_observations(y, t0)
            # replay the reference's loop
            t, exp_rows, exp_obs = t0, [], []
            for v in y.tolist():
                while t % k != 0:
                    exp_rows.append(float("nan")); t += 1
                exp_obs.append(len(exp_rows)); exp_rows.append(v); t += 1
            assert observed == exp_obs
            got = moves.reshape(-1).tolist()
            assert len(got) == len(exp_rows)
            assert all((math.isnan(a) and math.isnan(b)) or a == b for a, b in zip(got, exp_rows))


def test_constructor_checks_and_factory():
    with pytest.raises(ValueError):
        SISR("not a model", 10)
    with pytest.raises(NotImplementedError):
        SISR(ts.build("sv_ar1"), 10, nan_strategy="drop")
    with pytest.raises(NotImplementedError):
        ts.build("garch")
    with pytest.raises(ValueError):
        ts.build("sv_ar1", observe_every_step=0)
    f = SISR(ts.build("lg_ar1"), 1000, ess_threshold=0.5)
    assert f.particles == torch.Size([1000])
    f.set_batch_shape(torch.Size([7]))
    assert f.particles == torch.Size([1000, 7])
    with pytest.raises(NotImplementedError):
        f.set_batch_shape(torch.Size([2, 3]))
    g = f.copy()  # the reference forwards the already scaled threshold (filters/particle/base.py:165)
    assert g._resample_threshold == 0.5 * 1000 * 1000
    f.increase_particles(2)
    assert f.particles == torch.Size([2000, 7])
    with pytest.raises(ValueError):  # same pairing rule as proposals/linear.py:32-36
        proposals.LinearGaussianObservations().set_model(ts.build("sv_ar1"))


def test_filter_signatures_match_the_reference():
    """Same constructor keywords / defaults and method signatures as pyfilter's particle filters (filters/particle/base.py:19-27,
    filters/base.py:22-29), so user code switches by changing the import.  Needs the reference (build container)."""
    import inspect

    from oracle.ref_loader import load_reference, reference_available

    if not reference_available():
        pytest.skip("/root/reference not present")
    load_reference()
    from pyfilter.filters.base import BaseFilter
    from pyfilter.filters.particle import SISR as RefSISR
    from pyfilter.filters.particle.base import ParticleFilter

    mine = inspect.signature(SISR.__init__).parameters
    for ref_sig in (inspect.signature(ParticleFilter.__init__), inspect.signature(BaseFilter.__init__)):
        for name, p in ref_sig.parameters.items():
            if name in ("self", "kwargs"):
                continue
            assert name in mine, name
            if p.default is not inspect.Parameter.empty and name != "resampling":
                assert mine[name].default == p.default, name
    assert list(mine)[:6] == ["self", "model", "particles", "resampling", "proposal", "ess_threshold"]   # positional order
    for meth in ("batch_filter", "filter", "initialize", "initialize_with_result", "copy", "increase_particles", "set_batch_shape",
                 "initialize_model", "smooth", "predict", "correct"):
        a = [n for n in inspect.signature(getattr(SISR, meth)).parameters]
        b = [n for n in inspect.signature(getattr(RefSISR, meth)).parameters]
        assert a == b, (meth, a, b)
