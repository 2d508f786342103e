"""The reference's own particle-filter test class (tests/filters/test_particle.py: ``TestParticleFilters``) restated for this repository's
classes on the model of its ``linear_models()`` that the compiled zoo holds (1-D AR(1) observed with noise, tests/filters/models.py:12-26):
same parametrisation - filters x proposals x batch shapes x 10 % missing observations x ``copy`` -, same assertions, same 10 %
median-relative-deviation criterion.  pykalman is absent, so the comparator is the closed-form scalar Kalman filter / RTS smoother
(NaN observations skipped), and the data are simulated from the model itself."""
from functools import partial

import numpy as np
import pytest
import torch

from oracle import smc_oracle as O

pytestmark = pytest.mark.gpu
RELATIVE_TOLERANCE = 1e-1
SERIES_LENGTH = 100
BATCH_SIZES = [torch.Size([]), torch.Size([3])]
MISSING_PERC = [0.0, 0.1]
P = O.DEFAULT_PARAMS["lg_ar1"]


def median_relative_deviation(y_true, y):
    return np.median(np.abs((y_true - y) / y_true))


def _filters(particles=1_500, **kwargs):
    from pyfilter_b200.filters import particle as part

    out = []
    for pt in (part.APF, part.SISR):
        out.append(("%s-bootstrap" % pt.__name__, partial(pt, particles=particles, proposal=part.proposals.Bootstrap(), **kwargs)))
        for second in (False, True):
            prop = part.proposals.Linearized(n_steps=5, use_second_order=second)
            out.append(("%s-linearized%d" % (pt.__name__, second), partial(pt, particles=particles, proposal=prop, **kwargs)))
        out.append(("%s-lgo" % pt.__name__, partial(pt, particles=particles, proposal=part.proposals.LinearGaussianObservations(), **kwargs)))
        out.append(("%s-nested50" % pt.__name__, partial(pt, particles=particles, proposal=part.proposals.NestedProposal(50), **kwargs)))
    out.append(("GPF-gaussian", partial(part.GPF, particles=particles, **kwargs)))   # tests/filters/test_particle.py:22-23
    return out


FILTERS = _filters()
SMOOTH_FILTERS = _filters(particles=1_500, record_states=True)[:-1]   # skip_gpf=True (test_particle.py:172)


def _data(missing_perc):
    np.random.seed(123)
    torch.manual_seed(123)
    _, y = O.build_model("lg_ar1").simulate(SERIES_LENGTH)
    y = y.float()
    if missing_perc:
        idx = np.random.randint(1, SERIES_LENGTH, size=int(missing_perc * SERIES_LENGTH))
        y[idx] = float("nan")
    return y


def _kalman(y, smooth=False):
    """Scalar Kalman filter (and RTS smoother) of the AR(1) model; NaN observations are skipped (what pykalman does with masked values)."""
    a_, b_, q, h, r = P["alpha"], P["beta"], P["sigma"] ** 2, P["a"], P["s"] ** 2
    m, p = P["alpha"], P["sigma"] ** 2 / (1 - P["beta"] ** 2)
    mp, pp, mf, pf_, ll = [], [], [], [], 0.0
    for yt in y.numpy():
        m_pred, p_pred = a_ + b_ * m, b_ * b_ * p + q
        m, p = m_pred, p_pred
        if not np.isnan(yt):
            s = h * h * p + r
            resid = yt - (P["b"] + h * m)
            ll += -0.5 * (np.log(2 * np.pi * s) + resid * resid / s)
            k = p * h / s
            m, p = m + k * resid, (1 - k * h) * p
        mp.append(m_pred); pp.append(p_pred); mf.append(m); pf_.append(p)
    mf, pf_, mp, pp = map(np.array, (mf, pf_, mp, pp))
    if not smooth:
        return mf, ll
    ms = mf.copy()
    for t in range(len(mf) - 2, -1, -1):
        g = pf_[t] * b_ / pp[t + 1]
        ms[t] = mf[t] + g * (ms[t + 1] - mp[t + 1])
    return ms, ll


def _model():
    from pyfilter_b200 import timeseries as ts

    return ts.build("lg_ar1")


@pytest.mark.parametrize("name,filter_", FILTERS, ids=[n for n, _ in FILTERS])
@pytest.mark.parametrize("batch_size", BATCH_SIZES, ids=["nobatch", "batch3"])
@pytest.mark.parametrize("missing_perc", MISSING_PERC)
@pytest.mark.parametrize("test_copy", [False, True])
def test_filter_and_log_likelihood(name, filter_, batch_size, missing_perc, test_copy):
    """tests/filters/test_particle.py:67-111."""
    y = _data(missing_perc)
    kalman_mean, kalman_ll = _kalman(y)
    kalman_mean = kalman_mean[:, None, None] if len(batch_size) > 0 else kalman_mean[:, None]
    f = filter_(_model())
    f.set_batch_shape(batch_size)
    result = f.batch_filter(y, bar=False)
    if test_copy:
        old_result = result
        result = result.copy()
        assert result is not old_result
        for new_state, copy_state in zip(result.states, old_result.states):
            assert new_state is not copy_state
            assert (new_state.timeseries_state.value == copy_state.timeseries_state.value).all()
            assert (new_state.normalized_weights() == copy_state.normalized_weights()).all()
    assert len(result.states) == 1
    assert (median_relative_deviation(kalman_ll, result.loglikelihood.cpu().numpy()) < RELATIVE_TOLERANCE).all()
    means = result.filter_means[1:]
    assert means.shape == torch.Size([SERIES_LENGTH, *batch_size, 1])
    assert median_relative_deviation(kalman_mean, means.cpu().numpy()) < RELATIVE_TOLERANCE


@pytest.mark.parametrize("name,filter_", FILTERS[::2], ids=[n for n, _ in FILTERS[::2]])
@pytest.mark.parametrize("batch_size", BATCH_SIZES, ids=["nobatch", "batch3"])
@pytest.mark.parametrize("missing_perc", MISSING_PERC)
def test_predict(name, filter_, batch_size, missing_perc):
    """tests/filters/test_particle.py:113-135."""
    y = _data(missing_perc)
    model = _model()
    f = filter_(model)
    f.set_batch_shape(batch_size)
    result = f.batch_filter(y, bar=False)
    num_steps = 10
    path = result.latest_state.predict_path(model, num_steps)
    assert len(path.get_paths()) == 2
    x, yy = path.get_paths()
    assert x.shape == torch.Size([num_steps, *f.particles, *f.ssm.hidden.event_shape])


@pytest.mark.parametrize("name,filter_", FILTERS[::2], ids=[n for n, _ in FILTERS[::2]])
@pytest.mark.parametrize("batch_size", BATCH_SIZES, ids=["nobatch", "batch3"])
@pytest.mark.parametrize("missing_perc", MISSING_PERC)
def test_save_and_load(name, filter_, batch_size, missing_perc):
    """tests/filters/test_particle.py:137-168."""
    y = _data(missing_perc)
    f = filter_(_model())
    f.set_batch_shape(batch_size)
    result = f.batch_filter(y, bar=False)
    state_dict = result.state_dict()
    new_result = f.initialize_with_result()
    new_result.load_state_dict(state_dict)
    assert ((new_result.filter_means == result.filter_means).all() and (new_result.filter_variance == result.filter_variance).all()
            and (new_result.loglikelihood == result.loglikelihood).all())
    for new_s, old_s in zip(new_result.states, result.states):
        new_ts, old_ts = new_s.get_timeseries_state(), old_s.get_timeseries_state()
        assert (new_ts.value == old_ts.value).all() and (new_ts.time_index == old_ts.time_index).all()


@pytest.mark.parametrize("name,filter_", SMOOTH_FILTERS[::3], ids=[n for n, _ in SMOOTH_FILTERS[::3]])
@pytest.mark.parametrize("batch_size", BATCH_SIZES, ids=["nobatch", "batch3"])
@pytest.mark.parametrize("missing_perc", MISSING_PERC)
@pytest.mark.parametrize("method", ["ffbs", "fl"])
def test_smooth(name, filter_, batch_size, missing_perc, method):
    """tests/filters/test_particle.py:171-208 (FFBS for non-batched filters: the reference's batched branch does not run either)."""
    if method == "ffbs" and len(batch_size) > 0:
        pytest.skip("FFBS is implemented for non-batched filters")
    y = _data(missing_perc)
    kalman_mean, _ = _kalman(y, smooth=True)
    kalman_mean = kalman_mean[:, None, None] if len(batch_size) > 0 else kalman_mean[:, None]
    f = filter_(_model())
    f.set_batch_shape(batch_size)
    result = f.batch_filter(y, bar=False)
    assert len(result.states) == kalman_mean.shape[0] + 1
    smoothed = f.smooth(result.states, method=method)
    means = smoothed[1:].mean(1).unsqueeze(-1).cpu().numpy()
    n90 = int(0.9 * SERIES_LENGTH)
    if method != "fl":
        assert median_relative_deviation(kalman_mean[-n90:], means[-n90:]) < RELATIVE_TOLERANCE
    else:
        assert median_relative_deviation(kalman_mean[-10:], means[-10:]) < RELATIVE_TOLERANCE
