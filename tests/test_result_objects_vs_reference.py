"""Drop-in boundary, host objects: the theta-level operations SMC2 / NESS / PMMH apply to a filter's result -
``FilterResult.resample / exchange / copy / state_dict`` (reference filters/result.py:76-164) and the same on the state
(filters/particle/state.py:150-208) - give the same tensors as the reference's own classes on the same inputs.  CPU only; needs the
reference (build container), the oracle supplies the moments the device would have computed."""
import pytest
import torch

from oracle import smc_oracle as O
from oracle.ref_loader import reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")

N, B, T = 40, 6, 4


def _inputs(seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for t in range(T + 1):
        x = torch.randn(N, B, generator=g)
        w = torch.randn(N, B, generator=g)
        ll = torch.randn(B, generator=g) if t else torch.zeros(B)   # initialize() starts from zero (particle/base.py:100); the
        # reference aliases its running total to the initial state's tensor, which only a non-zero start would expose
        inds = torch.randint(0, N, (N, B), generator=g)
        out.append((t, x, w, ll, inds))
    return out


def _build(which, seed, record_states=True):
    if which == "ref":
        from oracle.ref_loader import load_reference

        load_reference()
        from pyfilter.filters.particle.state import ParticleFilterCorrection
        from pyfilter.filters.result import FilterResult
        from stochproc.timeseries import TimeseriesState

        def mk(t, x, w, ll, inds):
            return ParticleFilterCorrection(TimeseriesState(t, x.clone(), torch.Size([])), w.clone(), ll.clone(), inds.clone())
    else:
        from pyfilter_b200.filters.particle.state import ParticleFilterCorrection
        from pyfilter_b200.filters.result import FilterResult
        from pyfilter_b200.timeseries import TimeseriesState

        def mk(t, x, w, ll, inds):
            mean, var = O.filter_mean_and_variance(x, O.normalize(w.clone()), 0)
            return ParticleFilterCorrection(TimeseriesState(torch.tensor(t), x.clone(), torch.Size([])), w.clone(), ll.clone(),
                                            inds.clone(), mean.clone(), var.clone())
    states = [mk(*s) for s in _inputs(seed)]
    res = FilterResult(states[0], record_states, True)
    for s in states[1:]:
        res.append(s)
    return res


def _same(a, b, state_moments=True):
    assert torch.allclose(a.loglikelihood, b.loglikelihood, atol=1e-6)
    assert torch.allclose(a.filter_means.reshape(b.filter_means.shape), b.filter_means, atol=1e-6)
    assert torch.allclose(a.filter_variance.reshape(b.filter_variance.shape), b.filter_variance, atol=1e-6)
    assert len(a.states) == len(b.states)
    for k, (sa, sb) in enumerate(zip(a.states, b.states)):
        assert torch.equal(sa.timeseries_state.value, sb.timeseries_state.value)
        assert torch.equal(sa.weights, sb.weights)
        assert torch.equal(sa.previous_indices, sb.previous_indices)
        # The reference's running total IS the initial state's likelihood tensor (filters/result.py:33 takes it without a copy and
        # :131 adds in place), so its recorded state 0 reports the total; pyfilter_b200 keeps the state's own value (0).
        if k > 0:
            assert torch.allclose(sa.get_loglikelihood(), sb.get_loglikelihood())
        if state_moments:
            assert torch.allclose(sa.get_mean().reshape(sb.get_mean().shape), sb.get_mean(), atol=1e-6)
            assert torch.allclose(sa.get_variance().reshape(sb.get_variance().shape), sb.get_variance(), atol=1e-6)
        else:  # what the moments of the permuted columns are, whatever the reference reports (see the caller)
            W = O.normalize(sa.weights.clone())
            m, v = O.filter_mean_and_variance(sa.timeseries_state.value, W, 0)
            assert torch.allclose(sa.get_mean().reshape(m.shape), m, atol=1e-5)
            assert torch.allclose(sa.get_variance().reshape(v.shape), v, atol=1e-5)


def test_append_resample_exchange_copy_match_reference():
    mine, ref = _build("mine", 1), _build("ref", 1)
    _same(mine, ref)
    idx = torch.tensor([3, 3, 0, 5, 1, 1])           # theta-level resampling (inference/sequential/kernels/base.py:15-23)
    mine.resample(idx)
    ref.resample(idx)
    # Everything the callers read (likelihoods, moment history, particles, weights, ancestors) matches.  The per-state `_mean` /
    # `_var` do not: in the reference the deque of recorded moments holds the SAME tensor objects as the states, so
    # FilterResult.resample permutes them in place (filters/result.py:110-112) and the state permutes them again
    # (particle/state.py:157-158) - its states end up with the moments of the wrong columns.  pyfilter_b200 permutes once;
    # its values are checked against the moments recomputed from the permuted particles.
    _same(mine, ref, state_moments=False)
    other_m, other_r = _build("mine", 2), _build("ref", 2)
    mask = torch.tensor([True, False, False, True, True, False])   # accepted proposals (kernels/mh.py:54-58)
    mine.exchange(other_m, mask)
    ref.exchange(other_r, mask)
    _same(mine, ref, state_moments=False)
    cm, cr = mine.copy(), ref.copy()
    _same(cm, cr, state_moments=False)
    before = mine.filter_means.clone()
    cm.resample(torch.tensor([0, 0, 0, 0, 0, 0]))    # a copy does not alias its source
    assert torch.equal(mine.filter_means, before)
    _same(mine, ref, state_moments=False)


def _single(which, seed):
    """A result that keeps only the latest state (the default ``record_states=False``), as ``state_dict`` requires."""
    return _build(which, seed, record_states=False)


def test_state_dict_has_the_reference_layout_and_round_trips():
    """Same keys and tensors as the reference's ``state_dict`` (filters/result.py:134-146, container.py:113-123,
    particle/state.py:176-185), exact round trip (the reference's own test: tests/filters/test_particle.py:153-168), and a
    dictionary written by the reference loads into a pyfilter_b200 result."""
    mine, ref = _single("mine", 3), _single("ref", 3)
    sm, sr = mine.state_dict(), ref.state_dict()
    assert list(sm.keys()) == list(sr.keys())
    assert list(sm["tensor_tuples"].keys()) == list(sr["tensor_tuples"].keys())
    assert set(sm["state"].keys()) == set(sr["state"].keys())
    for k in sr["tensor_tuples"]:
        assert torch.allclose(sm["tensor_tuples"][k].reshape(sr["tensor_tuples"][k].shape), sr["tensor_tuples"][k], atol=1e-6), k
    assert torch.allclose(sm["log_likelihood"], sr["log_likelihood"])
    fresh = _single("mine", 4)
    fresh.load_state_dict(mine.state_dict())
    assert torch.equal(fresh.filter_means, mine.filter_means) and torch.equal(fresh.filter_variance, mine.filter_variance)
    assert torch.equal(fresh.loglikelihood, mine.loglikelihood)
    assert torch.equal(fresh.latest_state.timeseries_state.value, mine.latest_state.timeseries_state.value)
    assert torch.equal(fresh.latest_state.previous_indices, mine.latest_state.previous_indices)
    cross = _single("mine", 5)
    cross.load_state_dict(ref.state_dict())
    assert torch.allclose(cross.filter_means.reshape(ref.filter_means.shape), ref.filter_means)
    assert torch.equal(cross.latest_state.weights, ref.latest_state.weights)
    assert torch.equal(cross.latest_state.timeseries_state.value, ref.latest_state.timeseries_state.value)
