"""The drop-in claim at the level of the reference's own callers: the UNMODIFIED reference package (baseline/_ref, imported with the
stand-ins of oracle/standins for its absent dependencies) runs ITS ``SMC2`` - ``inference/sequential/smc2.py``, the particle
Metropolis-Hastings kernel ``kernels/mh.py`` with ``run_pmmh`` (``batch/mcmc/utils.py``), its ``InferenceContext`` and priors - on top
of THIS repository's ``APF``: ``set_batch_shape``, ``initialize_model(context)`` with a model builder that reads the context's parameters,
``initialize``, ``filter(y, state, result=)``, ``copy``, ``batch_filter``, ``FilterResult.resample / exchange / loglikelihood``.  Skipped
when the reference install is absent."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def _reference():
    if not os.path.isdir(os.path.join(REF, "pyfilter")):
        pytest.skip("baseline/_ref not installed (python __graft_entry__.py in the build container)")
    for p in (os.path.join(ROOT, "oracle", "standins"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pyfilter

    assert os.path.abspath(pyfilter.__file__).startswith(os.path.abspath(REF))
    return pyfilter


def test_reference_smc2_drives_this_repositorys_filter():
    _reference()
    from pyfilter import inference as inf
    from pyfilter.inference.sequential import SMC2
    from pyro.distributions import LogNormal, Normal

    from oracle import smc_oracle as O
    from pyfilter_b200 import timeseries as ts
    from pyfilter_b200.filters.particle import APF, proposals

    torch.manual_seed(2)
    _, y = O.build_model("sine_em", dict(gamma=0.5, sigma=2.0)).simulate(40)
    y = y.float().cuda()

    def build_model(context):   # what a pyfilter user writes (examples/*.ipynb), with this repository's model zoo behind it
        gamma = context.named_parameter("gamma", Normal(0.0, 1.0).cuda())
        sigma = context.named_parameter("sigma", LogNormal(0.0, 0.5).cuda())
        return ts.build("sine_em", gamma=gamma, sigma=sigma)

    with inf.make_context() as context:
        filt = APF(build_model, 256, proposal=proposals.LinearGaussianObservations(), seed=7)
        alg = SMC2(filt, 64, threshold=0.5)            # the reference's algorithm object
        state = alg.fit(y)
        w = state.normalized_weights()
        sigma = context.get_parameter("sigma")
        post = float((w * sigma).sum())
        ess = state.ess
    assert torch.isfinite(state.w).all() and state.filter_state.loglikelihood.shape == (64,)
    assert len(ess) == 41 and float(ess.min()) < 0.5 * 64 <= 64.0 + 1e-3      # the ESS fell below the threshold: the PMH kernel ran
    assert (ess[1:] > ess[:-1] + 10).any()                                      # ... and reset the weights (mh.py:107)
    assert 1.3 < post < 2.9, post                                               # data from sigma = 2 under a LogNormal(0, 0.5) prior (mean 1.13)
    assert state.filter_state.filter_means.shape[0] == 41


def _setup():
    _reference()
    from pyro.distributions import LogNormal, Normal

    from oracle import smc_oracle as O
    from pyfilter_b200 import timeseries as ts

    torch.manual_seed(2)
    _, y = O.build_model("sine_em", dict(gamma=0.5, sigma=2.0)).simulate(40)

    def build_model(context):
        gamma = context.named_parameter("gamma", Normal(0.0, 1.0).cuda())
        sigma = context.named_parameter("sigma", LogNormal(0.0, 0.5).cuda())
        return ts.build("sine_em", gamma=gamma, sigma=sigma)

    return y.float().cuda(), build_model


def test_reference_ness_drives_this_repositorys_filter():
    """The reference's NESS (inference/sequential/ness.py with kernels/online.py and the jittering kernels): every update resamples the
    theta-particles (``FilterResult.resample(indices, entire_history=False)``), jitters the context's parameters and hands the filter its
    new model (``initialize_model``)."""
    y, build_model = _setup()
    from pyfilter import inference as inf
    from pyfilter.inference.sequential import NESS

    from pyfilter_b200.filters.particle import APF, proposals

    with inf.make_context() as context:
        alg = NESS(APF(build_model, 128, proposal=proposals.LinearGaussianObservations(), seed=3), 128, threshold=0.9)
        state = alg.fit(y)
        post = float((state.normalized_weights() * context.get_parameter("sigma")).sum())
        ess = state.ess
    assert torch.isfinite(state.w).all() and len(ess) == 41
    assert (ess[1:] > ess[:-1] + 5).any()            # updates happened (weights reset, online.py:51)
    assert 1.3 < post < 2.9, post


def test_reference_pmmh_drives_this_repositorys_filter():
    """The reference's PMMH (inference/batch/mcmc/pmmh.py): chains = the filter's batch, every iteration re-filters the whole data with
    the candidate parameters (``copy`` + ``initialize_model`` + ``batch_filter``) and exchanges the accepted chains."""
    y, build_model = _setup()
    from pyfilter import inference as inf
    from pyfilter.inference.batch.mcmc import PMMH

    from pyfilter_b200.filters.particle import APF, proposals

    with inf.make_context() as context:
        alg = PMMH(APF(build_model, 256, proposal=proposals.LinearGaussianObservations(), seed=4), 40, num_chains=8)
        state = alg.fit(y)
        sigma = context.get_parameter("sigma")
    assert state.filter_state.loglikelihood.shape == (8,) and torch.isfinite(state.filter_state.loglikelihood).all()
    assert torch.isfinite(sigma).all() and (sigma > 0).all()
