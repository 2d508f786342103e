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
