"""Stand-in for the un-vendored dependency ``stochproc==0.3.0`` (pyproject.toml:32 of the reference).

TEST INFRASTRUCTURE ONLY.  stochproc is absent from /root/reference and from this image, so this package
re-states the small surface the reference's particle filters touch from its published
behaviour: ``x_t = loc + scale * eps`` for affine processes, ``mean_scale = (x + f*dt, g)`` for the
Euler-Maruyama discretisation and ``N(b + a x, s)`` for linear-Gaussian observations.  It exists so that the
UNMODIFIED reference package can be imported in the build container to pin ``oracle/`` and to generate the
golden vectors under ``tests/golden/``.  Nothing in ``pyfilter_b200`` imports it.
"""
from . import timeseries, distributions  # noqa: F401

__version__ = "0.3.0-standin"
