from . import proposals  # noqa: F401
from .base import ParticleFilter


class SISR(ParticleFilter):
    """Sequential Importance Sampling Resampling (reference filters/particle/sisr.py:7-56)."""

    algorithm_id = 0


class APF(ParticleFilter):
    """Auxiliary Particle Filter of Pitt and Shephard (reference filters/particle/apf.py:9-46)."""

    algorithm_id = 1
