"""Stand-in for ``pyro-ppl`` (pyproject.toml:30 of the reference): re-export of ``torch.distributions`` plus
``Distribution.to_event``.  TEST INFRASTRUCTURE ONLY (see oracle/standins/stochproc/__init__.py)."""
from . import distributions  # noqa: F401


def factor(name, value):
    raise NotImplementedError("pyro.factor is not part of the SMC hot path")
