import torch.distributions as _td
from torch.distributions import *  # noqa: F401,F403
from torch.distributions import Distribution, Independent
from . import transforms  # noqa: F401

if not hasattr(Distribution, "to_event"):
    def _to_event(self, n=None):
        return Independent(self, n if n is not None else len(self.batch_shape))

    Distribution.to_event = _to_event
