"""``pyfilter.filters.FilterResult`` (reference filters/result.py:14-164, container.py:10-18): running log-likelihood, deques of
filter means / variances (``record_moments``) and of states (``record_states``)."""
from collections import deque
from copy import deepcopy
from typing import List

import torch


def make_dequeue(maxlen=None) -> deque:
    """``container.py:10-18``: ``False`` -> keep 1, ``True`` -> keep all, an ``int`` -> keep that many."""
    return deque(maxlen=1 if maxlen is False else (None if isinstance(maxlen, bool) else maxlen))


class FilterResult(dict):
    def __init__(self, init_state, record_states=False, record_moments=True):
        super().__init__()
        self._loglikelihood = init_state.get_loglikelihood().clone()
        self._means = make_dequeue(record_moments)
        self._variances = make_dequeue(record_moments)
        self._states = make_dequeue(record_states)
        self.append(init_state)

    @property
    def loglikelihood(self) -> torch.Tensor:
        return self._loglikelihood

    @staticmethod
    def _stack(d):
        return torch.stack(tuple(d), dim=0) if d else torch.tensor([])

    @property
    def filter_means(self) -> torch.Tensor:
        """``{timesteps + 1, [batch shape], latent dimension}`` (initial state included, filters/result.py:40)."""
        return self._stack(self._means)

    @property
    def filter_variance(self) -> torch.Tensor:
        return self._stack(self._variances)

    @property
    def states(self) -> List:
        return list(self._states)

    @property
    def latest_state(self):
        return self._states[-1]

    def append(self, state):
        self._means.append(state.get_mean())
        self._variances.append(state.get_variance())
        self._loglikelihood = self._loglikelihood + state.get_loglikelihood()
        # A state straight from the engine holds zero-copy VIEWS of the device buffers, which the next move overwrites (particles and
        # weights ping-pong, the ancestors are rewritten in place).  A result that keeps more than the latest state therefore takes its
        # copy NOW, while the views still show this state - not when the next one arrives.
        if self._states.maxlen != 1 and getattr(state, "_engine", None) is not None:
            state = state.detach_copy()
        self._states.append(state)
        return self

    def extend_moments(self, means: torch.Tensor, variances: torch.Tensor):
        """Bulk append of the rows the device loop recorded (one row per move)."""
        self._means.extend(means.unbind(0))
        self._variances.extend(variances.unbind(0))

    def exchange(self, other: "FilterResult", mask: torch.Tensor):
        self._loglikelihood[mask] = other.loglikelihood[mask]
        for mine, theirs in ((self._means, other._means), (self._variances, other._variances)):
            for old, new in zip(mine, theirs):
                old[mask] = new[mask]
        for ns, os_ in zip(other.states, self.states):
            os_.exchange(ns, mask)
        return self

    def resample(self, indices: torch.Tensor, entire_history=True):
        self._loglikelihood.copy_(self._loglikelihood[indices])
        if entire_history:
            # new tensors, not the reference's in-place `tens.copy_(tens[indices])` (filters/result.py:110-112): the latest state
            # holds the same tensor objects as the deques and permutes them itself below - in place they would be permuted twice
            for d in (self._means, self._variances):
                for i in range(len(d)):
                    d[i] = d[i][indices]
        for s in self.states:
            s.resample(indices)
        return self

    # The reference's serialisation (filters/result.py:134-156 on top of container.py:113-147): the moment deques travel as stacked
    # tensors under "tensor_tuples" with their maxlen encoded in the key, then the latest state and the running log-likelihood.
    _TT_KEY = "tensor_deque_{maxlen}__{name}"

    def state_dict(self):
        from collections import OrderedDict

        tt = OrderedDict()
        tt[self._TT_KEY.format(maxlen=self._means.maxlen, name="filter_means")] = self.filter_means
        tt[self._TT_KEY.format(maxlen=self._variances.maxlen, name="filter_variances")] = self.filter_variance
        return OrderedDict([("tensor_tuples", tt), ("state", self.latest_state.state_dict()), ("log_likelihood", self.loglikelihood)])

    def load_state_dict(self, sd):
        if "tensor_tuples" in sd:
            for key, v in sd["tensor_tuples"].items():
                type_, name = key.replace("tensor_", "", 1).split("__", 1)
                maxlen = type_.split("_", 1)[1]
                dq = deque(v.unbind(0) if v.numel() else (), maxlen=None if maxlen == "None" else int(maxlen))
                if name == "filter_means":
                    self._means = dq
                elif name == "filter_variances":
                    self._variances = dq
        else:  # flat layout written by earlier versions of this package
            self._means = deque(sd["filter_means"].unbind(0), maxlen=self._means.maxlen)
            self._variances = deque(sd["filter_variances"].unbind(0), maxlen=self._variances.maxlen)
        self._loglikelihood = sd["log_likelihood"]
        assert len(self.states) == 1, "Can only handle case when we have 1 state!"
        self.latest_state.load_state_dict(sd["state"])

    def copy(self) -> "FilterResult":
        new = FilterResult.__new__(FilterResult)
        dict.__init__(new)
        new._loglikelihood = self._loglikelihood.clone()
        new._means = deque((m.clone() for m in self._means), maxlen=self._means.maxlen)
        new._variances = deque((v.clone() for v in self._variances), maxlen=self._variances.maxlen)
        new._states = deque((s.detach_copy() for s in self._states), maxlen=self._states.maxlen)
        return new

    def __repr__(self):
        return f"FilterResult(ll: {self._loglikelihood.__repr__()}, num_observations: {len(self._means)})"
