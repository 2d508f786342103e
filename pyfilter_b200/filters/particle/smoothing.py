"""Backward passes over recorded filter states (reference filters/particle/base.py:105-146)."""
import torch

from ... import _lib
from ..utils import batched_gather


def ffbs(filter_, states, uniforms=None) -> torch.Tensor:
    """Forward filtering - backward sampling as the reference writes it (``_do_sample_ffbs``, filters/particle/base.py:105-128): the
    last state is resampled with the filter's resampler; walking backwards, every smoothed particle draws its predecessor from
    ``Categorical(logits = log w_s + log p(x_{s+1}^i | x_s))`` - one O(N^2) device pass per recorded state
    (``smcb_filter_ffbs_step``).  Non-batched filters, like the reference's working branch.  ``uniforms`` (list of float64 ``(N,)``
    tensors, one per backward step) injects the draws (parity hook)."""
    if filter_.batch_shape:
        raise NotImplementedError("FFBS is implemented for non-batched filters")
    states = list(states)
    e = filter_._get_engine(2)
    lib = e.lib
    last = states[-1]
    idx = filter_._resampler(last.weights.clone())
    res = [batched_gather(last.timeseries_state.value, idx)]
    n = int(last.weights.shape[0])
    d = e.D
    seed = int(torch.randint(0, 2**62, (1,)).item())
    for k, state in enumerate(reversed(states[:-1])):
        x = state.timeseries_state.value.to("cuda", torch.float32).reshape(n, d).contiguous()
        lw = state.weights.to("cuda", torch.float32).contiguous()
        xn = res[-1].to(torch.float32).reshape(n, d).contiguous()
        U = None if uniforms is None else uniforms[k].to("cuda", torch.float64).contiguous()
        out_idx = torch.empty(n, device="cuda", dtype=torch.int64)
        out_x = torch.empty((n, d), device="cuda", dtype=torch.float32)
        _lib.check(lib.smcb_filter_ffbs_step(e.handle, x.data_ptr(), lw.data_ptr(), xn.data_ptr(), U.data_ptr() if U is not None else None,
                                             seed, k, out_idx.data_ptr(), out_x.data_ptr(), _lib.current_stream()))
        res.append(out_x.reshape(last.timeseries_state.value.shape))
    return torch.stack(res[::-1], dim=0)
