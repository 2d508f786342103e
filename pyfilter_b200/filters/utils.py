"""``pyfilter.filters.utils`` (reference filters/utils.py:4-21)."""
import torch

from .. import _lib


def _as_nbd(x: torch.Tensor, idx: torch.Tensor):
    """``x`` (N, [B], [d...]) and ``idx`` (N, [B]) -> contiguous (N, B, D) float32 / (N, B) int64 plus what undoes the reshape."""
    _lib.require_cuda()
    if not x.is_cuda or not idx.is_cuda:
        raise _lib.SmcbError("pyfilter_b200 operators take CUDA tensors; there is no CPU fallback")
    n = x.shape[0]
    nb = idx.dim() - 1
    if nb > 1:
        raise NotImplementedError("Currently do not support nested batches!")
    b = idx.shape[1] if nb else 1
    xc = x.to(torch.float32).reshape(n, b, -1).contiguous()
    return xc, idx.to(torch.int64).reshape(idx.shape[0], b).contiguous(), n, b, xc.shape[2]


def batched_gather(x: torch.Tensor, indices: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """Gathers ``x`` along the particle axis with one index per particle and batch element, broadcast over the event dimensions:
    ``out[i, b, ...] = x[indices[i, b], b, ...]``."""
    if dim != 0:
        raise NotImplementedError("particles live on dim 0")
    xc, ic, n, b, d = _as_nbd(x, indices)
    if ic.shape[0] != n:
        raise NotImplementedError("one index per particle")
    out = torch.empty_like(xc)
    lib = _lib.load_library()
    _lib.check(lib.smcb_batched_gather(xc.data_ptr(), n, b, d, ic.clone().data_ptr(), None, out.data_ptr(), _lib.current_stream()))
    return out.reshape(x.shape)


def trace_back(x: torch.Tensor, lineage: torch.Tensor, previous_indices: torch.Tensor):
    """One backward step of ancestral tracing (filters/particle/base.py:141-144): ``lineage <- previous_indices[lineage]`` followed by
    the gather of ``x`` along the new lineage.  Returns ``(x[lineage], lineage)``."""
    xc, lc, n, b, d = _as_nbd(x, lineage)
    pc = previous_indices.to(torch.int64).reshape(n, b).contiguous()
    lc = lc.clone()
    out = torch.empty_like(xc)
    lib = _lib.load_library()
    _lib.check(lib.smcb_batched_gather(xc.data_ptr(), n, b, d, lc.data_ptr(), pc.data_ptr(), out.data_ptr(), _lib.current_stream()))
    return out.reshape(x.shape), lc.reshape(lineage.shape)
