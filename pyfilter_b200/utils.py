"""``pyfilter.utils`` for the hot path: ``normalize`` (utils.py:49-64) and ``get_ess`` (utils.py:8-20) on the device."""
import torch

from . import _lib


def _strides_2d(w: torch.Tensor):
    if w.dim() == 1:
        return w.shape[0], 1, w.stride(0), 0
    if w.dim() == 2:
        return w.shape[0], w.shape[1], w.stride(0), w.stride(1)
    raise NotImplementedError("Currently do not support nested batches!")  # filters/base.py:116-117


def _check_weights(w: torch.Tensor):
    _lib.require_cuda()
    if not w.is_cuda:
        raise _lib.SmcbError("pyfilter_b200 operators take CUDA tensors; there is no CPU fallback")
    if w.dtype != torch.float32:
        w = w.float()
    return w


def normalize(weights: torch.Tensor) -> torch.Tensor:
    """Normalizes a 1D or 2D array of log weights over dim 0 (NaN / +inf count as -inf, a column that sums to zero becomes
    uniform).  Like the reference, the input is sanitised in place (``nan_to_num_``, utils.py:57)."""
    w = _check_weights(weights)
    n, b, sn, sb = _strides_2d(w)
    out = torch.empty_like(w, memory_format=torch.contiguous_format)
    on, ob, osn, osb = _strides_2d(out)
    lib = _lib.load_library()
    _lib.check(lib.smcb_normalize(w.data_ptr(), n, b, sn, sb, out.data_ptr(), osn, osb, None, _lib.current_stream()))
    weights.nan_to_num_(-float("inf"), posinf=-float("inf"))  # the reference mutates its argument (Appendix A-1)
    return out


def get_ess(weights: torch.Tensor, normalized: bool = False) -> torch.Tensor:
    """ESS = 1 / sum_i W_i^2 over dim 0 from an array of (log) weights."""
    w = _check_weights(weights)
    n, b, sn, sb = _strides_2d(w)
    ess = torch.empty(b, device=w.device, dtype=torch.float32)
    lib = _lib.load_library()
    _lib.check(lib.smcb_get_ess(w.data_ptr(), n, b, sn, sb, int(bool(normalized)), ess.data_ptr(), _lib.current_stream()))
    return ess.reshape(w.shape[1:])
