"""``pyfilter.resampling`` on the device (reference resampling.py:8-105).

``systematic`` returns ancestors that are bit-identical to the reference's CPU path
(``torch.cumsum`` -> ``torch.searchsorted``) for the same normalised weights and offsets ``u``; ``multinomial`` is
bit-identical to ``torch.multinomial``'s CPU kernel for the same float64 uniforms.  Both accept ``(N,)`` or ``(N, B)`` input
(particles along dim 0) and return an int64 tensor of the same shape - for 2-D input a transposed view with strides
``(1, N)``, exactly what the reference's ``moveaxis`` wrapper produces (resampling.py:16-19).
"""
from typing import Optional, Union

import torch

from . import _lib
from .utils import _check_weights, _strides_2d

_seed_counter = [0]


def _next_seed() -> int:
    """Philox key for a call without injected randomness: drawn from torch's CPU generator so ``torch.manual_seed`` governs it."""
    return int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())


def _prep(w: torch.Tensor):
    w = _check_weights(w)
    n, b, sn, sb = _strides_2d(w)
    out = torch.empty((b, n), device=w.device, dtype=torch.int64)  # rows contiguous; returned transposed like the reference
    return w, n, b, sn, sb, out


def systematic(w: torch.Tensor, normalized: bool = False, u: Optional[Union[torch.Tensor, float]] = None) -> torch.Tensor:
    """Performs systematic resampling on either a 1D or 2D array.

    Args:
        w: log weights (or normalised weights with ``normalized=True``) of shape ``(N,)`` or ``(N, B)``.
        normalized: whether the weights are normalized or not.
        u: overrides the sampled offsets, shape ``(B, 1)`` (testing hook of the reference, resampling.py:25,32).  Unlike the
            reference - whose wrapper drops the keyword for 1-D input - it is honoured for 1-D too.
    """
    w, n, b, sn, sb, out = _prep(w)
    lib = _lib.load_library()
    u_ptr, u_t = None, None
    if u is not None:
        u_t = torch.as_tensor(u, dtype=torch.float32, device=w.device).reshape(-1)
        if u_t.numel() == 1 and b > 1:
            u_t = u_t.expand(b)
        u_t = u_t.contiguous()
        if u_t.numel() != b:
            raise ValueError("`u` must hold one offset per column")
        u_ptr = u_t.data_ptr()
    seed = 0 if u is not None else _next_seed()
    _lib.check(lib.smcb_systematic(w.data_ptr(), n, b, sn, sb, int(bool(normalized)), u_ptr, seed, out.data_ptr(), 1, n,
                                   _lib.current_stream()))
    if not normalized:
        w.nan_to_num_(-float("inf"), posinf=-float("inf"))  # reference normalize() mutates the caller's tensor
    return out[0] if w.dim() == 1 else out.moveaxis(0, 1)


def multinomial(w: torch.Tensor, normalized: bool = False, U: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Performs multinomial sampling (``torch.multinomial(W, N, replacement=True)`` semantics).

    ``U`` (float64, ``(B, N)`` in draw order) injects the uniforms; otherwise they come from Philox."""
    w, n, b, sn, sb, out = _prep(w)
    lib = _lib.load_library()
    U_ptr = None
    if U is not None:
        U = torch.as_tensor(U, dtype=torch.float64, device=w.device).reshape(b, n).contiguous()
        U_ptr = U.data_ptr()
    seed = 0 if U is not None else _next_seed()
    _lib.check(lib.smcb_multinomial(w.data_ptr(), n, b, sn, sb, int(bool(normalized)), U_ptr, seed, out.data_ptr(), 1, n,
                                    _lib.current_stream()))
    if not normalized:
        w.nan_to_num_(-float("inf"), posinf=-float("inf"))
    return out[0] if w.dim() == 1 else out.moveaxis(0, 1)


def residual(w: torch.Tensor, normalized: bool = False, U: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Performs residual resampling (reference resampling.py:68-105): ``floor(N W_j)`` copies of particle ``j`` in particle order,
    then multinomial draws on the fractional parts.  One column only, like the reference.  ``U`` (float64, draw order) injects the
    uniforms of the multinomial part; otherwise they come from Philox."""
    if w.dim() > 1:
        raise NotImplementedError("Not implemented for multidimensional arrays!")  # resampling.py:78-79
    from .utils import normalize

    W = _check_weights(w) if normalized else normalize(w)
    n = W.shape[0]
    out = torch.empty(n, device=W.device, dtype=torch.int64)
    lib = _lib.load_library()
    U_ptr = None
    if U is not None:
        U_full = torch.zeros(n, dtype=torch.float64, device=W.device)
        U = torch.as_tensor(U, dtype=torch.float64, device=W.device).reshape(-1)
        U_full[: U.numel()] = U
        U_ptr = U_full.data_ptr()
    seed = 0 if U is not None else _next_seed()
    _lib.check(lib.smcb_residual(W.data_ptr(), n, 1, W.stride(0), 0, U_ptr, seed, out.data_ptr(), 1, n, _lib.current_stream()))
    return out
