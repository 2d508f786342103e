"""Column sharding of a batch of independent filters across GPUs.

The batch dimension of ``set_batch_shape`` (reference filters/base.py:93-119; the theta-particles of SMC2 / NESS,
inference/sequential/base.py:31-34) is embarrassingly parallel: every rank owns a contiguous block of columns and runs the
single-GPU kernels unchanged.  The only exchange is the vector of marginal log-likelihoods every rank needs for the theta-level
ESS test (inference/sequential/state.py:43-44, smc2.py:59-62): one all-gather of ``(B_local,)`` floats per step.
"""
from typing import Tuple

import torch


def column_shard(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block ``[lo, hi)`` of the ``batch`` columns owned by ``rank`` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_loglikelihood(ll_local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """All-gathers the per-column log-likelihoods of every rank into the full ``(batch,)`` vector, in column order."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [column_shard(batch, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    if ll_local.numel() != hi - lo:
        raise ValueError("local vector does not match this rank's shard")
    padded = torch.zeros(width, dtype=ll_local.dtype, device=ll_local.device)
    padded[: hi - lo] = ll_local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[: h - l] for o, (l, h) in zip(out, sizes)])


class LogLikelihoodGather:
    """The per-move exchange of the theta-sharded loop without per-call allocations: every rank contributes ``width`` floats
    (its shard padded to the widest one) to ONE preallocated ``(world * width,)`` buffer through ``all_gather_into_tensor``
    (a single NCCL kernel, no list of outputs, no concatenation); ``__call__`` returns a view of that buffer in column order when
    the shards are equal, else the compacted copy.  Same result as :func:`gather_loglikelihood`."""

    def __init__(self, batch: int, device, dtype=torch.float32, group=None):
        import torch.distributed as dist

        self.group, self.batch = group, batch
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.sizes = [column_shard(batch, r, self.world) for r in range(self.world)]
        self.width = max(hi - lo for lo, hi in self.sizes)
        self.even = all(hi - lo == self.width for lo, hi in self.sizes)
        self.send = torch.zeros(self.width, dtype=dtype, device=device)
        self.recv = torch.empty(self.world * self.width, dtype=dtype, device=device)

    def __call__(self, ll_local: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist

        lo, hi = self.sizes[self.rank]
        if ll_local.numel() != hi - lo:
            raise ValueError("local vector does not match this rank's shard")
        src = ll_local if (self.even and ll_local.is_contiguous()) else self.send
        if src is self.send:
            self.send[: hi - lo].copy_(ll_local)
        dist.all_gather_into_tensor(self.recv, src, group=self.group)
        if self.even:
            return self.recv
        return torch.cat([self.recv[r * self.width: r * self.width + (h - l)] for r, (l, h) in enumerate(self.sizes)])


class PeerExchange:
    """The per-move exchange of the theta-sharded loop WITHOUT a collective launch: the kernel that finalises a column stores its
    log-likelihood increment and running total (each with a sequence tag, one 8-byte store) straight into every rank's buffer over
    NVLink peer memory (``smcb_filter_attach_exchange``, include/smcb200.h); :meth:`wait` enqueues the small reader that polls this
    rank's buffer until all columns of the batch have arrived and returns ``(increments, totals)``, each ``(batch,)`` in column order.

    The buffers are torch symmetric memory (``torch.distributed._symmetric_memory``: allocation and address exchange are plumbing);
    ``buffers=`` takes ordinary CUDA tensors instead - one per rank, all visible to this process - which is how the single-GPU tests
    emulate several ranks.  Protocol: every rank calls ``wait()`` after every ``run()`` that publishes (two buffers alternate by
    parity of the sequence number, so a rank may be one exchange ahead of a peer, not two)."""

    def __init__(self, engine, batch: int, first_column: int, group=None, buffers=None, rank: int = None):
        import ctypes as C

        from . import _lib

        self.engine, self.batch = engine, int(batch)
        words = 2 * 2 * self.batch   # [parity][increment | total][column], 8 bytes each
        if buffers is None:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem

            group = group if group is not None else dist.group.WORLD
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
            self._buf = symm_mem.empty(words, dtype=torch.int64, device=torch.device("cuda", torch.cuda.current_device()))
            self._buf.zero_()
            self._hdl = symm_mem.rendezvous(self._buf, group)
            ptrs = [int(p) for p in self._hdl.buffer_ptrs]
            torch.cuda.synchronize()
            dist.barrier(group)   # every buffer is zeroed before anybody stores into it
        else:
            self.world, self.rank = len(buffers), int(rank)
            for b in buffers:
                assert b.is_cuda and b.dtype == torch.int64 and b.numel() >= words
            self._buf = buffers
            ptrs = [int(b.data_ptr()) for b in buffers]
        arr = (C.c_uint64 * self.world)(*ptrs)
        _lib.check(engine.lib.smcb_filter_attach_exchange(engine.handle, arr, self.world, self.rank, self.batch, int(first_column)))

    def wait(self):
        import ctypes as C

        from . import _lib

        out = C.c_void_p()
        _lib.check(self.engine.lib.smcb_filter_exchange_wait(self.engine.handle, C.byref(out), _lib.current_stream()))
        t = _lib.as_tensor(out.value, (2, self.batch), "<f4", self.engine)
        return t[0], t[1]

    def close(self):
        from . import _lib

        _lib.check(self.engine.lib.smcb_filter_attach_exchange(self.engine.handle, None, 0, 0, 0, 0))


def theta_ess(ll_total: torch.Tensor) -> torch.Tensor:
    """ESS of the theta-particles from their accumulated log-likelihoods (reference utils.py:8-20 on the theta weights)."""
    w = torch.softmax(ll_total - ll_total.max(), dim=0)
    return 1.0 / (w * w).sum()
