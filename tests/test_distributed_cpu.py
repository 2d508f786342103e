"""Host-side logic of the N > 1 path on CPU: gloo backend, world size 2 (columns shard, one all-gather)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyfilter_b200.sharding import LogLikelihoodGather, column_shard, gather_loglikelihood, theta_ess


def test_column_shard_partitions_every_batch():
    for batch in (1, 2, 7, 128, 1024, 1025):
        for world in (1, 2, 3, 8):
            blocks = [column_shard(batch, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == batch
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        column_shard(8, 2, 2)


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(batch, dtype=torch.float32) * -0.5 - 3.0   # stands for the per-column log-likelihoods
        lo, hi = column_shard(batch, rank, world)
        got = gather_loglikelihood(full[lo:hi].clone(), batch)
        ok = torch.equal(got, full) and abs(float(theta_ess(got)) - float(theta_ess(full))) < 1e-6
        gather = LogLikelihoodGather(batch, "cpu")   # the preallocated single-buffer form bench.py uses per move
        for rep in range(2):
            ok = ok and torch.equal(gather(full[lo:hi].clone() + rep), full + rep)
        # timing reduction used by bench.py: max over ranks
        tmax = torch.tensor([float(rank + 1)])
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ok = ok and float(tmax) == float(world)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [7, 1024])
def test_gather_loglikelihood_world2_gloo(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + batch % 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res == {0: True, 1: True}
