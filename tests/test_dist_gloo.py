"""The N>1 host path on CPU: two gloo ranks exercise the sharding plan and the max/sum reductions
bench.py uses, and check that per-rank synthetic read ranges concatenate to the single-rank input."""
import os
import socket

import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from oracle import synth
    from uq_b200 import shard
    red = shard.Reducer(dist, "cpu")
    first, n = shard.shard_range(rank, world_size, 37)
    fq = synth.make_fastq(kind="genome", n=n, length=30, seed=9, first=first, genome=500, pool=11)
    secs = 0.5 + rank                       # rank 1 is the slow one
    red.barrier()
    value = shard.job_throughput(red, n, secs)
    parts = [None] * world_size
    dist.all_gather_object(parts, fq)
    if rank == 0:
        out.put((value, b"".join(parts), [shard.split_evenly(10, r, world_size) for r in range(world_size)]))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    import torch.multiprocessing as mp
    from oracle import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    value, joined, even = q.get(timeout=120)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert value == pytest.approx((37 + 37) / 1.5)          # all reads / slowest rank
    whole = synth.make_fastq(kind="genome", n=74, length=30, seed=9, first=0, genome=500, pool=11)
    assert joined == whole                                  # contiguous read ranges, no overlap, no gap
    assert even == [(0, 5), (5, 5)]


def test_single_rank_reducer_is_identity():
    from uq_b200 import shard
    r = shard.Reducer()
    assert r.max(3.5) == 3.5 and r.sum(2) == 2.0
    assert shard.split_evenly(10, 2, 3) == (7, 3) and shard.split_evenly(10, 0, 3) == (0, 4)
