"""Global (one container) encode over 2 GPUs against the single-GPU encode of the same file - needs 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_multigpu_gpu.py -m gpu`); skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    ("genome", 24000, 60, dict(sort="DNA")),
    ("genome", 24000, 60, dict(sort="None", raw=["DNA", "QUAL", "QNAME"])),
    ("genome", 9000, 40, dict(sort="QUAL", raw=["DNA"])),
    ("casava", 26000, 50, dict(sort="QNAME")),
    ("casava", 8000, 50, dict(sort="DNA", raw=["QNAME", "QUAL"])),
    ("illumina", 7000, 30, dict()),
    ("genome", 9000, 40, dict(sort="DNA", pattern=["2.2", "1.1"])),
    ("genome", 9000, 40, dict(sort="None", raw=["DNA", "QUAL"], pattern=["3.1", "1.2"])),
]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import synth
    from uq_b200 import host, multigpu as mg
    from uq_b200.device import Context
    ts = torch.cuda.Stream(device=rank)                  # stream-ordered exchanges, as bench.py runs them
    ctx = Context(rank, stream=ts.cuda_stream)
    comm = mg.Comm(dist, "cuda:%d" % rank, stream=ts)
    results = []
    # every case through the partition-first sample sort, and the first three again through the merge variant
    runs = [(c, "0") for c in CASES] + [(c, "1") for c in CASES[:3]]
    for (kind, n, length, kw), merge in runs:
        os.environ["UQB_MG_MERGE"] = merge
        kwg = dict(genome=max(4 * length, n // 6), pool=max(1, n // 7)) if kind == "genome" else {}
        # uneven contiguous ranges; rank 0 is the largest (it must hold the first 10001 reads)
        wts = [0.55, 0.45] if world == 2 else [0.46] + [0.54 / (world - 1)] * (world - 1)
        cuts = [0] + [int(n * sum(wts[:k + 1])) for k in range(world - 1)] + [n]
        first, cnt = cuts[rank], cuts[rank + 1] - cuts[rank]
        shard = synth.make_fastq(kind=kind, n=cnt, length=length, seed=21, first=first, **kwg)
        fq = ctx.load_fastq(shard)
        res, cfg = mg.encode_sharded(ctx, comm, fq, **kw)
        members = mg.assemble(comm, res)
        res.free(); fq.free()
        if rank == 0:
            whole = synth.make_fastq(kind=kind, n=n, length=length, seed=21, first=0, **kwg)
            want, want_cfg = host.encode(whole, ctx=ctx, **kw)
            bad = []
            if sorted(members) != sorted(want):
                bad.append("names %s vs %s" % (sorted(members), sorted(want)))
            else:
                for k in want:
                    a, b = members[k], np.asarray(want[k])
                    if a.dtype != b.dtype or a.shape != b.shape or not np.array_equal(a, b):
                        bad.append("%s dtype %s/%s shape %s/%s equal %s" % (k, a.dtype, b.dtype, a.shape, b.shape,
                                                                           a.shape == b.shape and bool(np.array_equal(a, b))))
            import json
            if json.loads(json.dumps(cfg, default=str)) != json.loads(json.dumps(want_cfg, default=str)):
                bad.append("config differs")
            results.append((kind + (" merge" if merge == "1" else ""), kw, bad))
        # sharded decode of the container just written: the ranks' texts concatenate to the single-GPU decode
        members = comm.all_gather_object(members)[0]
        text, a, b = mg.decode_sharded(ctx, comm, members, cfg)
        texts = comm.all_gather_object(bytes(text))
        if rank == 0:
            single = host.decode(members, cfg, ctx=ctx).tobytes()
            results.append((kind + " decode", kw, [] if b"".join(texts) == single else ["sharded decode differs"]))
    os.environ["UQB_MG_MERGE"] = "0"
    # a larger case generated on the device, loaded through the streamed path with the reference line, async downloads
    n_each = 700_000
    if world != 2:
        if rank == 0:
            q.put(results)
        dist.destroy_process_group()
        return
    dev = ctx.synth("genome", n_each, 150, 1002, first=rank * n_each, genome=200_000, pool=300_000)
    data = dev.download().copy()
    dev.free()
    pin = ctx.pinned_empty(data.size)
    pin.array[:] = data
    line1 = comm.all_gather_object(bytes(data[:200]).split(b"\n")[0])[0]
    fq = ctx.load_fastq_streamed(pin, chunk_bytes=8 << 20, ref=line1, rbase=0 if rank == 0 else 1)
    bufs = {}
    def sink(name, nbytes):
        bufs[name] = ctx.pinned_empty(nbytes)
        return bufs[name].array
    res, cfg = mg.encode_sharded(ctx, comm, fq, sort="DNA", sink=sink)
    members = mg.assemble(comm, res)
    res.free(); fq.free()
    if rank == 0:
        whole = ctx.synth("genome", 2 * n_each, 150, 1002, first=0, genome=200_000, pool=300_000)
        wfq = ctx.adopt_fastq(whole)
        wm, wcfg = host.encode_device(ctx, wfq, sort="DNA")
        want = wm.download()
        bad = []
        for k in want:
            a, b = members[k], np.asarray(want[k])
            if a.dtype != b.dtype or a.shape != b.shape or not np.array_equal(a, b):
                bad.append("%s %s/%s %s/%s" % (k, a.dtype, b.dtype, a.shape, b.shape))
        if cfg["reads"] != 2 * n_each or cfg["QNAME_columns"] != wcfg["QNAME_columns"]:
            bad.append("config")
        results.append(("device synth 1.4M streamed", "sort DNA", bad))
    if rank == 0:
        q.put(results)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_multi_gpu_global_encode_equals_single_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    results = q.get(timeout=600)
    for p in procs: p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs)
    for kind, kw, bad in results:
        assert not bad, (kind, kw, bad)
