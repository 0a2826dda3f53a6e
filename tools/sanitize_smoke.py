#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer: every kernel family on tiny inputs."""
import io, lzma, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from conftest import golden_case
from oracle import uq_literal as lit
from uq_b200 import host
from uq_b200.device import Context

ctx = Context(0)
for name in ("c2_keyed_sortDNA", "c5_variable_sortQUAL", "c3_casava_sortQNAME", "c4_pad1_notricks1", "c1_raw_p22_11", "c7_offset_suffix_keyed"):
    fq, uq, kw = golden_case(name)
    want, _ = lit.read_container(uq)
    got, cfg = host.encode(fq, ctx=ctx, **kw)
    for k in want:
        assert np.array_equal(got[k], want[k]), (name, k)
    host.decode(got, cfg, ctx=ctx)
    pin = ctx.pinned_empty(len(fq)); pin.array[:] = np.frombuffer(fq, dtype=np.uint8)
    h = ctx.load_fastq_streamed(pin, chunk_bytes=16384)
    m, c = host.encode_device(ctx, h, **kw)
    m.download(); m.free(); h.free(); pin.free()
rng = np.random.default_rng(1)
for n, w in ((5000, 17), (3000, 113), (70000, 9)):
    t = rng.integers(0, 2, size=(n, w), dtype=np.uint8)
    t[: n // 3, : w - 1] = t[0, : w - 1]
    d = ctx.upload(t)
    p, k, u, nu = ctx.sort_rows(d, True, True, True)
    v = np.ascontiguousarray(t).view("V%d" % w).reshape(-1)
    assert np.array_equal(p.download(dtype=np.uint32).reshape(-1), np.argsort(v, kind="stable").astype(np.uint32))
dev = ctx.synth("illumina", 3000, 100, 5)
fq = ctx.adopt_fastq(dev)
m, c = host.encode_device(ctx, fq, sort="QUAL", pattern=["3.2", "1.1"], raw=["QUAL"])
m.download(); m.free()
print("sanitize smoke ok")
