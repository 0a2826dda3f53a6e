#!/usr/bin/env python3
"""Diagnostics: device decode throughput (tables already in HBM -> FASTQ text in HBM)."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from uq_b200 import host, _lib as L
from uq_b200.device import Context, DeviceArray

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
ctx = Context(0)
dev = ctx.synth("genome", n, 150, 1002, genome=10_000_000, pool=max(1, n // 5))
fq = ctx.adopt_fastq(dev)
members, cfg = host.encode_device(ctx, fq, sort="None", raw=["DNA", "QUAL", "QNAME"])
hm = members.download()
members.free(); fq.free()
ctx.timing(True)
for it in range(3):
    ctx.timing_reset(); ctx.sync(); t0 = time.perf_counter()
    out = host.decode(hm, cfg, ctx=ctx)
    dt = time.perf_counter() - t0
    rep = ctx.timing_report()
print("decode incl. upload/download: %.1f ms for %d reads (%.1f M reads/s, %.1f GB/s of FASTQ)" % (dt * 1e3, n, n / dt / 1e6, out.nbytes / dt / 1e9))
assert out.tobytes() == dev.download().tobytes()
for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:8]:
    print("  %-28s %4d launches %8.2f ms" % (k, v[0], v[1]))
