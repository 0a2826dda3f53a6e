#!/usr/bin/env python3
"""Diagnostics: wall clock of every device-level call (synchronised) over several encode steps."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uq_b200 import host, device
from uq_b200.device import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
sort = sys.argv[2] if len(sys.argv) > 2 else "DNA"
ctx = Context(0)
dev = ctx.synth("genome", n, 150, 1002, genome=10_000_000, pool=max(1, n // 5))
LOG = {}
def wrap(cls, name):
    f = getattr(cls, name)
    def g(self, *a, **k):
        c = self if isinstance(self, Context) else self.ctx
        c.sync(); t0 = time.perf_counter()
        r = f(self, *a, **k)
        c.sync(); LOG.setdefault(name, []).append(round((time.perf_counter() - t0) * 1e3, 2))
        return r
    setattr(cls, name, g)
for nm in ("sort_rows", "gather_rows", "narrow_u32", "columns_to_rows", "rows_to_columns", "layout"):
    wrap(Context, nm)
for nm in ("split", "analyze", "qname_scan", "pack", "qname_encode"):
    wrap(device.Fastq, nm)
wrap(device.DeviceArray, "free")
for s in range(4):
    LOG.clear()
    t0 = time.perf_counter()
    fq = ctx.adopt_fastq(dev)
    m, cfg = host.encode_device(ctx, fq, sort=sort)
    m.free(); fq.free(); ctx.sync()
    tot = (time.perf_counter() - t0) * 1e3
    fr = LOG.pop("free", [])
    print("step", s, "total %.1f ms" % tot, "mem", [x >> 20 for x in ctx.mem_info()], "frees %.1f ms (%d)" % (sum(fr), len(fr)))
    print("   ", json.dumps(LOG))
