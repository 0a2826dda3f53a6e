#!/usr/bin/env python3
"""Diagnostics: uqb_sort_rows on an adversarial QUAL table - binned qualities (four values), every row shares its first 64
bytes with every other row, duplicated tails with point differences - next to the bench's i.i.d. quality table of the same
size.  Round 0 (first 8 bytes) leaves ONE tie group; the per-group common-prefix rounds of sort.cu have to split it."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uq_b200 import host
from uq_b200.device import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
width = 113
ctx = Context(0)
rng = np.random.default_rng(7)
pool = max(1, n // 5)
tails = rng.integers(0, 4, size=(pool, width - 64), dtype=np.uint8) * 21 + 2        # four binned values
pick = rng.integers(0, pool, size=n)
t = np.empty((n, width), dtype=np.uint8)
t[:, :64] = 65
t[:, 64:] = tails[pick]
mut = rng.random(n) < 0.5                                                           # half of the rows differ from their pool entry in one byte
pos = rng.integers(64, width, size=n)
t[np.nonzero(mut)[0], pos[mut]] = 1
out = {"rows": n, "width": width}
d = ctx.upload(t)
for name, table in (("adversarial", d),):
    for it in range(2):
        ctx.sync(); ctx.timing(True); ctx.timing_reset()
        t0 = time.perf_counter()
        perm, key, uniq, nu = ctx.sort_rows(table, want_perm=True, want_key=True, want_uniq=True)
        ctx.sync()
        ms = (time.perf_counter() - t0) * 1e3
        rep = ctx.timing_report()
        for a in (perm, key, uniq): a.free()
    out[name] = {"wall_ms": round(ms, 2), "kernel_ms": round(sum(v[1] for v in rep.values()), 2), "unique": int(nu),
                 "kernels": sorted(([k, v[0], round(v[1], 3)] for k, v in rep.items() if v[1] > 0.05), key=lambda r: -r[2])[:12]}
d.free()
del t
# the bench's quality table at the same number of reads
dev = ctx.synth("genome", n, 150, 1002, genome=10_000_000, pool=max(1, n // 5))
fq = ctx.adopt_fastq(dev)
p = host.prepare(ctx, fq)
for it in range(2):
    ctx.sync(); ctx.timing(True); ctx.timing_reset()
    t0 = time.perf_counter()
    perm, key, uniq, nu = ctx.sort_rows(p["qual"], want_perm=True, want_key=True, want_uniq=True)
    ctx.sync()
    ms = (time.perf_counter() - t0) * 1e3
    rep = ctx.timing_report()
    for a in (perm, key, uniq): a.free()
out["bench_qual"] = {"wall_ms": round(ms, 2), "kernel_ms": round(sum(v[1] for v in rep.values()), 2), "unique": int(nu),
                     "kernels": sorted(([k, v[0], round(v[1], 3)] for k, v in rep.items() if v[1] > 0.05), key=lambda r: -r[2])[:12]}
out["ratio_wall"] = round(out["adversarial"]["wall_ms"] / out["bench_qual"]["wall_ms"], 2)
print(json.dumps(out, indent=1))
