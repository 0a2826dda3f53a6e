#!/usr/bin/env python3
"""Diagnostics: per-phase wall clock and full per-kernel event timing of one encode (not a benchmark)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uq_b200 import host
from uq_b200.device import Context

# usage: profile_phases.py [reads] [sort] [kind genome|casava|illumina] [read length] [pattern, e.g. 2.2 = raw tables in that layout]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
sort = sys.argv[2] if len(sys.argv) > 2 else "DNA"
kind = sys.argv[3] if len(sys.argv) > 3 else "genome"
length = int(sys.argv[4]) if len(sys.argv) > 4 else 150
pattern = sys.argv[5] if len(sys.argv) > 5 else None
ctx = Context(0)
if kind == "genome":
    dev = ctx.synth("genome", n, length, 1002, genome=10_000_000, pool=max(1, n // 5))
else:
    dev = ctx.synth(kind, n, length, 1003 if kind == "casava" else 1004)
opts = dict(sort=sort) if pattern is None else dict(sort=sort, raw=["DNA", "QUAL", "QNAME"], pattern=[pattern, pattern])
def step():
    fq = ctx.adopt_fastq(dev)
    m, cfg = host.encode_device(ctx, fq, **opts)
    m.free(); fq.free()
step(); step()
ctx.timing(True); ctx.timing_reset()
host.PHASE_LOG = {}
host.PHASE_KERNELS = {}
step()
rep = ctx.timing_report()
per_phase = {ph: sorted(([k, v[0], round(v[1], 3)] for k, v in ks.items()), key=lambda r: -r[2]) for ph, ks in host.PHASE_KERNELS.items()}
print(json.dumps({"workload": "%d %s reads x %d, %s" % (n, kind, length, opts), "fastq_gb": dev.nbytes / 1e9, "phases_ms": host.PHASE_LOG,
                  "phase_kernels": per_phase,
                  "kernels": sorted(([k, v[0], round(v[1], 3), round(v[2] / 1e9 / (v[1] / 1e3), 1) if v[1] > 0 and v[2] else None] for k, v in rep.items()), key=lambda r: -r[2])}, indent=1))
