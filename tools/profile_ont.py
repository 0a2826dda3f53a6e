#!/usr/bin/env python3
"""Diagnostics: per-phase / per-kernel timing of configs[4] (1 M variable-length reads of 1-20 kb, --sort QUAL)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth
from uq_b200 import host
from uq_b200.device import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = Context(0)
dev = ctx.synth("ont", n, (1000, 20000), 1005, len_table=synth.ont_length_table(1000, 20000))
def step():
    fq = ctx.adopt_fastq(dev)
    m, cfg = host.encode_device(ctx, fq, sort="QUAL")
    m.free(); fq.free()
step()
ctx.timing(True); ctx.timing_reset()
host.PHASE_LOG = {}
step()
rep = ctx.timing_report()
print(json.dumps({"fastq_gb": dev.nbytes / 1e9, "phases_ms": host.PHASE_LOG,
                  "kernels": sorted(([k, v[0], round(v[1], 3)] for k, v in rep.items()), key=lambda r: -r[2])[:14]}))
