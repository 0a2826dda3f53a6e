#!/usr/bin/env python3
"""Diagnostics: device decode (tables already in HBM -> FASTQ text in HBM), per-kernel event timing."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uq_b200 import host
from uq_b200.device import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
ctx = Context(0)
dev = ctx.synth("genome", n, 150, 1002, genome=10_000_000, pool=max(1, n // 5))
fq = ctx.adopt_fastq(dev)
st = {}
members, cfg = host.encode_device(ctx, fq, sort="None", raw=["DNA", "QUAL", "QNAME"], stages=st)
for it in range(2):
    text = host.decode_device(ctx, st["dna"], st["qual"], st["cols"], cfg)
    text.free()
ctx.sync()
ctx.timing(True); ctx.timing_reset()
text = host.decode_device(ctx, st["dna"], st["qual"], st["cols"], cfg)
rep = ctx.timing_report()
assert text.first_difference(dev) == -1
print(json.dumps(sorted(([k, v[0], round(v[1], 3)] for k, v in rep.items()), key=lambda r: -r[2])))
