#!/usr/bin/env python3
"""Diagnostics: the --test sweep (uq.py:855-889) served from HBM (host.MixFeed) against one fresh encode per mix.
usage: profile_testfeed.py [reads]   (configs[3]: 20 M reads x 100 bp; 8 raw sets x 4 sorts x 8 patterns)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uq_b200 import host
from uq_b200.device import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
ctx = Context(0)
dev = ctx.synth("illumina", n, 100, 1004)
pats = host.PATTERNS
raws = [('DNA', 'QUAL', 'QNAME'), ('DNA', 'QUAL'), ('QUAL', 'QNAME'), ('DNA', 'QNAME'), ('DNA',), ('QUAL',), ('QNAME',), (None,)]
sorts = ['DNA', 'QUAL', 'QNAME', None]

def fresh(sort, raw, pat):
    fq = ctx.adopt_fastq(dev)
    m, cfg = host.encode_device(ctx, fq, sort=sort if sort else 'None', raw=[r if r else 'none' for r in raw], pattern=[pat, pat])
    out = m.download()
    m.free(); fq.free()
    return out

fresh('DNA', raws[1], '2.2')                                  # warm-up
t0 = time.perf_counter()
k = 0
for raw in raws[:2]:                                          # a sample of the 256 fresh encodes (they all cost about the same)
    for sort in sorts:
        for pat in pats[:2]:
            fresh(sort, raw, pat); k += 1
ctx.sync()
ms_fresh = (time.perf_counter() - t0) * 1e3 / k

fq = ctx.adopt_fastq(dev)
t0 = time.perf_counter()
feed = host.MixFeed(ctx, fq)
ctx.sync()
ms_prepare = (time.perf_counter() - t0) * 1e3
l0 = ctx.launches
t0 = time.perf_counter()
mixes = members = 0
for raw in raws:
    for sort in sorts:
        for pat in pats:
            m = feed.members(sort=sort if sort else 'None', raw=[r if r else 'none' for r in raw], pattern=[pat, pat])
            mixes += 1; members += len(m)
ctx.sync()
ms_sweep = (time.perf_counter() - t0) * 1e3
print(json.dumps({"workload": "configs[3] --test sweep: %d reads x 100 bp, 8 raw sets x 4 sorts x 8 patterns = %d mixes, %d members served" % (n, mixes, members),
                  "fresh_encode_ms_per_mix": round(ms_fresh, 1), "feed_prepare_ms_once": round(ms_prepare, 1),
                  "feed_sweep_ms_total": round(ms_sweep, 1), "feed_ms_per_mix": round((ms_prepare + ms_sweep) / mixes, 2),
                  "kernel_launches_in_sweep": ctx.launches - l0,
                  "note": "both sides include the device->host copy of every member (that is what the compressor reads); "
                          "the feed downloads each distinct member once"}, indent=1))
