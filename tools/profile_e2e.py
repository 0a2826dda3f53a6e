#!/usr/bin/env python3
"""Diagnostics: host-side timestamps of the end-to-end (host buffer -> host buffers) encode."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uq_b200 import host
from uq_b200.device import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = Context(0)
dev = ctx.synth("genome", n, 150, 1002, genome=10_000_000, pool=max(1, n // 5))
fbytes = dev.nbytes
pin_in = ctx.pinned_empty(fbytes)
ctx.check(ctx.lib.uqb_array_download(ctx.h, dev.h, pin_in.ptr, fbytes))
dev.free()
pin_out = ctx.pinned_empty(int(fbytes * 0.45))
for it in range(3):
    cur = [0]
    def sink(name, nbytes):
        a = pin_out.array[cur[0]:cur[0] + nbytes]; cur[0] += (nbytes + 63) & ~63; return a
    ctx.sync(); t0 = time.perf_counter()
    fq = ctx.load_fastq_streamed(pin_in, chunk_bytes=chunk)
    t1 = time.perf_counter()
    ctx.sync(); t1s = time.perf_counter()
    host.PHASE_LOG = {} if it == 2 else None
    members, cfg = host.encode_device(ctx, fq, sort="DNA", sink=sink)
    t2 = time.perf_counter()
    ctx.sync(); t2s = time.perf_counter()
    members.download()
    t3 = time.perf_counter()
    print("iter %d: load returned %.1f (synced %.1f) encode returned %.1f (synced %.1f) downloads done %.1f ms   H2D %.1f GB/s" % (
        it, (t1 - t0) * 1e3, (t1s - t0) * 1e3, (t2 - t0) * 1e3, (t2s - t0) * 1e3, (t3 - t0) * 1e3, fbytes / 1e9 / (t1s - t0)))
    if host.PHASE_LOG: print("   phases", {k: round(v, 1) for k, v in host.PHASE_LOG.items()})
    members.free(); fq.free()
