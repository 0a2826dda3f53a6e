#!/usr/bin/env python3
"""Diagnostics: uqb_layout / uqb_unlayout (write_pattern uq.py:257-270 and its inverse) for the seven non-trivial patterns
at row widths from 25 bytes (100 bp DNA) to 17 501 bytes (20 kb QUAL): GB/s of table bytes read + stream bytes written."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from uq_b200 import host
from uq_b200.device import Context

ctx = Context(0)
target = int(float(sys.argv[1]) * 1e9) if len(sys.argv) > 1 else 2_000_000_000
out = {}
for width in (25, 38, 113, 190, 1001, 5001, 17501):
    n = max(1, target // width)
    table = ctx.alloc(n, width)
    row = {}
    for pat in host.PATTERNS:
        if pat == '0.1':
            continue
        s = ctx.layout(table, pat); s.free()                    # warm-up
        ctx.sync()
        ctx.span_begin()
        for _ in range(3):
            s = ctx.layout(table, pat); s.free()
        ms = ctx.span_end() / 3
        row[pat] = round(2 * n * width / 1e9 / (ms / 1e3), 1)
    s = ctx.layout(table, '2.2')
    ctx.sync(); ctx.span_begin()
    for _ in range(3):
        t = ctx.unlayout(s, n, width, '2.2'); t.free()
    ms = ctx.span_end() / 3
    row['unlayout 2.2'] = round(2 * n * width / 1e9 / (ms / 1e3), 1)
    s.free(); table.free()
    out["width %d x %d rows" % (width, n)] = row
    print(width, row, file=sys.stderr)
print(json.dumps({"unit": "GB/s (read + written)", "layouts": out}, indent=1))
