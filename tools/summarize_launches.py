#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the LAST encode step."""
import collections, csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
data = [(r[ik].split("(")[0], float(r[iv].replace(",", "")), r[iu]) for r in rows[1:] if len(r) > iv]
starts = [i for i, d in enumerate(data) if d[0] == "k_newline_count"]
last = data[starts[-1]:]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v, u in last:
    agg[k][0] += 1
    agg[k][1] += v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
tot = sum(v[1] for v in agg.values())
print("launches in the last step: %d, summed kernel time %.2f ms" % (len(last), tot))
print("%-36s %8s %10s %7s" % ("kernel", "launches", "ms", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-36s %8d %10.3f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
