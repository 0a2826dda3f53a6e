#!/usr/bin/env python3
"""Diagnostics: per-source-line instruction counts and stall samples of one kernel of an .ncu-rep.

ncu's offline `--page source --csv` lists SASS only; this joins it (by instruction order) with the line table that
`nvdisasm -g` prints for the same kernel of the in-tree library, and aggregates per source line.

usage: ncu_lines.py REPORT.ncu-rep KERNEL_REGEX MANGLED_SUBSTRING [cubin-name-substring] [launch-skip]
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, kregex, mangled = sys.argv[1:4]
cub = sys.argv[4] if len(sys.argv) > 4 else ""
skip = sys.argv[5] if len(sys.argv) > 5 else "0"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "uq_b200", "libuqb200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
lines = None
for f in sorted(os.listdir(tmp)):
    if cub and cub not in f:
        continue
    out = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    m = re.search(r"^\.text\.(\S*%s\S*):\n(.*?)(?=^//-+ \.|\Z)" % re.escape(mangled), out, re.S | re.M)
    if m:
        cur, lines = None, []
        for l in m.group(2).splitlines():
            mm = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if mm:
                cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            elif re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
                lines.append((cur, l.split("*/", 1)[1].strip()))
        break
if lines is None:
    sys.exit("kernel %s not found in the cubins" % mangled)
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# the page holds one section per matching launch (the filters are not applied to an imported report): keep the
# heaviest section
heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[heads[0]]
ci = hdr.index("Instructions Executed")
sections = [[r for r in rows[a + 1:b] if len(r) == len(hdr)] for a, b in zip(heads, heads[1:] + [len(rows)])]
body = max(sections, key=lambda sec: sum(int(r[ci] or 0) for r in sec))
ci, cs, cn = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Warp Stall Sampling (Not-issued Samples)")
if len(body) != len(lines):
    print("warning: %d SASS rows in the report vs %d in the cubin (rebuilt since the capture?)" % (len(body), len(lines)), file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
for (ln, _), r in zip(lines, body):
    a = agg[ln]
    a[0] += int(r[ci] or 0); a[1] += int(r[cs] or 0); a[2] += int(r[cn] or 0); a[3] += 1
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
src = {}
print("line      inst%  stall%  notissued%  sass  source")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:(100000 if os.environ.get("NCU_LINES_ALL") else 45)]:
    text = ""
    if ln:
        fn = os.path.join(root, "uq_b200", "csrc", ln[0])
        if fn not in src and os.path.exists(fn):
            src[fn] = open(fn).read().splitlines()
        if fn in src and ln[1] <= len(src[fn]):
            text = src[fn][ln[1] - 1].strip()[:110]
    print("%-9s %5.1f  %5.1f   %5.1f      %4d  %s" % ("%s:%d" % (ln[0][:3], ln[1]) if ln else "?", 100 * a[0] / ti, 100 * a[1] / ts, 100 * a[2] / ts, a[3], text))
print("total warp instructions:", ti)
