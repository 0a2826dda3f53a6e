#!/usr/bin/env python3
"""Diagnostics: one block of headline metrics + top stall reasons per kernel launch of an .ncu-rep (ncu --set full)."""
import csv, io, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']
sel = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
as_json = "--json" in sys.argv
records = []
for r in rows[2:]:
    name = r[idx['Kernel Name']]
    if sel not in name:
        continue
    if as_json:
        rec = {"kernel": name.split("(")[0]}
        for w in want:
            if w in idx:
                rec[w] = "%s %s" % (r[idx[w]], units[idx[w]])
        st = [(h, float(r[i] or 0)) for h, i in idx.items() if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
        st.sort(key=lambda x: -x[1])
        rec["top_stalls_per_issue"] = {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''): round(v, 2) for h, v in st[:6]}
        records.append(rec)
        continue
    print('----', name[:70])
    for w in want:
        if w in idx:
            print('  %-70s %s %s' % (w, r[idx[w]], units[idx[w]]))
    st = [(h, float(r[i] or 0)) for h, i in idx.items() if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
    st.sort(key=lambda x: -x[1])
    print('  stalls/issue:', ', '.join('%s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for h, v in st[:7]))

if as_json:
    import json
    print(json.dumps(records, indent=1))
