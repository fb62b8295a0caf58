#!/usr/bin/env python3
"""Summarise an .ncu-rep: headline counters per launch, stall reasons, opcode mix and hottest code regions.
usage: ncu_summary.py report.ncu-rep [launch_index]"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread', 'launch__grid_size',
        'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('=== launch', r[hdr.index('ID')], r[hdr.index('Kernel Name')][:70])
    for k in KEYS:
        if k in hdr:
            print('  %-70s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
    st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i]]
    if not st:
        st = [(float(r[i]), h) for i, h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('.pct') and r[i]]
    for v, h in sorted(st, reverse=True)[:8]:
        print('  stall %-64s %.3f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__warp_issue_stalled_', ''), v))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hidx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
i0 = hidx[which]; i1 = hidx[which + 1] if which + 1 < len(hidx) else len(rows)
h = rows[i0]
ci, cs, ct, cn = h.index('Instructions Executed'), h.index('Source'), h.index('Avg. Threads Executed'), h.index('# Samples')
body = [r for r in rows[i0 + 1:i1] if len(r) > ci and r[ci].isdigit()]
tot = sum(int(r[ci]) for r in body); ns = sum(int(r[cn] or 0) for r in body)
print('--- source page: %d SASS rows, %d warp-instr executed, %d samples' % (len(body), tot, ns))
c, s = Counter(), Counter()
for r in body:
    t = r[cs].split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    c[op] += int(r[ci]); s[op] += int(r[cn] or 0)
print('opcode mix:', ', '.join('%s %.1f%%/%.1f%%' % (o, 100 * v / tot, 100 * s[o] / max(ns, 1)) for o, v in c.most_common(22)))
B = 100
for k in range(0, len(body), B):
    seg = body[k:k + B]; v = sum(int(r[ci]) for r in seg); sm = sum(int(r[cn] or 0) for r in seg)
    if v / tot > 0.01 or sm / max(ns, 1) > 0.01:
        th = sum(float(r[ct]) * int(r[ci]) for r in seg) / max(1, v)
        print('  rows %4d-%4d  exec %5.1f%%  samples %5.1f%%  avg-threads %4.1f   %s' % (k, k + B, 100 * v / tot, 100 * sm / max(ns, 1), th, seg[0][cs][:60]))
