#!/usr/bin/env python3
"""Attribute executed warp-instructions of an ncu capture to CUDA source lines.
usage: ncu_lines.py report.ncu-rep library.so kernel_mangled_substring [launch_index] [top_n]
Joins the ncu source page (per-SASS-row executed counts) with `nvdisasm -g` line info of the same binary by row order."""
import csv, io, re, subprocess, sys, tempfile, os
from collections import Counter

rep, so, ksub = sys.argv[1], sys.argv[2], sys.argv[3]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 45
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split('\n')
start = [i for i, l in enumerate(dis) if l.startswith('.text.') and ksub in l][0]
end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith('//--------------------- .text.')), len(dis))
lines = []  # (file, line) per SASS instruction, in order
cur = None
for l in dis[start:end]:
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    elif re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append(cur)
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hidx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
i0 = hidx[which]; i1 = hidx[which + 1] if which + 1 < len(hidx) else len(rows)
h = rows[i0]
ci, ct, cn = h.index('Instructions Executed'), h.index('Thread Instructions Executed'), h.index('# Samples')
body = [r for r in rows[i0 + 1:i1] if len(r) > ci and r[ci].isdigit()]
print('sass rows: ncu %d, nvdisasm %d' % (len(body), len(lines)))
n = min(len(body), len(lines))
ex, th, sm = Counter(), Counter(), Counter()
for r, ln in zip(body[:n], lines[:n]):
    ex[ln] += int(r[ci]); th[ln] += int(r[ct]); sm[ln] += int(r[cn] or 0)
tot, tots = sum(ex.values()), sum(sm.values())
srcs = {}
print('total warp-instr %d, thread-instr %d (avg active %.1f)' % (tot, sum(th.values()), sum(th.values()) / max(tot, 1)))
for ln, v in ex.most_common(topn):
    text = ''
    if ln:
        f = ln[0]
        if f not in srcs:
            for d in ('tennisbot_rl_b200/csrc', 'include'):
                pth = os.path.join(d, f)
                if os.path.exists(pth):
                    srcs[f] = open(pth).read().split('\n'); break
            else:
                srcs[f] = []
        if ln[1] - 1 < len(srcs[f]):
            text = srcs[f][ln[1] - 1].strip()[:90]
    print('%5.2f%% smp %5.2f%% thr %4.1f  %s:%s  %s' % (100 * v / tot, 100 * sm[ln] / max(tots, 1), th[ln] / max(v, 1), ln[0] if ln else '?', ln[1] if ln else '', text))
