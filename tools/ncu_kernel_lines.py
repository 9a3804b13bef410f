#!/usr/bin/env python
"""Per kernel, the source lines with the most stall samples / executed instructions and their stall mix, from an
`ncu --page source --csv --print-source cuda,sass` export.  usage: ncu_kernel_lines.py file.csv kernel-substr [top_n]"""
import csv
import sys
from collections import defaultdict

path, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = csv.reader(open(path, newline=''))
cur_file = names = fn = None
agg = defaultdict(lambda: defaultdict(float))
for r in rows:
    if not r:
        continue
    if r[0] in ('File Path', 'File Name'):
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        fn = r[1]
        continue
    if r[0] == 'Line No':
        names = r
        continue
    if names is None or r[0] == '' or fn is None or want not in fn:
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    n = len(names)
    key = (cur_file, line, r[1].strip()[:100])
    for col in ('# Samples', 'Instructions Executed', 'Thread Instructions Executed', 'stall_long_sb', 'stall_wait', 'stall_no_inst',
                'stall_math', 'stall_short_sb', 'stall_branch_resolving', 'stall_barrier', 'stall_lg', 'stall_mio', 'stall_not_selected',
                'stall_selected', 'stall_dispatch'):
        if col in names:
            try:
                agg[key][col] += float(r[names.index(col) - n] or 0)
            except (ValueError, IndexError):
                pass
ts = sum(v['# Samples'] for v in agg.values()) or 1
ti = sum(v['Instructions Executed'] for v in agg.values()) or 1
print('kernel ~ %s: samples %d, warp instructions %d' % (want, ts, ti))
tot = defaultdict(float)
for v in agg.values():
    for k, x in v.items():
        tot[k] += x
print('stall mix: ' + ', '.join('%s %.1f%%' % (k[6:], 100 * tot[k] / ts) for k in sorted(tot, key=lambda k: -tot[k]) if k.startswith('stall_')))
byfile = defaultdict(lambda: [0.0, 0.0])
for (f, l, src), v in agg.items():
    byfile[f][0] += v['# Samples']
    byfile[f][1] += v['Instructions Executed']
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0])[:8]:
    print('   %-24s samples %5.1f%%  inst %5.1f%%' % (f, 100 * v[0] / ts, 100 * v[1] / ti))
for (f, l, src), v in sorted(agg.items(), key=lambda kv: -kv[1]['# Samples'])[:top]:
    st = sorted(((k[6:], x) for k, x in v.items() if k.startswith('stall_')), key=lambda kx: -kx[1])[:2]
    print('%5.1f%% smp %5.1f%% inst  %-20s %s  | %s' % (100 * v['# Samples'] / ts, 100 * v['Instructions Executed'] / ti, '%s:%d' % (f, l),
                                                       src[:80], ' '.join('%s %.0f%%' % (k, 100 * x / max(v['# Samples'], 1)) for k, x in st)))
