#!/usr/bin/env python
"""Per-source-line stall reasons from `ncu --page source --csv --print-source cuda,sass`.
usage: ncu_line_stalls.py file.csv.gz [top_n]"""
import csv
import gzip
import sys
from collections import defaultdict

path = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
op = gzip.open if path.endswith('.gz') else open
f = None
names = None
L = []
for r in csv.reader(op(path, 'rt', newline='')):
    if not r:
        continue
    if r[0] == 'File Path':
        f = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        names = r
        n = len(r)
        continue
    if r[0] in ('', 'Function Name') or names is None:
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    off = len(r) - n
    d = {}
    for i in range(4, n):
        d.setdefault(names[i], r[i + off])
    try:
        smp = int(d['# Samples'])
        inst = int(d['Instructions Executed'])
    except (KeyError, ValueError):
        continue
    st = {k: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit()}
    L.append((f, ln, r[1][:70], smp, inst, st))
T = sum(x[3] for x in L) or 1
I = sum(x[4] for x in L) or 1
agg = defaultdict(int)
for x in L:
    for k, v in x[5].items():
        agg[k] += v
print('total samples', T, 'warp instructions', I)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]:
    print('  %-24s %6.2f%%' % (k, 100 * v / T))
print()
for x in sorted(L, key=lambda x: -x[3])[:top_n]:
    top = sorted(x[5].items(), key=lambda kv: -kv[1])[:4]
    print('%5.2f%% smp %5.2f%% inst %s:%d %s | %s' % (100 * x[3] / T, 100 * x[4] / I, x[0], x[1], x[2][:50],
          ' '.join('%s=%d%%' % (k[6:], 100 * v / max(x[3], 1)) for k, v in top)))
