#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` export: per kernel, the source lines with
the most stall samples / executed instructions.  usage: ncu_source_summary.py file.csv[.gz] [top_n] [kernel-substr]"""
import csv
import gzip
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ''
op = gzip.open if path.endswith('.gz') else open
rows = csv.reader(op(path, 'rt', newline=''))
cur_file, cur_fn, hdr = None, None, None
data = defaultdict(lambda: defaultdict(lambda: [0, 0, 0, '']))  # fn -> (file,line) -> [samples, inst, thread_inst, src]
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1]
        continue
    if r[0] == 'Function Name':
        cur_fn = r[1]
        continue
    if r[0] == 'Line No':
        hdr = {n: i for i, n in enumerate(r)}
        continue
    if hdr is None or r[0] == '':
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    d = data[cur_fn][(cur_file.split('/')[-1], line)]
    n = len(hdr)
    try:  # source text may contain unescaped quotes: index the metric columns from the end
        d[0] += int(r[hdr['# Samples'] - n] or 0)
        d[1] += int(r[hdr['Instructions Executed'] - n] or 0)
        d[2] += int(r[hdr['Thread Instructions Executed'] - n] or 0)
    except (ValueError, IndexError):
        continue
    d[3] = r[1].strip()[:110]
for fn, lines in data.items():
    if want not in fn:
        continue
    ts = sum(v[0] for v in lines.values()) or 1
    ti = sum(v[1] for v in lines.values()) or 1
    print('==== %s   samples=%d  warp-inst=%d' % (fn[:90], ts, ti))
    byfile = defaultdict(lambda: [0, 0])
    for (f, l), v in lines.items():
        byfile[f][0] += v[0]
        byfile[f][1] += v[1]
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print('   file %-22s samples %5.1f%%  inst %5.1f%%' % (f, 100.0 * v[0] / ts, 100.0 * v[1] / ti))
    for (f, l), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100.0 * v[0] / ts, 100.0 * v[1] / ti, f, l, v[3]))
