#!/usr/bin/env python
"""Per-kernel share of the device time in an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: ncu_launch_shares.py launches.csv"""
import collections
import csv
import re
import sys

txt = open(sys.argv[1]).read()
rows = csv.DictReader(txt[txt.index('"ID"'):].splitlines())
ns = collections.Counter()
count = collections.Counter()
scale = {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 'nsecond': 1.0, 'usecond': 1e3, 'msecond': 1e6, 'second': 1e9}
for r in rows:
    if r['Metric Name'] != 'gpu__time_duration.sum':
        continue
    k = re.sub(r'<.*', '', r['Kernel Name'].replace('void ', '').replace('<unnamed>::', '').split('(')[0])
    ns[k] += float(r['Metric Value'].replace(',', '')) * scale.get(r['Metric Unit'], 1.0)
    count[k] += 1
total = sum(ns.values())
print('%d launches, %.1f ms of kernel time (under ncu: cold caches, serialised)' % (sum(count.values()), total / 1e6))
for k, v in ns.most_common():
    print('%-32s %6d launches  %9.2f ms  %5.1f %%' % (k, count[k], v / 1e6, 100.0 * v / total))
