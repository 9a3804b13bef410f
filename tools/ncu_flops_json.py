#!/usr/bin/env python
"""Build profiles/executed_flops.json: executed FP64 work per unit of each hot kernel, per bench workload.

Input (written under gpurun_out/ by tools/ncu_capture.sh <tag>): for each workload W
  <tag>_flops_W.csv   `ncu --csv` of every launch of `bench.py --workload W ...`: thread-level DADD / DMUL / DFMA counts,
                      duration, FP64 pipe activity and DRAM bytes of each launch
  <tag>_units_W.json  the unit counts of the same process (bench.py --dump-units): samples, DP attempts, passes
Output: for every kernel, flop = DADD + DMUL + 2 DFMA summed over all launches, divided by passes x units (a DP attempt
for the geodesic kernel, a stored sample for the radiation kernels, all frequencies of the workload included), the
duration-weighted FP64 pipe activity, DRAM bytes per unit, and the hash of the kernel sources the numbers belong to
(bench.py marks the file stale when the sources have changed since).

usage: tools/ncu_flops_json.py <tag> [workload ...]"""
import collections
import csv
import datetime
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import EXECUTED_JSON, source_hash  # noqa: E402

METRIC = {'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum': 'dadd', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum': 'dmul',
          'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum': 'dfma', 'gpu__time_duration.sum': 'ns',
          'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active': 'pipe', 'dram__bytes_read.sum': 'rd', 'dram__bytes_write.sum': 'wr'}
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1.0, 'us': 1e3, 'ms': 1e6, 'nsecond': 1.0, 'usecond': 1e3,
         'msecond': 1e6, 'second': 1e9}


def main():
    tag = sys.argv[1]
    out = {'comment': 'executed FP64 work per unit, from ncu (tools/ncu_capture.sh, tools/ncu_flops_json.py); read by bench.py',
           'captured': '%s, gpurun_out/%s_flops_*.csv' % (datetime.date.today().isoformat(), tag), 'source_hash': source_hash(),
           'entries': {}}
    try:
        old = json.load(open(EXECUTED_JSON))
        if old.get('source_hash') == out['source_hash']:
            out['entries'] = old.get('entries', {})
    except (OSError, ValueError):
        pass
    names = sys.argv[2:] or sorted(f[len(tag) + 7:-4] for f in os.listdir(os.path.join(ROOT, 'gpurun_out'))
                                   if f.startswith(tag + '_flops_') and f.endswith('.csv'))
    for w in names:
        txt = open(os.path.join(ROOT, 'gpurun_out', '%s_flops_%s.csv' % (tag, w))).read()
        units = json.load(open(os.path.join(ROOT, 'gpurun_out', '%s_units_%s.json' % (tag, w))))
        acc = collections.defaultdict(lambda: collections.defaultdict(float))
        launches = collections.Counter()
        per_launch = {}
        for r in csv.DictReader(txt[txt.index('"ID"'):].splitlines()):
            k = re.sub(r'<.*', '', re.sub(r'void <unnamed>::', '', r['Kernel Name']).split('(')[0])
            m = METRIC.get(r['Metric Name'])
            if m is None:
                continue
            v = float(r['Metric Value'].replace(',', '')) * SCALE.get(r['Metric Unit'], 1.0)
            if m == 'pipe':
                per_launch[(k, r['ID'])] = v
                continue
            acc[k][m] += v
            if m == 'ns':
                launches[k] += 1
                per_launch[(k, r['ID'], 'ns')] = v
        st = units['stats_rank0']
        entry = {}
        for k, m in acc.items():
            if not (k.startswith('geodesic') or k.startswith('pol_') or k.startswith('radiate_')):
                continue
            unit = 'attempt' if k.startswith('geodesic_dp') else 'sample'
            n = units['passes'] * (st['num_attempts'] if unit == 'attempt' else st['num_samples'])
            flop = m['dadd'] + m['dmul'] + 2.0 * m['dfma']
            pipe = sum(per_launch[(k, i)] * per_launch[(k, i, 'ns')] for (kk, i, *rest) in list(per_launch) if kk == k and not rest
                       and (k, i, 'ns') in per_launch)
            entry[k] = {'flop_per_unit': flop / n, 'unit': unit, 'dfma_share_of_fp64_instructions': m['dfma'] / max(m['dadd'] + m['dmul'] + m['dfma'], 1.0),
                        'fp64_pipe_active': pipe / m['ns'] / 100.0 if m['ns'] > 0 and pipe > 0 else None,
                        'dram_bytes_per_unit': (m['rd'] + m['wr']) / n if (m['rd'] + m['wr']) > 0 else None,
                        'tflops_under_ncu': flop / (m['ns'] * 1e-9) / 1e12, 'launches': launches[k],
                        'captured_at': '%dx%d, %d frequencies' % (units['resolution'], units['resolution'], units['frequencies'])}
        out['entries'][w] = entry
        print(w, json.dumps(entry, indent=1))
    with open(EXECUTED_JSON, 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write('\n')


if __name__ == '__main__':
    main()
