#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` of the hot kernels: duration, executed FP64 operations (thread-level
DADD/DMUL/DFMA counts -> flop, DFMA = 2) and their rate, FP64 pipe activity, DRAM bytes, registers, occupancy, issue
slot use and the top warp-stall reasons.  usage: ncu_raw_summary.py report.ncu-rep [...]   (needs ncu on PATH)"""
import csv
import io
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'dadd'),
    ('smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'dmul'),
    ('smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'dfma'),
    ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'fp64 pipe active %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu pipe %'),
    ('smsp__issue_active.avg.pct', 'issue slots busy %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__occupancy_limit_registers', 'CTAs/SM (register limit)'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('l1tex__t_sector_hit_rate.pct', 'L1 hit %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('smsp__thread_inst_executed_per_inst_executed.ratio', 'active threads / warp instruction'),
]
UNIT_SCALE = {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0, 'nsecond': 1e-9, 'usecond': 1e-6, 'msecond': 1e-3, 'second': 1.0, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6,
              'Gbyte': 1e9, 'Tbyte': 1e12}


def num(text):
    return float(text.replace(',', ''))


for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(head)}
    for r in rows[2:]:
        print('== %s :: %s  grid %s block %s' % (path.split('/')[-1], r[col['Kernel Name']][:70], r[col['Grid Size']], r[col['Block Size']]))
        vals = {}
        for metric, label in WANT:
            if metric not in col or r[col[metric]] == '':
                continue
            v, u = num(r[col[metric]]), units[col[metric]]
            vals[label] = v * UNIT_SCALE.get(u, 1.0)
            print('   %-38s %18s %s' % (label, r[col[metric]], u))
        if all(k in vals for k in ('dadd', 'dmul', 'dfma', 'duration')):
            flop = vals['dadd'] + vals['dmul'] + 2.0 * vals['dfma']
            print('   %-38s %18.3f TFLOP/s  (%.3e flop; DFMA share of FP64 instructions %.0f %%)' % (
                'executed FP64 rate', flop / vals['duration'] / 1e12, flop,
                100.0 * vals['dfma'] / max(vals['dadd'] + vals['dmul'] + vals['dfma'], 1.0)))
        if 'dram read' in vals and 'duration' in vals:
            print('   %-38s %18.1f GB/s' % ('dram throughput (read + write)', (vals['dram read'] + vals['dram write']) / vals['duration'] / 1e9))
        stalls = []
        for name, i in col.items():
            if name.startswith('smsp__average_warps_issue_stalled_') and name.endswith('_per_issue_active.ratio') and r[i] != '':
                stalls.append((num(r[i]), name[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
        if not stalls:
            for name, i in col.items():
                if name.startswith('smsp__average_warp_latency_issue_stalled_') and r[i] != '':
                    stalls.append((num(r[i]), name[len('smsp__average_warp_latency_issue_stalled_'):].split('.')[0]))
        stalls.sort(reverse=True)
        if stalls:
            print('   top stalls (warps per issue): ' + ', '.join('%s %.2f' % (n, v) for v, n in stalls[:6]))
        print()
