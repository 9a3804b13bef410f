#!/usr/bin/env python
"""The drop-in executable on the C4 frame with one device and with all devices of the box (BLACKLIGHT_DEVICES): wall
times of the whole run (read, trace, radiate, write) and a bitwise comparison of the two output files.
usage: dropin_multi.py [resolution]"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blacklight_b200.cases import C4_PHYSICS, Case  # noqa: E402
import blacklight_b200 as bl  # noqa: E402

res = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
exe = os.path.join(ROOT, 'blacklight_b200', 'bin', 'blacklight_b200')
n_dev = bl.load_library().bl_device_count()
with tempfile.TemporaryDirectory() as d:
    case = Case(d, 'simulation.input', dict(C4_PHYSICS, camera_resolution=res))
    out = {}
    for tag, devs in (('one', '0'), ('all', '0-%d' % (n_dev - 1))):
        path, outdir = case._input(tag, {})
        env = dict(os.environ, BLACKLIGHT_DEVICES=devs)
        t0 = time.perf_counter()
        p = subprocess.run([exe, path], cwd=d, env=env, capture_output=True, text=True)
        dt = time.perf_counter() - t0
        if p.returncode != 0:
            print(p.stdout[-2000:], p.stderr[-2000:])
            raise SystemExit('drop-in run failed on devices ' + devs)
        out[tag] = dict(np.load(os.path.join(outdir, 'image.npz')))
        print('devices %-5s: %.2f s wall for the whole executable; %s' % (devs, dt, ' | '.join(
            l.strip() for l in p.stdout.splitlines() if 'Integrating' in l or 'Elapsed' in l)))
    same = sorted(out['one']) == sorted(out['all']) and all(
        np.array_equal(out['one'][k], out['all'][k], equal_nan=True) for k in out['one'])
    print('%d devices, %dx%d C4 frame: outputs bitwise identical: %s' % (n_dev, res, res, same))
    if not same:
        raise SystemExit(1)
