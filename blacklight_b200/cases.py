"""Synthetic workloads of the benchmark and the tests: a parameter file in the reference's schema (templates with the
values of the reference's example inputs in blacklight_b200/inputs/) plus, for the simulation model, a mock Athena++
snapshot written next to it (blacklight_b200/mock_snapshot.py).  No oracle or test code is involved: bench.py builds
its inputs through this module; the tests extend Case with the reference runs (tests/harness.py)."""
import os

import numpy as np

from . import Config, parse_input_text, run_input_file
from . import mock_snapshot

INPUTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'inputs')

# BASELINE.json configs[3]: polarized kappa-distribution synchrotron, four frequencies, log-spaced 86-345 GHz
C4_PHYSICS = {'image_polarization': 'true', 'image_num_frequencies': 4, 'image_frequency_start': '8.6e10',
              'image_frequency_end': '3.45e11', 'image_frequency_spacing': 'log', 'plasma_kappa_frac': '1.0',
              'plasma_kappa': '4.0', 'plasma_w': '1.0'}


def load_input(name):
    with open(os.path.join(INPUTS, name)) as f:
        return parse_input_text(f.read())


def write_input(path, kv):
    with open(path, 'w') as f:
        for k, v in kv.items():
            f.write('%s = %s\n' % (k, v))


class Case:
    """One configuration in its own directory: <dir>/<tag>.input, <dir>/data/mock.athdf, <dir>/out_<tag>/"""

    def __init__(self, workdir, base, overrides=None, mock=None, threads=None):
        self.dir = str(workdir)
        os.makedirs(os.path.join(self.dir, 'data'), exist_ok=True)
        self.kv = load_input(base)
        self.kv.update({k: str(v) for k, v in (overrides or {}).items()})
        self.kv['num_threads'] = str(threads or os.cpu_count() or 1)
        self.grid = None
        self.sim = self.kv['model_type'] == 'simulation'
        if self.sim:
            mock = dict(mock or {})
            blocks = tuple(mock.pop('blocks', (1, 1, 1)))
            self.kv['simulation_file'] = os.path.join(self.dir, 'data', 'mock.athdf')
            self.grid = mock_snapshot.make_mock(self.kv['simulation_file'], blocks, **mock)

    def _input(self, tag, extra):
        kv = dict(self.kv)
        out = os.path.join(self.dir, 'out_' + tag)
        os.makedirs(out, exist_ok=True)
        kv['output_file'] = os.path.join(out, 'image.npz')
        kv.update(extra)
        path = os.path.join(self.dir, tag + '.input')
        write_input(path, kv)
        return path, out

    def config(self, device=0, tile_rays=0, extra=None):
        path, _ = self._input('gpu', extra or {})
        return Config(path, device=device, tile_rays=tile_rays)

    def run_gpu_file(self, device=0, extra=None, tag='gpufile', devices=None):
        """Full drop-in run through blh_run_input_file (devices: a list of CUDA ordinals -> blh_run_input_file_devices);
        returns (npz dict, timings)."""
        path, out = self._input(tag, extra or {})
        t = run_input_file(path, device=device, devices=devices)
        return dict(np.load(os.path.join(out, 'image.npz'))), t

    def grid_arrays(self):
        return mock_snapshot.grid_view_arrays(self.grid)


def parse_timers(stdout):
    """The five-line timing report of the reference (and of the drop-in executable): name -> seconds."""
    t = {}
    for line in stdout.splitlines():
        if ':' in line and line.strip().endswith(' s'):
            k, v = line.rsplit(':', 1)
            try:
                t[k.strip()] = float(v.strip()[:-2])
            except ValueError:
                pass
    return t
