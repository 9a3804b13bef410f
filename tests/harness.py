"""Shared helpers for the parity tests: build a work directory with an input file (+ mock snapshot),
run the unmodified reference (oracle/_ref/blacklight) and/or the CUDA path on it, load results."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

import blacklight_b200 as bl  # noqa: E402
from blacklight_b200 import mock_snapshot  # noqa: E402

INPUTS = os.path.join(ROOT, 'tests', 'inputs')
REF_BIN = os.path.join(ROOT, 'oracle', '_ref', 'blacklight')


def load_input(name):
    with open(os.path.join(INPUTS, name)) as f:
        return bl.parse_input_text(f.read())


def write_input(path, kv):
    with open(path, 'w') as f:
        for k, v in kv.items():
            f.write('%s = %s\n' % (k, v))


class Case:
    """One configuration in its own directory: <dir>/case.input, <dir>/data/mock.athdf, <dir>/out_*/"""

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

    def run_reference(self, checkpoints=True):
        """Run the unmodified reference; returns dict(npz=..., geo=..., samp=..., timers=...)."""
        extra = {}
        if checkpoints:
            extra.update({'checkpoint_geodesic_save': 'true', 'checkpoint_geodesic_load': 'false',
                          'checkpoint_geodesic_file': os.path.join(self.dir, 'out_ref', 'geo.ckpt')})
            if self.sim:
                extra.update({'checkpoint_sample_save': 'true', 'checkpoint_sample_load': 'false',
                              'checkpoint_sample_file': os.path.join(self.dir, 'out_ref', 'samp.ckpt')})
        path, out = self._input('ref', extra)
        proc = subprocess.run([REF_BIN, path], cwd=self.dir, capture_output=True, text=True, timeout=3600)
        if proc.returncode != 0 or 'Calculation completed' not in proc.stdout:
            raise RuntimeError('reference failed: ' + proc.stdout + proc.stderr)
        res = {'npz': dict(np.load(os.path.join(out, 'image.npz'))), 'stdout': proc.stdout, 'stderr': proc.stderr}
        res['timers'] = parse_timers(proc.stdout)
        if checkpoints:
            import refio  # oracle/: checkpoint readers, reference runs only
            res['geo'] = refio.read_geodesic_checkpoint(extra['checkpoint_geodesic_file'])
            if self.sim:
                res['samp'] = refio.read_sample_checkpoint(extra['checkpoint_sample_file'],
                                                           interp=self.kv['simulation_interp'] == 'true')
        return res

    def config(self, device=0, tile_rays=0, extra=None):
        path, _ = self._input('gpu', extra or {})
        return bl.Config(path, device=device, tile_rays=tile_rays)

    def run_gpu_file(self, device=0, extra=None, tag='gpufile'):
        """Full drop-in run through blh_run_input_file; returns (npz dict, timings)."""
        path, out = self._input(tag, extra or {})
        t = bl.run_input_file(path, device=device)
        return dict(np.load(os.path.join(out, 'image.npz'))), t

    def grid_arrays(self):
        return mock_snapshot.grid_view_arrays(self.grid)


def parse_timers(stdout):
    t = {}
    for line in stdout.splitlines():
        if ':' in line and line.strip().endswith(' s'):
            k, v = line.rsplit(':', 1)
            try:
                t[k.strip()] = float(v.strip()[:-2])
            except ValueError:
                pass
    return t


def rel_err(a, b, floor_frac=1e-12):
    """Per-pixel relative difference with an absolute floor of floor_frac * max|b|; NaN patterns must agree."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), 'NaN pattern differs'
    ok = ~nan_a
    if not ok.any():
        return 0.0
    scale = np.maximum(np.abs(b[ok]), floor_frac * np.nanmax(np.abs(b)))
    scale = np.where(scale > 0, scale, 1.0)
    return float(np.max(np.abs(a[ok] - b[ok]) / scale))


def flux_rel(a, b):
    fa, fb = np.nanmean(a), np.nanmean(b)
    return abs(fa - fb) / abs(fb) if fb != 0 else abs(fa)
