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

from blacklight_b200 import cases  # noqa: E402
from blacklight_b200.cases import load_input, parse_timers, write_input  # noqa: E402,F401

REF_BIN = os.path.join(ROOT, 'oracle', '_ref', 'blacklight')


class Case(cases.Case):
    """The benchmark's workload case plus the checker: a run of the unmodified reference on the same inputs."""

    def run_reference(self, checkpoints=True):
        """Run the unmodified reference; returns dict(npz=..., geo=..., samp=..., timers=...)."""
        extra = {}
        sample_only = checkpoints == 'sample'   # large cases: the geodesic checkpoint would be several GB
        if checkpoints:
            if not sample_only:
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
            if not sample_only:
                res['geo'] = refio.read_geodesic_checkpoint(extra['checkpoint_geodesic_file'])
            if self.sim:
                res['samp'] = refio.read_sample_checkpoint(extra['checkpoint_sample_file'],
                                                           interp=self.kv['simulation_interp'] == 'true')
        return res


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
