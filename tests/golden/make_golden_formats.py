#!/usr/bin/env python
"""Golden fixtures for the snapshot formats other than Athena++ .athdf: the UNMODIFIED reference
(oracle/_ref/blacklight) run on deterministic mock dumps written by blacklight_b200/mock_snapshot.py.

Run where the reference is built:  python tests/golden/make_golden_formats.py
Each fixture formats_<name>.npz holds the reference's image arrays; FORMAT_CASES is imported by the parity test, which
rebuilds the same dump and renders it through the drop-in executable path.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), 'oracle'))   # refio: generation only

FMKS = dict(poly_xt=0.82, poly_alpha=14.0, mks_smooth=0.5)
POL = {'image_polarization': 'true'}
# name -> (format, simulation_coord, writer keyword arguments, input overrides)
FORMAT_CASES = {
    'athenak_16': ('athenak', 'cks', dict(location_size=4, variable_size=4), {'camera_resolution': 16, 'simulation_a': '0.5'}),
    'athenak_f64_pol_12': ('athenak', 'cks', dict(location_size=8, variable_size=8),
                           dict(POL, camera_resolution=12, simulation_a='0.5')),
    'iharm3d_mks_16': ('iharm3d', 'sks', dict(hslope=0.7), {'camera_resolution': 16}),
    'iharm3d_fmks_nearest_16': ('iharm3d', 'fmks', dict(hslope=0.3, fmks=FMKS), {'camera_resolution': 16, 'simulation_interp': 'false'}),
    'iharm3d_fmks_pol_12': ('iharm3d', 'fmks', dict(hslope=0.3, fmks=FMKS), dict(POL, camera_resolution=12, simulation_interp='false')),
    'harm3d_16': ('harm3d', 'sks', {}, {'camera_resolution': 16}),
}


def write_case(name, workdir):
    """Write the case's mock dump and input file into workdir; returns (input path, output npz path)."""
    from blacklight_b200 import mock_snapshot as ms
    from harness import load_input, write_input
    fmt, coord, kw, over = FORMAT_CASES[name]
    snap = os.path.join(workdir, 'mock.' + fmt)
    if fmt == 'athenak':
        ms.write_athenak(snap, ms.to_blocks(ms.mock_fields_cks(n=24), (2, 1, 2)), gamma_adi=13.0 / 9.0, time=1.0, spin=0.5, **kw)
    elif fmt == 'iharm3d':
        ms.write_iharm3d(snap, n_r=40, n_th=24, n_ph=16, gamma_adi=13.0 / 9.0, time=2.0, **kw)
    else:
        ms.write_harm3d(snap, ms.mock_fields(n_r=40, n_th=24, n_ph=16), gamma_adi=13.0 / 9.0, time=3.0)
    kv = load_input('simulation.input')
    kv.update({k: str(v) for k, v in over.items()})
    kv.update({'simulation_format': fmt, 'simulation_file': snap, 'simulation_coord': coord, 'num_threads': '4',
               'output_file': os.path.join(workdir, 'image.npz')})
    if fmt != 'athenak':
        kv['simulation_a'] = '0.0'
        kv.pop('simulation_block_interp', None)
    kv.pop('plasma_gamma', None)
    path = os.path.join(workdir, 'case.input')
    write_input(path, kv)
    return path, kv['output_file']


def dump_crc(input_path):
    """CRC-32 of the mock dump an input file written by write_case points at (pins the generator to the fixture)."""
    import zlib
    for line in open(input_path):
        if line.startswith('simulation_file'):
            with open(line.split('=', 1)[1].strip(), 'rb') as f:
                return zlib.crc32(f.read())
    raise ValueError('no simulation_file in ' + input_path)


def defined_pixels(name, workdir, path, shape):
    """FMKS only: the reference's scaled zone lookup forms indices one zone past the last row / plane
    (simulation_sampling.cpp:412-446); past the last cell of a variable's plane it reads the next variable's first
    cells, and for the last variable whatever follows its array in memory.  Pixels whose rays take such a sample are
    not defined by the reference; they are found here from its own sampling checkpoint (unpolarized run, same rays)."""
    import refio
    from harness import REF_BIN
    fmt, coord, kw, over = FORMAT_CASES[name]
    if coord != 'fmks':
        return np.ones(shape, bool)
    lines = [ln for ln in open(path).read().splitlines() if not ln.startswith('checkpoint_')]
    lines = [ln.replace('image_polarization = true', 'image_polarization = false') for ln in lines]
    samp, geo = os.path.join(workdir, 'samp.ckpt'), os.path.join(workdir, 'geo.ckpt')
    lines += ['checkpoint_sample_save = true', 'checkpoint_sample_load = false', 'checkpoint_sample_file = ' + samp,
              'checkpoint_geodesic_save = true', 'checkpoint_geodesic_load = false', 'checkpoint_geodesic_file = ' + geo]
    taps = os.path.join(workdir, 'taps.input')
    with open(taps, 'w') as f:
        f.write('\n'.join(lines) + '\n')
    proc = subprocess.run([REF_BIN, taps], cwd=workdir, capture_output=True, text=True, timeout=3600)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    interp = over.get('simulation_interp', 'true') == 'true'
    s, g = refio.read_sample_checkpoint(samp, interp=interp), refio.read_geodesic_checkpoint(geo)
    n_r, n_th, n_ph = 40, 24, 16
    S = s['sample_nan'].shape[1]
    x, y, z = (g['sample_pos'][..., c] for c in (1, 2, 3))
    valid = (np.arange(S)[None, :] < g['sample_num'][:, None]) & (s['sample_nan'] == 0) & (np.sqrt(x * x + y * y + z * z) <= 50.0)
    k, j, i = (s['sample_inds'][..., c].astype(np.int64) for c in (1, 2, 3))
    reach = 0 if not interp else n_th * n_r + n_r + 1          # the farthest corner of a trilinear stencil
    past = valid & ((k * n_th + j) * n_r + i + reach >= n_ph * n_th * n_r)
    return ~past.any(axis=1).reshape(shape)


def main():
    from harness import REF_BIN
    for name in (sys.argv[1:] or FORMAT_CASES):
        with tempfile.TemporaryDirectory() as d:
            path, out = write_case(name, d)
            proc = subprocess.run([REF_BIN, path], cwd=d, capture_output=True, text=True, timeout=3600)
            assert proc.returncode == 0 and 'Calculation completed' in proc.stdout, proc.stdout + proc.stderr
            npz = dict(np.load(out))
            keep = {k: v for k, v in npz.items() if k.endswith('_nu')}
            keep['defined'] = defined_pixels(name, d, path, npz['I_nu'].shape)
            keep['dump_crc'] = np.uint32(dump_crc(path))
            np.savez_compressed(os.path.join(HERE, 'formats_%s.npz' % name), **keep)
            print(name, {k: v.shape for k, v in keep.items() if k.endswith('_nu')}, float(np.nanmax(keep['I_nu'])), 'undefined pixels', int((~keep['defined']).sum()))


if __name__ == '__main__':
    main()
