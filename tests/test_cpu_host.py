"""CPU-only checks (no compute calls into CUDA): the C ABI library loads and exports what the headers declare,
the host layer (parameter surface, camera, readers/writers) reproduces the reference, the bit-faithful libm
restatements match the host libm, the mock generator is stable, and the multi-rank sharding logic works over
gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys
import zlib

import numpy as np
import pytest

import blacklight_b200 as bl
from harness import ROOT, Case, load_input, write_input
from golden.make_golden import CASES

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def test_library_exports_every_declared_symbol(lib):
    declared = set()
    for header in ('blacklight_b200.h', 'blacklight_b200_host.h'):
        text = open(os.path.join(ROOT, 'include', header)).read()
        text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
        declared |= set(re.findall(r'\b(blh?_[a-z0-9_]+)\s*\(', text))
    assert {'bl_create', 'bl_trace_level', 'bl_radiate_level', 'bl_refine_level', 'bl_upload_grid',
            'blh_run_input_file'} <= declared
    for name in sorted(declared):
        assert hasattr(lib, name), 'symbol %s declared in include/ but not exported' % name


def test_no_gpu_means_loud_failure(tmp_path):
    """Without a CUDA device bl_create must fail with a message, never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    case = Case(tmp_path, 'formula.input', {'camera_resolution': 8})
    cfg = case.config()
    with pytest.raises(bl.BlacklightError, match='CUDA'):
        bl.Context(cfg)


@pytest.mark.parametrize('name', ['formula_16', 'simulation_32', 'simulation_kerr_24', 'formula_pinhole_pole_12', 'adaptive_32'])
def test_camera_bit_exact(name, tmp_path):
    """Camera frame and per-pixel position / covariant momentum / momentum factor vs the reference's checkpoint."""
    base, over, mock = CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + '.npz'))
    kv = load_input(base)
    kv.update({k: str(v) for k, v in over.items()})
    path = os.path.join(tmp_path, 'c.input')
    write_input(path, kv)
    cfg = bl.Config(path)
    frame = cfg.camera_frame()
    for k, v in frame.items():
        assert np.array_equal(v, gold['geo_' + k]), k
    pos, dirs, fac = cfg.camera_root()
    assert np.array_equal(pos, gold['geo_camera_pos'])
    assert np.array_equal(dirs, gold['geo_camera_dir'])
    assert np.array_equal(fac, gold['geo_momentum_factors'])


def test_camera_refined_blocks(tmp_path):
    """Children of flagged parents: order (parents in index order x 4 children), pixel positions on the finer grid."""
    kv = load_input('adaptive.input')
    path = os.path.join(tmp_path, 'a.input')
    write_input(path, kv)
    cfg = bl.Config(path)
    res, bs = 32, 8
    nb = res // bs
    parents = np.array([[v, u] for v in range(nb) for u in range(nb)], np.int32)
    flags = np.zeros(nb * nb, np.uint8)
    flags[[5, 10]] = 1
    locs, pos, dirs, fac = cfg.camera_refined(1, parents, flags)
    assert locs.tolist() == [[2, 2], [2, 3], [3, 2], [3, 3], [4, 4], [4, 5], [5, 4], [5, 5]]
    # a level-1 pixel coincides with the level-0 camera of twice the resolution
    kv2 = dict(kv, camera_resolution='64', adaptive_max_level='0')
    path2 = os.path.join(tmp_path, 'b.input')
    write_input(path2, kv2)
    pos64, dir64, fac64 = bl.Config(path2).camera_root()
    blk = 0
    v, u = locs[blk]
    for row in (0, 3, 7):
        for col in (0, 5):
            m_fine = (v * bs + row) * 64 + (u * bs + col)
            m_blk = blk * bs * bs + row * bs + col
            assert np.array_equal(pos[m_blk], pos64[m_fine]) and np.array_equal(dirs[m_blk], dir64[m_fine])
            assert fac[m_blk] == fac64[m_fine]
    # a rank that owns only some of the level's blocks builds exactly their pixels (blh_camera_blocks)
    mine = [6, 1, 3]
    bpix = bs * bs
    p2, d2, f2 = cfg.camera_blocks(1, locs[mine])
    for n, b in enumerate(mine):
        assert np.array_equal(p2[n * bpix:(n + 1) * bpix], pos[b * bpix:(b + 1) * bpix])
        assert np.array_equal(d2[n * bpix:(n + 1) * bpix], dirs[b * bpix:(b + 1) * bpix])
        assert np.array_equal(f2[n * bpix:(n + 1) * bpix], fac[b * bpix:(b + 1) * bpix])


def test_input_surface_errors(tmp_path):
    kv = load_input('simulation.input')
    def cfg_of(d):
        p = os.path.join(tmp_path, 'e.input')
        write_input(p, d)
        return bl.Config(p)
    with pytest.raises(bl.BlacklightError, match=r'Unknown key \(bogus_key\) in input file\.'):
        cfg_of(dict(kv, bogus_key='1'))
    with pytest.raises(bl.BlacklightError, match='Must have positive camera_resolution.'):
        cfg_of(dict(kv, camera_resolution='0'))
    with pytest.raises(bl.BlacklightError, match='Unknown string used for boolean value.'):
        cfg_of(dict(kv, ray_flat='yes'))
    with pytest.raises(bl.BlacklightError, match='Must have nonnegative ray_max_retries.'):
        cfg_of(dict(kv, ray_max_retries='0'))
    with pytest.raises(bl.BlacklightError, match='No image or rendering selected.'):
        cfg_of(dict(kv, image_light='false'))
    missing = dict(kv)
    del missing['camera_r']
    with pytest.raises(bl.BlacklightError, match='camera_r'):
        cfg_of(missing)
    with pytest.raises(bl.BlacklightError, match=r'Polarized transport only supports kappa in \[3.5, 5\]\.'):
        cfg_of(dict(kv, image_polarization='true', plasma_kappa_frac='0.5', plasma_kappa='6.0', plasma_w='1.0'))
    # comments and arbitrary whitespace are stripped exactly like the reference does
    p = os.path.join(tmp_path, 'w.input')
    with open(p, 'w') as f:
        for k, v in kv.items():
            f.write('  %s   =  %s   # trailing comment\n\n' % (k, v))
    assert bl.Config(p).resolution == 64


def test_glibc_math_restatements_match_host_libm(tmp_path):
    """hypot / pow used by the geodesic kernel vs this host's libm on 10^7 inputs each (bit-exact)."""
    src = os.path.join(tmp_path, 't.cpp')
    with open(src, 'w') as f:
        f.write(r'''
#include "%s/blacklight_b200/csrc/glibc_math.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
int main() {
  srand48(12345);
  long bad_pow = 0, bad_hyp = 0, bad_h3 = 0;
  for (long i = 0; i < 10000000; i++) {
    double x = std::pow(10.0, drand48() * 24 - 20) * (1 + drand48());
    double y = (i %% 3) ? -0.2 : (drand48() - 0.5) * 6;
    if (blmath::pow_glibc(x, y) != std::pow(x, y)) bad_pow++;
    double a = (drand48() - 0.5) * std::pow(10.0, drand48() * 8 - 2), b = (drand48() - 0.5) * std::pow(10.0, drand48() * 8 - 2);
    if (i %% 7 == 0) b = a * 1e-17 * drand48();
    if (i %% 11 == 0) b = 0.0;
    if (blmath::hypot_glibc(a, b) != std::hypot(a, b)) bad_hyp++;
    if (i %% 5 == 0 && blmath::hypot3_libstdcxx(a, b, 50 + 1000 * drand48()) != std::hypot(a, b, 50 + 0 * drand48())) {}
  }
  for (long i = 0; i < 1000000; i++) {
    double a = (drand48() - 0.5) * 30, b = (drand48() - 0.5) * 30, c = 50 + drand48() * 1000;
    if (blmath::hypot3_libstdcxx(a, b, c) != std::hypot(a, b, c)) bad_h3++;
  }
  std::printf("%%ld %%ld %%ld\n", bad_pow, bad_hyp, bad_h3);
  return 0;
}
''' % ROOT)
    exe = os.path.join(tmp_path, 't')
    subprocess.run(['/usr/bin/g++', '-O2', '-std=c++17', '-ffp-contract=off', '-mfma', src, '-o', exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    assert out == ['0', '0', '0'], 'mismatches (pow, hypot, hypot3): %s' % out


def test_mock_snapshot_is_stable_and_readable(tmp_path):
    """The synthetic snapshot generator is deterministic (CRC pinned; it was checked bit-for-bit against the
    reference's scripts/generate_mock_simulation.py) and its .athdf round-trips through our own reader."""
    from blacklight_b200 import mock_snapshot
    grid = mock_snapshot.make_mock(os.path.join(tmp_path, 'm.athdf'), blocks=(7, 2, 4))
    single = mock_snapshot.make_mock(None)
    assert zlib.crc32(single['prim'].tobytes()) == 0x52C06F21
    assert single['prim'].shape == (8, 1, 128, 64, 77)
    assert grid['prim'].shape == (8, 56, 32, 32, 11)
    # multi-block partition holds exactly the single-block cells
    assert np.array_equal(grid['prim'][:, 0], single['prim'][:, 0, :32, :32, :11])
    assert os.path.getsize(os.path.join(tmp_path, 'm.athdf')) > grid['prim'].nbytes


def _reader_case(tmp_path, fmt, snap, over=None):
    from harness import load_input, write_input
    kv = load_input('simulation.input')
    kv.update({'simulation_format': fmt, 'simulation_file': snap, 'camera_resolution': '8'})
    kv.update(over or {})
    path = os.path.join(str(tmp_path), fmt + '.input')
    write_input(path, kv)
    return bl.Config(path)


@pytest.mark.parametrize('location_size,variable_size', [(4, 4), (8, 8)])
def test_athenak_reader(tmp_path, location_size, variable_size, capfd):
    """simulation_format = athenak: pre-header, parameter dump, per-block records; variables
    located by name; faces rebuilt from block edges in double; eint -> pressure in float32; the adiabatic index
    taken from <mhd> gamma (reference simulation_reader.cpp:434-588,915-1131,1225-1290)."""
    from blacklight_b200 import mock_snapshot as ms
    grid = ms.add_entropy(ms.to_blocks(ms.mock_fields_cks(n=12), (2, 3, 1)))
    snap = os.path.join(str(tmp_path), 'mock.bin')
    ms.write_athenak(snap, grid, gamma_adi=1.5, time=7.25, location_size=location_size, variable_size=variable_size,
                     spin=0.5)
    kv = {'simulation_coord': 'cks', 'simulation_a': '0.5', 'plasma_model': 'code_kappa', 'simulation_kappa_name': 'r0'}
    cfg = _reader_case(tmp_path, 'athenak', snap, kv)
    g = bl.read_snapshot(cfg)
    assert (g['n_b'], g['n_k'], g['n_j'], g['n_i'], g['n_var']) == (6, 12, 4, 6, 9)
    assert g['time'] == 7.25 and g['plasma_gamma'] == 1.5
    assert np.array_equal(g['locations'], grid['locations']) and np.array_equal(g['levels'], grid['levels'])
    loc_t = np.float32 if location_size == 4 else np.float64
    for b in range(6):
        lo, hi = float(loc_t(grid['x1f'][b, 0])), float(loc_t(grid['x1f'][b, -1]))
        f = np.array([lo] + [lo + i * ((hi - lo) / 6) for i in range(1, 6)] + [hi])
        assert np.array_equal(g['x1f'][b], f)
        assert np.array_equal(g['x1v'][b], 0.5 * (f[:-1] + f[1:]))
    var_t = np.float32 if variable_size == 4 else np.float64
    prim = grid['prim'].astype(np.float64)
    eint = (prim[1] / 0.5).astype(var_t).astype(np.float32) * np.float32(0.5)
    order = [0, 2, 3, 4, None, 5, 6, 7]   # internal order rho, uu1, uu2, uu3, pgas, bb1, bb2, bb3, kappa
    for v, src in enumerate(order):
        want = eint if src is None else grid['prim'][src]
        assert np.array_equal(g['prim'][v], want), v
    assert np.array_equal(g['prim'][8], grid['kappa'])
    assert [g[k] for k in ('ind_rho', 'ind_uu1', 'ind_uu2', 'ind_uu3', 'ind_pgas', 'ind_bb1', 'ind_bb2', 'ind_bb3',
                           'ind_kappa')] == list(range(9))
    capfd.readouterr()
    # an input-file spin that differs from the dump's is warned about, in the reference's words
    bl.read_snapshot(_reader_case(tmp_path, 'athenak', snap, dict(kv, simulation_a='0.25')))
    assert 'Given spin of 0.25 does not match file value of 0.5; ignoring the latter.' in capfd.readouterr().err
    with pytest.raises(bl.BlacklightError, match='Unable to locate electron entropy values'):
        bl.read_snapshot(_reader_case(tmp_path, 'athenak', snap, dict(kv, simulation_kappa_name='s_e')))
    with open(snap, 'r+b') as f:
        f.write(b'Athena binary output version=1.0')
    with pytest.raises(bl.BlacklightError, match='Unknown AthenaK file format.'):
        bl.read_snapshot(cfg)


def test_athdf_and_harm3d_readers_return_the_generator_arrays(tmp_path):
    """blh_snapshot_read for the other two formats: the .athdf reader returns exactly the arrays the generator
    wrote (float32 coordinates widened); the harm3d reader returns one spherical Kerr-Schild block of the same
    shape with the header's time and adiabatic index."""
    from blacklight_b200 import mock_snapshot as ms
    d = str(tmp_path)
    grid = ms.make_mock(os.path.join(d, 'm.athdf'), blocks=(1, 2, 2), n_r=16, n_th=8, n_ph=8)
    want = ms.grid_view_arrays(grid)
    g = bl.read_snapshot(_reader_case(tmp_path, 'athena', os.path.join(d, 'm.athdf')))
    for k in ('levels', 'locations', 'x1f', 'x2f', 'x3f', 'x1v', 'x2v', 'x3v', 'prim'):
        assert np.array_equal(g[k], want[k]), k
    for k in ('n_b', 'n_var', 'ind_rho', 'ind_pgas', 'ind_uu1', 'ind_bb3', 'n_3_root'):
        assert g[k] == want[k], k
    fields = ms.mock_fields(n_r=16, n_th=8, n_ph=8)
    ms.write_harm3d(os.path.join(d, 'm.harm3d'), fields, gamma_adi=1.4, time=3.0)
    kv = {'simulation_coord': 'sks'}
    cfg = _reader_case(tmp_path, 'harm3d', os.path.join(d, 'm.harm3d'), kv)
    h = bl.read_snapshot(cfg)
    assert (h['n_b'], h['n_k'], h['n_j'], h['n_i']) == (1, 8, 8, 16)
    assert h['time'] == 3.0
    np.testing.assert_allclose(h['x1v'][0], np.sqrt(fields['rf'][:-1] * fields['rf'][1:]), rtol=1e-12)   # centres in ln r
    np.testing.assert_allclose(h['prim'][h['ind_rho'], 0], fields['prim'][0], rtol=1e-6)
    # the conversions run over the host threads the input file names: same bits with one thread
    h1 = bl.read_snapshot(_reader_case(tmp_path, 'harm3d', os.path.join(d, 'm.harm3d'), dict(kv, num_threads='1')))
    assert np.array_equal(h1['prim'], h['prim']) and np.array_equal(h1['x2v'], h['x2v'])
    # rho, pgas, u^phi and the field recover the generator's (its cell centres are arithmetic, the reader's
    # logarithmic, so the radial velocity differs at first order in the cell size)
    for q, name in ((1, 'ind_pgas'), (4, 'ind_uu3'), (5, 'ind_bb1'), (6, 'ind_bb2'), (7, 'ind_bb3')):
        want = fields['prim'][q].astype(np.float64)
        got = h['prim'][h[name], 0].astype(np.float64)
        assert np.max(np.abs(got - want)) <= 2e-2 * np.max(np.abs(want)), name


@pytest.mark.parametrize('fmks', [None, dict(poly_xt=0.82, poly_alpha=14.0, mks_smooth=0.5)])
def test_iharm3d_reader(tmp_path, fmks, capfd):
    """simulation_format = iharm3d: nested HDF5 groups, scalar datasets, prims (n1, n2, n3, n_prim);
    MKS coordinates converted to spherical Kerr-Schild, FMKS kept native with the (r, theta) -> (x1, x2) table; the
    vector conversion (reference simulation_geometry.cpp:95-236) must recover the generator's standard
    normal-frame velocity and coordinate-frame field, which the generator transformed the other way independently."""
    from blacklight_b200 import mock_snapshot as ms
    snap = os.path.join(str(tmp_path), 'mock.h5')
    hslope = 0.3 if fmks else 0.7
    w = ms.write_iharm3d(snap, n_r=24, n_th=16, n_ph=8, gamma_adi=1.5, time=4.5, hslope=hslope, fmks=fmks)
    kv = {'simulation_coord': 'fmks' if fmks else 'sks', 'simulation_a': '0.0'}
    from harness import load_input
    base = load_input('simulation.input')
    cfg = _reader_case(tmp_path, 'iharm3d', snap, kv)
    g = bl.read_snapshot(cfg)
    assert (g['n_b'], g['n_k'], g['n_j'], g['n_i'], g['n_var']) == (1, 8, 16, 24, 8)
    assert g['time'] == 4.5
    assert g['plasma_gamma'] == (float(base['plasma_gamma']) if 'plasma_gamma' in base else 1.5)
    lr = 0.5 * (w['lrf'][:-1] + w['lrf'][1:])
    x2 = 0.5 * (w['x2f'][:-1] + w['x2f'][1:])
    if fmks:
        np.testing.assert_allclose(g['x1v'][0], lr, rtol=1e-14)          # native coordinates
        np.testing.assert_allclose(g['x2v'][0], x2, rtol=1e-14)
        # the table inverts theta(x1, x2) on a 2048 x 2048 lattice of (r, theta).  (Not to the bisection tolerance of
        # 1e-8: the reference's loop, restated as is, moves x2 to the midpoint of the next bracket before testing
        # the previous midpoint's theta, simulation_geometry.cpp:381-395.)
        m = g['sks_map']
        assert m.shape == (2, 2048, 2048)
        r_in = np.exp(w['lrf'][0])
        assert g['sks_map_r_in'] == r_in and abs(g['sks_map_dtheta'] - np.pi / 2047) < 1e-15
        jj, ii = np.meshgrid(np.arange(0, 2048, 89), np.arange(0, 2048, 97), indexing='ij')
        np.testing.assert_allclose(m[0, jj, ii], np.log(r_in + ii * g['sks_map_dr']), rtol=1e-14)
        th = ms.fmks_theta(m[0, jj, ii], m[1, jj, ii], hslope, r_in, fmks['poly_xt'], fmks['poly_alpha'], fmks['mks_smooth'])
        assert np.max(np.abs(th - jj * g['sks_map_dtheta'])) < 1e-4
        np.testing.assert_allclose(g['simulation_bounds'], [r_in, np.exp(w['lrf'][-1]), 0.0, np.pi, 0.0, 2.0 * np.pi], atol=1e-12)
    else:
        np.testing.assert_allclose(g['x1v'][0], np.exp(lr), rtol=1e-14)
        np.testing.assert_allclose(g['x2v'][0], np.pi * x2 + (1.0 - hslope) / 2.0 * np.sin(2.0 * np.pi * x2), rtol=1e-14)
    # host threads only split independent cells / table rows: one thread gives the same bits
    g1 = bl.read_snapshot(_reader_case(tmp_path, 'iharm3d', snap, dict(kv, num_threads='1')))
    assert np.array_equal(g1['prim'], g['prim'])
    if fmks:
        assert np.array_equal(g1['sks_map'], g['sks_map'])
    ph = 0.5 * (w['phf'][:-1] + w['phf'][1:])
    want = ms.mock_fields_at(w['r'][None], w['th'][None], ph[:, None, None])
    gm1 = np.float32(g['plasma_gamma'] - 1.0)
    names = ('ind_rho', 'ind_pgas', 'ind_uu1', 'ind_uu2', 'ind_uu3', 'ind_bb1', 'ind_bb2', 'ind_bb3')
    for q, name in enumerate(names):
        got = g['prim'][g[name], 0].astype(np.float64)
        ref = want[q] if q != 1 else (want[1] / 0.5).astype(np.float32).astype(np.float64) * float(gm1)
        scale = np.max(np.abs(ref))
        tol = 3e-7 if not fmks or q < 2 else 2e-5     # the FMKS generator differentiates theta(x1, x2) numerically
        assert np.max(np.abs(got - ref)) <= tol * scale, (name, np.max(np.abs(got - ref)) / scale)


@pytest.mark.parametrize('fmt', ['athena', 'athenak', 'iharm3d', 'harm3d'])
def test_time_series_reread_equals_fresh_read(fmt, tmp_path):
    """Second file of a time series read on top of the first (blh_snapshot_reread: the layout, variable positions and
    adiabatic index of the first file are kept, as simulation_reader.cpp does after its first call) gives exactly the
    arrays of reading that file alone, with its own time."""
    from blacklight_b200 import mock_snapshot as ms
    d = str(tmp_path)
    files, kv = [], {}
    for n, amp in enumerate((0.1, 0.3)):
        f = os.path.join(d, 'mock.%d' % n)
        if fmt == 'athena':
            ms.write_athdf(f, ms.to_blocks(ms.mock_fields(n_r=16, n_th=8, n_ph=8, pert_amp=amp, pert_n_ph=1), (2, 1, 2)), time=5.0 * n)
        elif fmt == 'athenak':
            grid = ms.to_blocks(ms.mock_fields_cks(n=8), (2, 1, 1))
            grid['prim'] = grid['prim'] * np.float32(1.0 + amp)
            ms.write_athenak(f, grid, gamma_adi=1.5, time=5.0 * n, spin=0.5)
            kv = {'simulation_coord': 'cks', 'simulation_a': '0.5'}
        elif fmt == 'iharm3d':
            ms.write_iharm3d(f, n_r=16, n_th=8, n_ph=8, gamma_adi=1.5, time=5.0 * n, hslope=0.7, pert_amp=amp, pert_n_ph=1)
            kv = {'simulation_coord': 'sks'}
        else:
            ms.write_harm3d(f, ms.mock_fields(n_r=16, n_th=8, n_ph=8, pert_amp=amp, pert_n_ph=1), gamma_adi=1.5, time=5.0 * n)
            kv = {'simulation_coord': 'sks'}
        files.append(f)
    cfg = _reader_case(tmp_path, fmt, files[0], kv)
    first, fresh = bl.read_snapshot(cfg, files[0]), bl.read_snapshot(cfg, files[1])
    again = bl.read_snapshot(cfg, files[0], then=files[1])
    assert again['time'] == fresh['time'] == 5.0 and first['time'] == 0.0
    assert not np.array_equal(first['prim'], fresh['prim'])
    for k in ('prim', 'x1f', 'x2f', 'x3f', 'x1v', 'x2v', 'x3v', 'levels', 'locations'):
        assert np.array_equal(again[k], fresh[k]), k


def test_parallel_crc32_equals_zlib():
    """The npz writer's CRC-32 (chunks on all host threads, combined in GF(2)) against zlib, across the serial / parallel
    threshold and with ragged tails."""
    import zlib
    lib = bl.load_library()
    rng = np.random.default_rng(11)
    for n in (0, 1, 7, 4096, (8 << 20) - 1, (8 << 20) + 5, 37 * (1 << 20) + 123):
        buf = rng.integers(0, 256, n, dtype=np.uint8)
        assert lib.blh_crc32(buf.ctypes.data, n) == (zlib.crc32(buf.tobytes()) & 0xffffffff), n


def test_camera_rows_equal_the_rows_of_the_full_camera(tmp_path):
    """blh_camera_rows (a device's share of the frame in the multi-GPU driver) is bit for bit the same rows of
    blh_camera_root, plane and pinhole cameras."""
    for over in ({'camera_resolution': 12}, {'camera_resolution': 10, 'camera_type': 'pinhole', 'camera_th': '30.0'}):
        case = Case(tmp_path / ('c%d' % len(over)), 'formula.input', over)
        cfg = case.config()
        res = cfg.resolution
        pos, dirs, fac = cfg.camera_root()
        rows = np.arange(1, res, 3)
        p2, d2, f2 = cfg.camera_rows(rows)
        idx = (rows[:, None] * res + np.arange(res)[None, :]).ravel()
        assert np.array_equal(p2, pos[idx]) and np.array_equal(d2, dirs[idx]) and np.array_equal(f2, fac[idx])
        with pytest.raises(bl.BlacklightError):
            cfg.camera_rows([res])


def test_open_snapshots_do_not_share_reader_state(tmp_path):
    """Two snapshot handles open at the same time, and a reread issued from another thread: what the first file of a
    series fixed (AthenaK variable positions and record size, harm3d / iharm3d coordinate parameters and modified x2
    centres) belongs to the handle, so interleaved reads of other files cannot change what a reread returns."""
    import ctypes
    import threading
    from blacklight_b200 import mock_snapshot as ms
    lib = bl.load_library()
    d = str(tmp_path)
    series = {}
    for fmt in ('athenak', 'iharm3d', 'harm3d'):
        files = []
        for n, amp in enumerate((0.1, 0.3)):
            f = os.path.join(d, '%s.%d' % (fmt, n))
            if fmt == 'athenak':
                grid = ms.to_blocks(ms.mock_fields_cks(n=8), (2, 1, 1))
                grid['prim'] = grid['prim'] * np.float32(1.0 + amp)
                ms.write_athenak(f, grid, gamma_adi=1.5, time=5.0 * n, spin=0.5)
                kv = {'simulation_coord': 'cks', 'simulation_a': '0.5'}
            elif fmt == 'iharm3d':
                ms.write_iharm3d(f, n_r=16, n_th=8, n_ph=8, gamma_adi=1.5, time=5.0 * n, hslope=0.7, pert_amp=amp, pert_n_ph=1)
                kv = {'simulation_coord': 'sks'}
            else:
                ms.write_harm3d(f, ms.mock_fields(n_r=12, n_th=6, n_ph=4, pert_amp=amp, pert_n_ph=1), gamma_adi=1.5, time=5.0 * n)
                kv = {'simulation_coord': 'sks'}
            files.append(f)
        sub = tmp_path / fmt
        sub.mkdir()
        series[fmt] = (_reader_case(sub, fmt, files[0], kv), files)

    def prim_of(handle):
        v, t, g = bl.GridView(), ctypes.c_double(), ctypes.c_double()
        lib.blh_snapshot_view(handle, ctypes.byref(v), ctypes.byref(t), ctypes.byref(g))
        n = v.n_var * v.n_b * v.n_k * v.n_j * v.n_i
        return np.frombuffer((ctypes.c_char * (4 * n)).from_address(v.prim), dtype=np.float32).copy(), t.value

    handles = {}
    for fmt, (cfg, files) in series.items():      # all three first files open at once
        h = ctypes.c_void_p()
        assert lib.blh_snapshot_read(cfg._h, os.fsencode(files[0]), ctypes.byref(h)) == 0, lib.blh_last_error()
        handles[fmt] = h
    results = {}

    def reread(fmt):                               # on a fresh thread, after the other formats were read on the main one
        rc = lib.blh_snapshot_reread(handles[fmt], os.fsencode(series[fmt][1][1]))
        results[fmt] = (rc,) + (prim_of(handles[fmt]) if rc == 0 else (None, None))

    for fmt in ('athenak', 'harm3d', 'iharm3d'):
        t = threading.Thread(target=reread, args=(fmt,))
        t.start()
        t.join()
    for fmt, (cfg, files) in series.items():
        rc, prim, time = results[fmt]
        assert rc == 0, fmt
        fresh = bl.read_snapshot(cfg, files[1])
        assert time == fresh['time'] == 5.0
        assert np.array_equal(prim, fresh['prim'].ravel()), fmt
        lib.blh_snapshot_free(handles[fmt])


def test_npz_writer_and_athdf_reader_through_driver_without_gpu(tmp_path):
    """blh_run_input_file must fail loudly without a GPU, after parsing the file."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    case = Case(tmp_path, 'simulation.input', {'camera_resolution': 8})
    with pytest.raises(bl.BlacklightError, match='CUDA'):
        case.run_gpu_file()


def _shard_worker(rank, world, port, res, q):
    import torch
    import torch.distributed as dist
    from blacklight_b200.sharding import assemble, shard_rows
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    rows, idx = shard_rows(res, rank, world)
    local = torch.from_numpy(np.stack([np.sin(idx * 0.37), idx.astype(np.float64)]))  # stand-in for (Q, n) image
    gathered = [torch.empty_like(local) for _ in range(world)] if rank == 0 else None
    dist.gather(local, gathered, dst=0)
    if rank == 0:
        full = assemble([g.numpy() for g in gathered], res, world)
        m = np.arange(res * res)
        ok = np.array_equal(full[1], m.astype(np.float64)) and np.array_equal(full[0], np.sin(m * 0.37))
        q.put(bool(ok))
    dist.destroy_process_group()


def test_row_sharding_and_gather_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, 12, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def _gather_rows_worker(rank, world, port, res, q):
    import torch
    import torch.distributed as dist
    from blacklight_b200.sharding import gather_rows, shard_rows
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    ok = True
    for quantities in (3, 0):       # Q = 0: a rendering-only frame has no image quantities to exchange
        _, idx = shard_rows(res, rank, world)
        mine = torch.from_numpy((idx[None, :] * (np.arange(quantities)[:, None] + 1.0)).reshape(quantities, len(idx)))
        parts = [torch.empty((quantities, len(shard_rows(res, r, world)[1])), dtype=torch.float64) for r in range(world)] if rank == 0 else None
        full = torch.zeros((quantities, res, res), dtype=torch.float64) if rank == 0 else None
        gather_rows(mine, parts, full, res, rank, world, dist)
        if rank == 0:
            m = np.arange(res * res, dtype=np.float64)
            ok = ok and all(np.array_equal(full[k].numpy().ravel(), m * (k + 1.0)) for k in range(quantities))
    dist.barrier()                  # both ranks are still in step after the empty frame
    if rank == 0:
        q.put(bool(ok))
    dist.destroy_process_group()


@pytest.mark.parametrize('res', [12, 16])
def test_gather_rows_world_size_2(res):
    """bench.py's final exchange (blacklight_b200.sharding.gather_rows): rank 0 ends up with the frame in the reference's
    pixel order, a frame without image quantities exchanges nothing, and neither rank is left behind."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000 + res
    procs = [ctx.Process(target=_gather_rows_worker, args=(r, 2, port, res, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


class _FakeConfig:
    """Stands in for blacklight_b200.Config in the sharded-adaptive host logic test: an 8x8 image of 2x2 blocks."""
    resolution, block_size = 8, 2


class _FakeContext:
    """"Image" of a ray = a tag of its pixel (level 0: the raster index, so assembly order is checkable); refine blocks
    whose first pixel's tag is a multiple of 3."""
    level0_block_major = True

    def trace_level_pixels(self, level, rows=None, blocks=None):
        locs = np.asarray(blocks, np.int32).reshape(-1, 2)
        if level == 0:
            i, j = np.divmod(np.arange(4), 2)
            self.fac = ((locs[:, 0, None] * 2 + i[None, :]) * 8 + locs[:, 1, None] * 2 + j[None, :]).astype(np.float64).ravel()
        else:
            self.fac = 1000.0 * level + (locs[:, 0].repeat(4) * 64 + locs[:, 1].repeat(4)) * 4.0 + np.tile(np.arange(4.0), len(locs))
        return {'num_bad_geodesics': 0}

    def radiate_level(self, level, num_render=0):
        return self.fac[None, :].copy(), None, {'num_samples': len(self.fac)}

    def refine_level(self, level, locs):
        f = (np.round(self.fac.reshape(len(locs), -1)[:, 0]) % 3 == 0).astype(np.uint8)
        return f, int(f.sum())


def _adaptive_worker(rank, world, port, q):
    import torch.distributed as dist
    from blacklight_b200 import multigpu
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    out = multigpu.run_distributed(multigpu.adaptive_worker(_FakeConfig(), _FakeContext(), rank, world, 2), rank, world)
    if rank == 0:
        q.put([(L['locs'].tolist(), L['image'].tolist()) for L in out])
    dist.destroy_process_group()


def test_adaptive_block_sharding_world_size_2_matches_single_rank():
    """Host logic of blacklight_b200/multigpu.py over a gloo group of 2: flag all-gather, identical child lists,
    image gather and assembly -- against the same generator run as a single rank in this process."""
    import torch.multiprocessing as mp
    from blacklight_b200 import multigpu
    single = multigpu.run_local([multigpu.adaptive_worker(_FakeConfig(), _FakeContext(), 0, 1, 2)])[0]
    assert len(single) == 3 and np.array_equal(single[0]['image'][0], np.arange(64.0))
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_adaptive_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert len(got) == len(single)
    for (locs, image), L in zip(got, single):
        assert np.array_equal(np.array(locs).reshape(-1, 2), L['locs'])
        assert np.array_equal(np.array(image), L['image'])


def test_branch_free_math_against_libm_and_mpmath(tmp_path):
    """blacklight_b200/csrc/bf_math.cuh compiled for the host: exp_bf / log_bf against libm, bessel_k01 against
    40-digit values (the radiation kernels use these in their per-frequency coefficient code; the image tolerance
    is 1e-6, the functions are good to a few ulp)."""
    import subprocess
    mp = pytest.importorskip('mpmath')
    src = os.path.join(tmp_path, 'bf_check.cpp')
    with open(src, 'w') as f:
        f.write('#include <cstdio>\n#include <cmath>\n#include "%s"\n' % os.path.join(ROOT, 'blacklight_b200', 'csrc', 'bf_math.cuh'))
        f.write('int main() { double x; while (scanf("%lf", &x) == 1) { double k0, k1; bfm::bessel_k01(x > 0 ? x : 1.0, k0, k1);\n'
                '  printf("%.17e %.17e %.17e %.17e %.17e\\n", x, bfm::exp_bf(x), bfm::log_bf(x > 0 ? x : 1.0), k0, k1); } }\n')
    exe = os.path.join(tmp_path, 'bf_check')
    subprocess.run(['g++', '-O2', '-ffp-contract=off', '-o', exe, src, '-lm'], check=True)
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(-700.0, 700.0, 20000), np.exp(rng.uniform(np.log(1e-3), np.log(150.0), 3000)), [2.0, 1.0, 100.0]])
    out = subprocess.run([exe], input='\n'.join('%.17e' % x for x in xs), capture_output=True, text=True, check=True).stdout
    v = np.array([[float(t) for t in line.split()] for line in out.strip().splitlines()])
    x = v[:, 0]
    assert np.max(np.abs(v[:, 1] - np.exp(x)) / np.exp(x)) < 1e-15
    pos = x > 0
    ref_log = np.log(x[pos])
    assert np.max(np.abs(v[pos, 2] - ref_log) / np.maximum(np.abs(ref_log), 1e-3)) < 2e-15
    mp.mp.dps = 40
    sel = np.where((x > 1e-3) & (x < 150.0))[0][-300:]
    for i in sel:
        k0, k1 = float(mp.besselk(0, mp.mpf(x[i]))), float(mp.besselk(1, mp.mpf(x[i])))
        assert abs(v[i, 3] - k0) <= 4e-15 * k0 and abs(v[i, 4] - k1) <= 4e-15 * k1, x[i]
