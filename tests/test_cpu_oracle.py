"""Pin the plain-C restatement (oracle/blacklight_oracle.c) against the golden fixtures produced by the
unmodified reference: sample counts, flags and every stored sample bit for bit, cell indices exactly,
images to rounding.  Runs on CPU only (the restatement is test infrastructure)."""
import os
import zlib

import numpy as np
import pytest

import blacklight_b200 as bl
from harness import ROOT, load_input, write_input
from golden.make_golden import CASES

from blacklight_b200 import mock_snapshot
import oracle_lib

GOLDEN = os.path.join(ROOT, 'tests', 'golden')

if not os.path.exists(oracle_lib.LIB):
    pytest.skip('oracle restatement not built (run __graft_entry__.build())', allow_module_level=True)


def setup(name, tmp_path):
    base, over, mock = CASES[name]
    kv = load_input(base)
    kv.update({k: str(v) for k, v in over.items()})
    path = os.path.join(tmp_path, 'o.input')
    write_input(path, kv)
    cfg = bl.Config(path)   # host layer only: camera arrays (checked bit-exact in test_cpu_host.py)
    gold = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    return kv, cfg, gold, mock


def check_samples(s, gold):
    assert np.array_equal(s['num'], gold['sample_num'])
    assert np.array_equal(s['flags'], gold['sample_flags'])
    assert s['steps'] == int(gold['geodesic_num_steps'])
    mask = np.arange(s['cap'])[None, :] < s['num'][:, None]
    crc = zlib.crc32(s['pos'][mask].tobytes() + s['dir'][mask].tobytes() + s['len'][mask].tobytes())
    assert crc == int(gold['samples_crc'])
    return mask


@pytest.mark.parametrize('name', ['formula_16', 'formula_pinhole_pole_12', 'formula_rk4_max_steps_12', 'formula_photon_12',
                                  'formula_additive_12'])
def test_oracle_formula(name, tmp_path):
    kv, cfg, gold, _ = setup(name, tmp_path)
    pos, dirs, fac = cfg.camera_root()
    s = oracle_lib.trace(kv, float(kv['formula_spin']), pos, dirs)
    check_samples(s, gold)
    image = oracle_lib.formula_image(kv, s, fac, gold['frequency'])
    res = cfg.resolution
    ref = gold['I_nu']
    got = image[0].reshape(res, res)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.max(np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-300)) < 1e-12


@pytest.mark.parametrize('name', ['simulation_32', 'simulation_nearest_24', 'simulation_blocks_24', 'simulation_kerr_24',
                                  'simulation_rk4_16', 'simulation_rk2_kerr_16', 'simulation_amr_16'])
def test_oracle_simulation(name, tmp_path):
    kv, cfg, gold, mock = setup(name, tmp_path)
    mock = dict(mock or {})
    grid = mock_snapshot.grid_view_arrays(mock_snapshot.make_mock(None, tuple(mock.pop('blocks', (1, 1, 1))), **mock))
    pos, dirs, fac = cfg.camera_root()
    s = oracle_lib.trace(kv, float(kv['simulation_a']), pos, dirs)
    mask = check_samples(s, gold)
    image, inds = oracle_lib.simulation_image(kv, s, fac, grid)
    valid = mask & (inds[..., 0] >= 0)
    assert int(valid.sum()) == int(gold['valid_count'])
    assert zlib.crc32(np.ascontiguousarray(inds[valid]).tobytes()) == int(gold['inds_crc'])
    res = cfg.resolution
    ref, got = gold['I_nu'], image.reshape(res, res)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    scale = np.maximum(np.abs(ref[ok]), 1e-12 * np.nanmax(np.abs(ref)))
    assert np.max(np.abs(got[ok] - ref[ok]) / scale) < 1e-10


def test_format_fixtures_match_their_generator(tmp_path):
    """tests/golden/formats_*.npz were produced by the unmodified reference from deterministic mock dumps
    (AthenaK, iharm3d MKS / FMKS, harm3d); the dumps the GPU parity test rebuilds must be those same bytes."""
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden_formats import FORMAT_CASES, dump_crc, write_case
    for name in FORMAT_CASES:
        gold = np.load(os.path.join(GOLDEN, 'formats_%s.npz' % name))
        d = os.path.join(str(tmp_path), name)
        os.makedirs(d)
        path, _ = write_case(name, d)
        assert dump_crc(path) == int(gold['dump_crc']), name
        assert gold['defined'].shape == gold['I_nu'].shape and gold['defined'].mean() > 0.9
        assert np.nanmax(gold['I_nu']) > 0.0


def test_fmks_table_and_zone_lookup_against_reference_checkpoint(tmp_path):
    """FMKS host logic without a GPU: the (r, theta) -> (x1, x2) table our iharm3d reader builds, pushed through a numpy
    restatement of the reference's scaled zone lookup (simulation_sampling.cpp:397-418), reproduces the cell indices
    and fractions of the unmodified reference's sampling checkpoint exactly -- on the reference's own geodesics."""
    import subprocess
    import sys
    from harness import REF_BIN
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    sys.path.insert(0, GOLDEN)
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import refio
    from make_golden_formats import write_case
    d = str(tmp_path)
    path, _ = write_case('iharm3d_fmks_nearest_16', d)
    lines = [ln for ln in open(path).read().splitlines() if not ln.startswith(('checkpoint_', 'simulation_interp'))]
    lines += ['simulation_interp = true', 'checkpoint_sample_save = true', 'checkpoint_sample_load = false',
              'checkpoint_sample_file = %s/samp.ckpt' % d, 'checkpoint_geodesic_save = true',
              'checkpoint_geodesic_load = false', 'checkpoint_geodesic_file = %s/geo.ckpt' % d]
    with open(path, 'w') as f:
        f.write('\n'.join(lines) + '\n')
    proc = subprocess.run([REF_BIN, path], cwd=d, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    s = refio.read_sample_checkpoint(os.path.join(d, 'samp.ckpt'), interp=True)
    g = refio.read_geodesic_checkpoint(os.path.join(d, 'geo.ckpt'))
    G = bl.read_snapshot(bl.Config(path))
    x, y, z = (g['sample_pos'][..., c] for c in (1, 2, 3))
    with np.errstate(all='ignore'):
        r = np.sqrt(x * x + y * y + z * z)          # a = 0: the Kerr-Schild radius
        th = np.arccos(z / r)
    S = s['sample_nan'].shape[1]
    valid = (np.arange(S)[None, :] < g['sample_num'][:, None]) & (s['sample_nan'] == 0) & (r <= 50.0)
    x1, x2, m = r[valid], th[valid], G['sks_map']
    f_i, i_ind = np.modf((x1 - G['sks_map_r_in']) / G['sks_map_dr'])
    f_j, j_ind = np.modf(x2 / G['sks_map_dtheta'])
    i, j = i_ind.astype(int), j_ind.astype(int)
    nat_x1 = (1.0 - f_i) * m[0, j, i] + f_i * m[0, j, i + 1]
    nat_x2 = (1.0 - f_j) * m[1, j + 1, i] + f_j * m[1, j + 1, i]
    x1f, x2f = G['x1f'][0], G['x2f'][0]
    frac_i, zone_i = np.modf((nat_x1 - x1f[0]) / (x1f[1] - x1f[0]))
    frac_j, zone_j = np.modf(nat_x2 / (x2f[1] - x2f[0]))
    inds, fracs = s['sample_inds'][valid], s['sample_fracs'][valid]
    assert valid.sum() > 10000
    assert np.array_equal(inds[:, 3], zone_i.astype(int)) and np.array_equal(inds[:, 2], zone_j.astype(int))
    assert np.array_equal(fracs[:, 2], frac_i) and np.array_equal(fracs[:, 1], frac_j)


def test_oracle_formula_auxiliary_images(tmp_path):
    """Auxiliary images of the formula model (time, length, lambda, emission, tau, crossings; unpolarized.cpp:60-185)
    restated in C against the unmodified reference's fixture."""
    kv, cfg, gold, _ = setup('formula_aux_12', tmp_path)
    pos, dirs, fac = cfg.camera_root()
    s = oracle_lib.trace(kv, float(kv['formula_spin']), pos, dirs)
    check_samples(s, gold)
    aux = oracle_lib.formula_aux(kv, s, fac, gold['frequency'], cfg.camera_frame()['cam_x'])
    res = cfg.resolution
    for name in ('time', 'length', 'lambda', 'emission', 'tau', 'crossings'):
        ref = gold[name]
        got = (aux[name][0] if aux[name].ndim == 2 else aux[name]).reshape(res, res)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), name
        ok = ~np.isnan(ref)
        scale = np.maximum(np.maximum(np.abs(ref[ok]), 1e-12 * np.nanmax(np.abs(ref))), 1e-300)   # tau is 0 here: no absorption
        assert np.max(np.abs(got[ok] - ref[ok]) / scale) < 1e-11, name
    assert np.nanmax(gold['crossings']) >= 1 and np.nanmin(gold['time']) < 0.0


def test_oracle_true_color_frequencies(tmp_path):
    """BASELINE config 5 (true-colour): ten unpolarized frequencies on one set of geodesics; the restatement, run once
    per frequency of the reference's own frequency list, against the reference's multi-frequency fixture."""
    kv, cfg, gold, mock = setup('true_color_16', tmp_path)
    grid = mock_snapshot.grid_view_arrays(mock_snapshot.make_mock(None))
    pos, dirs, fac = cfg.camera_root()
    s = oracle_lib.trace(kv, float(kv['simulation_a']), pos, dirs)
    check_samples(s, gold)
    res = cfg.resolution
    freqs = np.atleast_1d(gold['frequency'])
    assert len(freqs) == 10 and gold['I_nu'].shape == (10, res, res)
    for l, nu in enumerate(freqs):
        image, _ = oracle_lib.simulation_image(dict(kv, image_frequency=repr(float(nu))), s, fac, grid, want_inds=False)
        ref, got = gold['I_nu'][l], image.reshape(res, res)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), l
        ok = ~np.isnan(ref)
        scale = np.maximum(np.abs(ref[ok]), 1e-12 * np.nanmax(np.abs(ref)))
        assert np.max(np.abs(got[ok] - ref[ok]) / scale) < 1e-10, l


def test_oracle_render(tmp_path):
    """False-colour rendering (rendering.cpp:25-179; flat-space rays, no light image): fills with their optical-depth
    law along the proper length, threshold crossings alpha-blended, colours through the sRGB -> XYZ conversion of
    the input reader -- restated in C against the reference's fixture."""
    kv, cfg, gold, mock = setup('render_32', tmp_path)
    grid = mock_snapshot.grid_view_arrays(mock_snapshot.make_mock(None))
    pos, dirs, fac = cfg.camera_root()
    s = oracle_lib.trace(kv, float(kv['simulation_a']), pos, dirs)
    check_samples(s, gold)
    _, _, rendering = oracle_lib.simulation_image(kv, s, fac, grid, want_inds=False, render=True)
    ref = gold['rendering']
    got = rendering.reshape(ref.shape)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    scale = np.maximum(np.abs(ref[ok]), 1e-12 * np.nanmax(np.abs(ref)))
    assert np.max(np.abs(got[ok] - ref[ok]) / scale) < 1e-9
    assert np.nanmax(ref) > 0.0


def test_oracle_simulation_auxiliary_images(tmp_path):
    """All 27 auxiliary images of the simulation model -- time, length, lambda, emission, tau, crossings and the
    lambda- / emission-averaged and tau-integrated cell values (rho, n_e, p_gas, Theta_e, B, sigma, 1/beta;
    unpolarized.cpp:60-200, simulation_coefficients.cpp:376-387) -- restated in C against the reference's fixture."""
    kv, cfg, gold, mock = setup('simulation_aux_16', tmp_path)
    grid = mock_snapshot.grid_view_arrays(mock_snapshot.make_mock(None))
    pos, dirs, fac = cfg.camera_root()
    s = oracle_lib.trace(kv, float(kv['simulation_a']), pos, dirs)
    check_samples(s, gold)
    image, aux = oracle_lib.simulation_image(kv, s, fac, grid, want_inds=False, camera_x=cfg.camera_frame()['cam_x'])
    res = cfg.resolution
    worst = {}
    for name in ['I_nu'] + oracle_lib.AUX_NAMES:
        ref = gold[name]
        got = (image if name == 'I_nu' else aux[name]).reshape(res, res)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), name
        ok = ~np.isnan(ref)
        scale = np.maximum(np.maximum(np.abs(ref[ok]), 1e-12 * np.nanmax(np.abs(ref))), 1e-300)
        worst[name] = float(np.max(np.abs(got[ok] - ref[ok]) / scale))
        assert worst[name] < 1e-9, (name, worst[name])
    assert np.nanmax(gold['tau']) > 0.0 and np.nanmax(gold['tau_int_Theta_e']) > 0.0


@pytest.mark.parametrize('name', ['iharm3d_mks_16', 'harm3d_16', 'athenak_16'])
def test_readers_through_the_restatement_against_reference_fixtures(name, tmp_path):
    """Host readers without a GPU: the arrays our iharm3d / harm3d / AthenaK readers hand to bl_upload_grid (coordinates
    converted to spherical Kerr-Schild and primitives to the standard frames for the Harm formats; Cartesian Kerr-Schild
    blocks rebuilt from their edges for AthenaK), rendered by the plain-C restatement of the
    reference's sampling + thermal transfer, against the unmodified reference's image of the same dump."""
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden_formats import write_case
    gold = np.load(os.path.join(GOLDEN, 'formats_%s.npz' % name))
    path, _ = write_case(name, str(tmp_path))
    with open(path) as f:
        kv = bl.parse_input_text(f.read())
    cfg = bl.Config(path)
    g = bl.read_snapshot(cfg)
    order = ('ind_rho', 'ind_pgas', 'ind_uu1', 'ind_uu2', 'ind_uu3', 'ind_bb1', 'ind_bb2', 'ind_bb3')   # the restatement's
    grid = dict(g, prim=np.ascontiguousarray(np.stack([g['prim'][g[k]] for k in order]), np.float32))
    pos, dirs, fac = cfg.camera_root()
    s = oracle_lib.trace(kv, float(kv['simulation_a']), pos, dirs)
    image, _ = oracle_lib.simulation_image(kv, s, fac, grid, want_inds=False)
    res = cfg.resolution
    ref, got = gold['I_nu'], image.reshape(res, res)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    scale = np.maximum(np.abs(ref[ok]), 1e-12 * np.nanmax(np.abs(ref)))
    assert np.max(np.abs(got[ok] - ref[ok]) / scale) < 1e-10


@pytest.mark.parametrize('fixture,over', [
    ('cpu_simulation_power_law_16', {'plasma_power_frac': '0.4', 'plasma_p': '3.0', 'plasma_gamma_min': '4.0',
                                     'plasma_gamma_max': '1000.0'}),
    ('cpu_simulation_mixed_electrons_16', {'plasma_power_frac': '0.2', 'plasma_p': '3.0', 'plasma_gamma_min': '4.0',
                                           'plasma_gamma_max': '1000.0', 'plasma_kappa_frac': '0.5', 'plasma_kappa': '4.0',
                                           'plasma_w': '1.0'}),
    ('cpu_simulation_energy_temperature_16', {'plasma_use_p': 'false', 'plasma_gamma': '1.5',
                                              'plasma_gamma_i': '1.6666666666666667',
                                              'plasma_gamma_e': '1.3333333333333333'}),
])
def test_oracle_nonthermal_electrons(fixture, over, tmp_path):
    """Thermal + power-law (+ kappa) electrons: the restatement's constants (tgamma forms and the truncated
    hypergeometric series, simulation_coefficients.cpp:53-105,740-773) and per-sample emissivities / absorptivities
    (:559-585, :608-653, including the kappa absorptivity that an unpolarized run of the reference zeroes), and the
    electron temperature from the internal energies when plasma_use_p = false (:340-346), against the
    unmodified reference's images (tests/golden/cpu_*.npz, made with the harness's Case.run_reference)."""
    kv = load_input('simulation.input')
    kv.update(dict(over, camera_resolution='16'))
    path = os.path.join(tmp_path, 'o.input')
    write_input(path, kv)
    cfg = bl.Config(path)
    grid = mock_snapshot.grid_view_arrays(mock_snapshot.make_mock(None))
    pos, dirs, fac = cfg.camera_root()
    s = oracle_lib.trace(kv, float(kv['simulation_a']), pos, dirs)
    image, _ = oracle_lib.simulation_image(kv, s, fac, grid, want_inds=False)
    ref = np.load(os.path.join(GOLDEN, fixture + '.npz'))['I_nu']
    got = image.reshape(16, 16)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    scale = np.maximum(np.abs(ref[ok]), 1e-12 * np.nanmax(np.abs(ref)))
    assert np.max(np.abs(got[ok] - ref[ok]) / scale) < 1e-10
    # and it differs from the purely thermal image, i.e. the non-thermal terms are exercised
    thermal = np.load(os.path.join(GOLDEN, 'simulation_32.npz'))['I_nu']
    assert abs(np.nanmax(ref) / np.nanmax(thermal) - 1.0) > 0.01


@pytest.mark.parametrize('fixture,over', [
    ('cpu_simulation_code_kappa_16', {}),
    ('cpu_simulation_code_kappa_nearest_16', {'simulation_interp': 'false'}),
])
def test_oracle_code_kappa(fixture, over, tmp_path):
    """plasma_model = code_kappa: the electron-entropy variable sampled like the other primitives (nearest cell, or
    trilinear with its own non-positive fallback and float storage, simulation_sampling.cpp:726,811-833) and the
    electron temperature from it (simulation_coefficients.cpp:351-358), against the unmodified reference's images."""
    kv = load_input('simulation.input')
    kv.update(dict(over, plasma_model='code_kappa', simulation_kappa_name='r0', camera_resolution='16'))
    path = os.path.join(tmp_path, 'o.input')
    write_input(path, kv)
    cfg = bl.Config(path)
    g = mock_snapshot.grid_view_arrays(mock_snapshot.make_mock(None, entropy=True))
    order = ('ind_rho', 'ind_pgas', 'ind_uu1', 'ind_uu2', 'ind_uu3', 'ind_bb1', 'ind_bb2', 'ind_bb3', 'ind_kappa')
    grid = dict(g, prim=np.ascontiguousarray(np.stack([g['prim'][g[k]] for k in order]), np.float32))
    pos, dirs, fac = cfg.camera_root()
    s = oracle_lib.trace(kv, float(kv['simulation_a']), pos, dirs)
    image, _ = oracle_lib.simulation_image(kv, s, fac, grid, want_inds=False)
    ref = np.load(os.path.join(GOLDEN, fixture + '.npz'))['I_nu']
    got = image.reshape(16, 16)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    scale = np.maximum(np.abs(ref[ok]), 1e-12 * np.nanmax(np.abs(ref)))
    assert np.max(np.abs(got[ok] - ref[ok]) / scale) < 1e-10
    thermal = np.load(os.path.join(GOLDEN, 'simulation_32.npz'))['I_nu']
    assert abs(np.nanmax(ref) / np.nanmax(thermal) - 1.0) > 0.01


@pytest.mark.parametrize('name', ['cuts_a', 'cuts_b', 'cuts_c', 'cuts_d', 'cuts_e'])
def test_oracle_cuts(name, tmp_path):
    """Every cut of the input surface: camera plane near / far, spheres, midplane angle and height (both signs),
    arbitrary plane (simulation_sampling.cpp:245-295) and the min / max cuts on the seven cell values
    (simulation_coefficients.cpp:361-375), against the unmodified reference's images; the parameter sets are the
    generator's (tests/golden/make_golden.py CPU_CASES)."""
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden import CPU_CASES
    kv = load_input('simulation.input')
    kv.update(CPU_CASES['cpu_simulation_%s_16' % name])
    path = os.path.join(tmp_path, 'o.input')
    write_input(path, kv)
    cfg = bl.Config(path)
    grid = mock_snapshot.grid_view_arrays(mock_snapshot.make_mock(None))
    pos, dirs, fac = cfg.camera_root()
    s = oracle_lib.trace(kv, float(kv['simulation_a']), pos, dirs)
    image, _ = oracle_lib.simulation_image(kv, s, fac, grid, want_inds=False, cut_camera_x=cfg.camera_frame()['cam_x'])
    ref = np.load(os.path.join(GOLDEN, 'cpu_simulation_%s_16.npz' % name))['I_nu']
    got = image.reshape(16, 16)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.array_equal(got == 0.0, ref == 0.0)
    scale = np.maximum(np.abs(ref), 1e-12 * np.nanmax(np.abs(ref)))
    assert np.max(np.abs(got - ref) / scale) < 1e-10
    # each case removes emission relative to the uncut image
    full = np.load(os.path.join(GOLDEN, 'simulation_32.npz'))['I_nu']
    assert np.nansum(ref) / ref.size < 0.9 * np.nansum(full) / full.size


@pytest.mark.parametrize('fixture', ['cpu_simulation_fallback_values_16', 'cpu_simulation_fallback_entropy_16'])
def test_oracle_fallback_values(fixture, tmp_path):
    """fallback_nan = false with a camera field of view wider than the grid: samples outside the grid carry
    fallback_rho / fallback_pgas with zero velocity and field (simulation_sampling.cpp:695-708), poorly terminated
    rays are integrated rather than blanked; visible in the cell-value averages.  All 27 auxiliary images against the
    unmodified reference's."""
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden import CPU_CASES
    kv = load_input('simulation.input')
    kv.update(CPU_CASES[fixture])
    path = os.path.join(tmp_path, 'o.input')
    write_input(path, kv)
    cfg = bl.Config(path)
    grid = mock_snapshot.grid_view_arrays(mock_snapshot.make_mock(None, entropy=kv['plasma_model'] == 'code_kappa'))
    if kv['plasma_model'] == 'code_kappa':   # the restatement takes the entropy variable last
        order = ('ind_rho', 'ind_pgas', 'ind_uu1', 'ind_uu2', 'ind_uu3', 'ind_bb1', 'ind_bb2', 'ind_bb3', 'ind_kappa')
        grid = dict(grid, prim=np.ascontiguousarray(np.stack([grid['prim'][grid[k]] for k in order]), np.float32))
    pos, dirs, fac = cfg.camera_root()
    s = oracle_lib.trace(kv, float(kv['simulation_a']), pos, dirs)
    image, aux = oracle_lib.simulation_image(kv, s, fac, grid, want_inds=False, camera_x=cfg.camera_frame()['cam_x'])
    gold = np.load(os.path.join(GOLDEN, fixture + '.npz'))
    for name in ['I_nu'] + oracle_lib.AUX_NAMES:
        ref = gold[name]
        got = (image if name == 'I_nu' else aux[name]).reshape(16, 16)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), name
        ok = ~np.isnan(ref)
        scale = np.maximum(np.maximum(np.abs(ref[ok]), 1e-12 * np.nanmax(np.abs(ref))), 1e-300)
        assert float(np.max(np.abs(got[ok] - ref[ok]) / scale)) < 1e-9, name
    # the fallback samples (52 < r < 80, outside the grid) are in those averages: another fallback density moves them
    _, other = oracle_lib.simulation_image(dict(kv, fallback_rho='1.0e-3'), s, fac, grid, want_inds=False,
                                           camera_x=cfg.camera_frame()['cam_x'])
    assert np.nanmin(np.abs(other['lambda_ave_rho'] / aux['lambda_ave_rho'] - 1.0)) > 1e-3
    if kv['plasma_model'] == 'code_kappa':   # and the fallback entropy sets the temperature there
        _, other = oracle_lib.simulation_image(dict(kv, fallback_kappa='1.0e7'), s, fac, grid, want_inds=False,
                                               camera_x=cfg.camera_frame()['cam_x'])
        assert np.nanmin(np.abs(other['lambda_ave_Theta_e'] / aux['lambda_ave_Theta_e'] - 1.0)) > 1e-3


@pytest.mark.parametrize('name', ['cpu_slow_light_blend_12', 'cpu_slow_light_nearest_slice_12',
                                  'cpu_slow_light_blend_nearest_cell_12'])
def test_oracle_slow_light(name, tmp_path):
    """slow_light_on: which files of a time series are resident for the first image (simulation_reader.cpp:211-262),
    the time slice of each sample at its coordinate time + snapshot time and the nearest-slice / blended values
    (simulation_sampling.cpp:297-349, 736-775, 840-905), against the unmodified reference's image."""
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden import SLOW_CASES, SLOW_DT_FILE, slow_light_setup
    kv, grids = slow_light_setup(str(tmp_path), SLOW_CASES[name])
    path = os.path.join(tmp_path, 'o.input')
    write_input(path, kv)
    cfg = bl.Config(path)
    file_times = [SLOW_DT_FILE * n for n in range(len(grids))]
    snapshot_time = float(kv['slow_t_start'])
    window = oracle_lib.slow_window(file_times, int(kv['slow_chunk_size']), snapshot_time)
    assert window == [8, 7, 6, 5, 4, 3, 2, 1]
    views = [mock_snapshot.grid_view_arrays(grids[n]) for n in window]
    grid = dict(views[0], prim=np.ascontiguousarray(np.stack([v['prim'] for v in views]), np.float32))
    pos, dirs, fac = cfg.camera_root()
    s = oracle_lib.trace(kv, float(kv['simulation_a']), pos, dirs)
    slow = dict(times=[file_times[n] for n in window], snapshot_time=snapshot_time)
    image, _ = oracle_lib.simulation_image(kv, s, fac, grid, want_inds=False, slow=slow)
    ref = np.load(os.path.join(GOLDEN, name + '.npz'))['I_nu']
    got = image.reshape(12, 12)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    scale = np.maximum(np.abs(ref[ok]), 1e-12 * np.nanmax(np.abs(ref)))
    assert np.max(np.abs(got[ok] - ref[ok]) / scale) < 1e-10
    # the image is not that of any single file of the window
    for v in views:
        frozen, _ = oracle_lib.simulation_image(kv, s, fac, v, want_inds=False)
        assert np.nanmax(np.abs(frozen.reshape(12, 12)[ok] - ref[ok]) / scale) > 1e-3


@pytest.mark.parametrize('name', ['block_interp_amr', 'block_interp_amr_tilted', 'block_interp_blocks'])
def test_oracle_block_interpolation(name, tmp_path):
    """simulation_block_interp = true: trilinear anchors one cell beyond a MeshBlock resolved on the neighbouring block of
    the same, the coarser or the finer level, phi periodic (FindNearbyInds / InterpolateAdvanced,
    simulation_sampling.cpp:506-553, 1068-1331, 1365-1386), on a two-level mesh and a single-level multi-block one,
    against the unmodified reference's images."""
    import sys
    sys.path.insert(0, GOLDEN)
    from make_golden import CPU_CASES
    over = dict(CPU_CASES['cpu_simulation_%s_16' % name])
    mock = dict(over.pop('_mock'))
    kv = load_input('simulation.input')
    kv.update(over)
    path = os.path.join(tmp_path, 'o.input')
    write_input(path, kv)
    cfg = bl.Config(path)
    grid = mock_snapshot.grid_view_arrays(mock_snapshot.make_mock(None, tuple(mock.pop('blocks')), **mock))
    pos, dirs, fac = cfg.camera_root()
    s = oracle_lib.trace(kv, float(kv['simulation_a']), pos, dirs)
    image, _ = oracle_lib.simulation_image(kv, s, fac, grid, want_inds=False)
    ref = np.load(os.path.join(GOLDEN, 'cpu_simulation_%s_16.npz' % name))['I_nu']
    got = image.reshape(16, 16)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    scale = np.maximum(np.abs(ref[ok]), 1e-12 * np.nanmax(np.abs(ref)))
    assert np.max(np.abs(got[ok] - ref[ok]) / scale) < 1e-10
    # and the anchors matter: in-block interpolation gives another image
    plain, _ = oracle_lib.simulation_image(dict(kv, simulation_block_interp='false'), s, fac, grid, want_inds=False)
    assert np.nanmax(np.abs(plain.reshape(16, 16)[ok] - ref[ok]) / scale) > 1e-4


def test_refinement_restatement_against_reference_fixture(tmp_path):
    """Adaptive refinement decision (EvaluateBlock, radiation_adaptive.cpp:163-312; child order camera.cpp:445-459): the
    numpy restatement applied to the unmodified reference's level-0 image reproduces the reference's list of level-1
    blocks -- and the host layer's blh_camera_refined derives the same list from those flags."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import refine_oracle
    base, over, _ = CASES['adaptive_32']
    kv = load_input(base)
    kv.update({k: str(v) for k, v in over.items()})
    gold = np.load(os.path.join(GOLDEN, 'adaptive_32.npz'))
    res, bs = int(kv['camera_resolution']), int(kv['adaptive_block_size'])
    nb = res // bs
    locs = np.array([[v, u] for v in range(nb) for u in range(nb)], np.int32)
    blocks = gold['I_nu'].reshape(nb, bs, nb, bs).transpose(0, 2, 1, 3).reshape(nb * nb, bs, bs)
    flags = refine_oracle.refinement_flags(blocks, locs, 0, kv)
    assert int(flags.sum()) * 4 == int(gold['adaptive_num_blocks'][1]) == len(gold['adaptive_block_locs_1'])
    assert np.array_equal(refine_oracle.child_locs(locs, flags), gold['adaptive_block_locs_1'])
    path = os.path.join(tmp_path, 'a.input')
    write_input(path, kv)
    kids, _, _, _ = bl.Config(path).camera_refined(1, locs, flags)
    assert np.array_equal(kids, gold['adaptive_block_locs_1'])
    # every criterion of the restatement runs: make each the only active one on a block with a known answer
    ramp = np.add.outer(np.arange(8.0), 2.0 * np.arange(8.0)) + 1.0
    off = {k: '-1.0' for k in ('adaptive_val_frac', 'adaptive_abs_grad_frac', 'adaptive_rel_grad_frac',
                               'adaptive_abs_lapl_frac', 'adaptive_rel_lapl_frac')}
    for key, cut, image, want in (('val', 10.0, ramp, True), ('val', 100.0, ramp, False),
                                  ('abs_grad', 2.0, ramp, True), ('abs_grad', 3.0, ramp, False),
                                  ('rel_grad', 0.05, ramp, True), ('abs_lapl', 1e-9, ramp, False),
                                  ('abs_lapl', 1.0, ramp ** 2, True), ('rel_lapl', 1e-3, ramp ** 2, True),
                                  ('rel_lapl', 10.0, ramp ** 2, False)):
        k2 = dict(kv, **off)
        k2['adaptive_%s_frac' % key], k2['adaptive_%s_cut' % key] = '0.5', str(cut)
        assert refine_oracle.evaluate_block(image, k2) is want, (key, cut)
    assert refine_oracle.evaluate_block(np.full((8, 8), np.nan), dict(kv, adaptive_val_frac='0.0')) is False


@pytest.mark.parametrize('name', ['adaptive_value_32', 'adaptive_abs_grad_32', 'adaptive_rel_grad_32', 'adaptive_abs_lapl_32',
                                  'adaptive_rel_lapl_region_32'])
def test_refinement_restatement_each_criterion_two_levels(name, tmp_path):
    """Each refinement criterion on its own (value, absolute / relative gradient, absolute / relative Laplacian; the last
    with a forced region), two levels deep: the numpy restatement applied to the reference's level-0 and level-1 block
    images reproduces the reference's level-1 and level-2 block lists (radiation_adaptive.cpp:163-312)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    sys.path.insert(0, GOLDEN)
    import refine_oracle
    from make_golden import ADAPT_CASES
    kv = load_input('adaptive.input')
    kv.update(ADAPT_CASES[name])
    gold = np.load(os.path.join(GOLDEN, name + '.npz'))
    res, bs = int(kv['camera_resolution']), int(kv['adaptive_block_size'])
    nb = res // bs
    locs = np.array([[v, u] for v in range(nb) for u in range(nb)], np.int32)
    blocks = gold['I_nu'].reshape(nb, bs, nb, bs).transpose(0, 2, 1, 3).reshape(nb * nb, bs, bs)
    path = os.path.join(tmp_path, 'o.input')
    write_input(path, kv)
    cfg = bl.Config(path)
    for level in (0, 1):
        flags = refine_oracle.refinement_flags(blocks, locs, level, kv)
        assert 0 < int(flags.sum()) < len(flags)
        assert int(flags.sum()) * 4 == int(gold['adaptive_num_blocks'][level + 1])
        kids, _, _, _ = cfg.camera_refined(level + 1, locs, flags)     # the host layer's child list from the same flags
        locs = refine_oracle.child_locs(locs, flags)
        assert np.array_equal(kids, locs)
        assert np.array_equal(locs, gold['adaptive_block_locs_%d' % (level + 1)])
        blocks = gold['adaptive_I_nu_%d' % (level + 1)]
