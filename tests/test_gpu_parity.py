"""Parity of the CUDA path (through the C ABI) against the unmodified reference.

Two sources of truth: the committed golden fixtures (tests/golden/*.npz, produced by make_golden.py from
the reference built by oracle/Makefile) and, when oracle/_ref/blacklight travelled with the repo, a live run
of the reference on the same inputs.  Thresholds are north_star's: exact sample_flags / sample_num / sample
cell indices, <= 1e-6 relative per-pixel intensity, <= 1e-9 relative total flux.
"""
import os
import sys
import zlib

import numpy as np
import pytest

import blacklight_b200 as bl
from blacklight_b200.cases import C4_PHYSICS
from harness import REF_BIN, Case, flux_rel, rel_err
from golden.make_golden import CASES

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
PIXEL_TOL = 1e-6
FLUX_TOL = 1e-9

UNPOLARIZED = ['formula_16', 'formula_aux_12', 'simulation_32', 'simulation_nearest_24', 'simulation_blocks_24',
               'simulation_aux_16', 'simulation_kerr_24']


def run_gpu_level0(case, taps=False):
    cfg = case.config()
    ctx = bl.Context(cfg)
    if case.sim:
        ctx.upload_grid(case.grid_arrays())
    pos, dirs, fac = cfg.camera_root()
    ctx.trace_level(0, pos, dirs, fac)
    if taps and case.sim:
        ctx.set_taps(True)
    R = int(case.kv.get('render_num_images', 0)) if case.sim else 0
    image, render, stats = ctx.radiate_level(0, num_render=R)
    return cfg, ctx, image, render, stats


def image_arrays(case, image, res):
    """Split the (Q, N) image into the reference's named npz arrays (numpy_format.cpp:129-283)."""
    kv = case.kv
    F = int(kv['image_num_frequencies'])
    on = lambda k: kv.get(k, 'false') == 'true'
    pol = case.sim and on('image_light') and on('image_polarization')
    out, q = {}, 0
    shape = (res, res) if F == 1 else (F, res, res)
    if on('image_light'):
        if pol:
            block = image[q:q + 4 * F].reshape(F, 4, -1)
            for s, name in enumerate(('I_nu', 'Q_nu', 'U_nu', 'V_nu')):
                out[name] = block[:, s].reshape(shape)
            q += 4 * F
        else:
            out['I_nu'] = image[q:q + F].reshape(shape)
            q += F
    for key, name, per_freq in (('image_time', 'time', False), ('image_length', 'length', False),
                                ('image_lambda', 'lambda', True), ('image_emission', 'emission', True),
                                ('image_tau', 'tau', True)):
        if on(key):
            n = F if per_freq else 1
            out[name] = image[q:q + n].reshape(shape if per_freq else (res, res))
            q += n
    cells = ('rho', 'n_e', 'p_gas', 'Theta_e', 'B', 'sigma', 'beta_inverse')
    for key, prefix in (('image_lambda_ave', 'lambda_ave_'), ('image_emission_ave', 'emission_ave_'),
                        ('image_tau_int', 'tau_int_')):
        if case.sim and on(key):
            block = image[q:q + 7 * F].reshape(F, 7, -1)
            for c, cname in enumerate(cells):
                out[prefix + cname] = block[:, c].reshape(shape)
            q += 7 * F
    if on('image_crossings'):
        out['crossings'] = image[q].reshape(res, res)
        q += 1
    assert q == image.shape[0]
    return out


def check_images(mine, ref, what):
    for name, arr in mine.items():
        assert name in ref, name
        err = rel_err(arr, ref[name])
        assert err <= PIXEL_TOL, '%s %s: per-pixel relative error %.3e' % (what, name, err)
        if name.endswith('_nu'):
            assert flux_rel(arr, ref[name]) <= FLUX_TOL or np.nanmax(np.abs(ref[name])) == 0, '%s %s flux' % (what, name)


@pytest.mark.parametrize('name', UNPOLARIZED)
def test_golden_unpolarized(name, gpu, tmp_path):
    base, over, mock = CASES[name]
    gold = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    case = Case(tmp_path, base, over, mock=mock)
    cfg, ctx, image, _, stats = run_gpu_level0(case, taps=True)
    res = cfg.resolution
    s = ctx.download_samples(0)
    # termination flags, sample counts: exact
    assert np.array_equal(s['flags'], gold['sample_flags'])
    assert np.array_equal(s['num'], gold['sample_num'])
    assert stats['geodesic_num_steps'] == int(gold['geodesic_num_steps'])
    # every stored sample of every ray: bit-identical (checksum) + explicit probes
    S = s['pos'].shape[1]
    mask = np.arange(S)[None, :] < s['num'][:, None]
    crc = zlib.crc32(s['pos'][mask].tobytes() + s['dir'][mask].tobytes() + s['len'][mask].tobytes())
    assert crc == int(gold['samples_crc']), 'geodesic samples are not bit-identical to the reference'
    for r in gold['probe_rays']:
        n = s['num'][r]
        assert np.array_equal(s['pos'][r, :n], gold['probe_pos_%d' % r])
        assert np.array_equal(s['dir'][r, :n], gold['probe_dir_%d' % r])
        assert np.array_equal(s['len'][r, :n], gold['probe_len_%d' % r])
    # sampled cell indices: exact on every sample the reference defines
    if case.sim:
        t = ctx.download_sample_inds(0, interp=case.kv['simulation_interp'] == 'true')
        valid = mask & (t['cut'] == 0) & (t['nan'] == 0) & (t['fallback'] == 0)
        assert int(valid.sum()) == int(gold['valid_count'])
        assert int(t['nan'][mask].sum()) == int(gold['sample_nan_count'])
        assert zlib.crc32(np.ascontiguousarray(t['inds'][valid]).tobytes()) == int(gold['inds_crc'])
    check_images(image_arrays(case, image, res), gold, name)
    ctx.close()


@pytest.mark.parametrize('base,over,mock', [
    ('simulation.input', {'camera_resolution': 64}, None),
    ('simulation.input', {'camera_resolution': 48, 'simulation_a': '0.5', 'camera_th': '60.0', 'camera_type': 'pinhole'}, {'blocks': (1, 4, 8)}),
    ('simulation.input', {'camera_resolution': 40, 'ray_integrator': 'rk4', 'ray_step': '0.02'}, None),
    ('simulation.input', {'camera_resolution': 40, 'ray_integrator': 'rk2', 'ray_step': '0.02'}, None),
    ('simulation.input', {'camera_resolution': 40, 'plasma_power_frac': '0.3', 'plasma_p': '3.0', 'plasma_gamma_min': '4.0', 'plasma_gamma_max': '1000.0'}, None),
    ('simulation.input', {'camera_resolution': 32, 'plasma_kappa_frac': '0.4', 'plasma_kappa': '4.2', 'plasma_w': '0.8',
                          'plasma_power_frac': '0.2', 'plasma_p': '2.5', 'plasma_gamma_min': '2.0', 'plasma_gamma_max': '500.0',
                          'image_num_frequencies': 3, 'image_frequency_start': '8.6e10', 'image_frequency_end': '6.9e11',
                          'image_frequency_spacing': 'log'}, None),
    ('simulation.input', {'camera_resolution': 32, 'fallback_nan': 'false', 'fallback_rho': '1.0e-6', 'fallback_pgas': '1.0e-8', 'camera_r': '80.0', 'camera_width': '60.0'}, None),
    ('simulation.input', {'camera_resolution': 32, 'cut_omit_near': 'true', 'cut_omit_in': '3.0', 'cut_midplane_theta': '30.0', 'cut_rho_min': '1.0e-18'}, None),
    ('simulation.input', {'camera_resolution': 32, 'plasma_use_p': 'false', 'plasma_gamma_i': '1.6666666666666667',
                          'plasma_gamma_e': '1.3333333333333333', 'plasma_gamma': '1.5'}, None),
    ('simulation.input', {'camera_resolution': 28, 'cut_omit_far': 'true', 'cut_omit_out': '30.0', 'cut_midplane_z': '6.0',
                          'cut_plane': 'true', 'cut_plane_origin': '1.0,0.0,0.5', 'cut_plane_normal': '0.2,1.0,0.1',
                          'cut_sigma_max': '-1.0', 'cut_beta_inverse_max': '5.0', 'cut_theta_e_max': '50.0'}, None),
    ('simulation.input', {'camera_resolution': 24, 'image_num_frequencies': 2, 'image_frequency_start': '1.0e11',
                          'image_frequency_end': '4.0e11', 'image_frequency_spacing': 'log', 'image_time': 'true',
                          'image_length': 'true', 'image_lambda': 'true', 'image_emission': 'true', 'image_tau': 'true',
                          'image_lambda_ave': 'true', 'image_emission_ave': 'true', 'image_tau_int': 'true',
                          'image_crossings': 'true'}, {'blocks': (1, 2, 2)}),
    ('formula.input', {'camera_resolution': 24, 'image_num_frequencies': 3, 'image_frequency_start': '1.0e11', 'image_frequency_end': '4.0e11', 'image_frequency_spacing': 'lin_wave', 'camera_th': '0.0'}, None),
    ('formula.input', {'camera_resolution': 20, 'ray_flat': 'true', 'formula_l0': '1.0'}, None),
])
def test_live_reference_unpolarized(base, over, mock, gpu, tmp_path):
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    case = Case(tmp_path, base, over, mock=mock)
    ref = case.run_reference()
    cfg, ctx, image, _, stats = run_gpu_level0(case, taps=True)
    s = ctx.download_samples(0)
    g = ref['geo']
    assert np.array_equal(s['flags'], g['sample_flags'])
    assert np.array_equal(s['num'], g['sample_num'])
    S = s['pos'].shape[1]
    mask = np.arange(S)[None, :] < s['num'][:, None]
    assert np.array_equal(s['pos'][mask], g['sample_pos'][mask])
    assert np.array_equal(s['dir'][mask], g['sample_dir'][mask])
    assert np.array_equal(s['len'][mask], g['sample_len'][mask])
    if case.sim:
        interp = case.kv['simulation_interp'] == 'true'
        t = ctx.download_sample_inds(0, interp=interp)
        rs = ref['samp']
        assert np.array_equal(t['nan'][mask], rs['sample_nan'][mask])
        assert np.array_equal(t['fallback'][mask], rs['sample_fallback'][mask])
        valid = mask & (t['cut'] == 0) & (t['nan'] == 0) & (t['fallback'] == 0)
        assert np.array_equal(t['inds'][valid], rs['sample_inds'][valid])
        if interp:
            assert np.max(np.abs(t['fracs'][valid] - rs['sample_fracs'][valid])) < 1e-9
    check_images(image_arrays(case, image, cfg.resolution), ref['npz'], str(over))
    ctx.close()


AMR_MOCK = dict(blocks=(2, 2, 4), n_r=32, n_th=16, n_ph=32)


def _amr_refine(bi, bj, bk):
    return bi == 0 and bj == 1


@pytest.mark.parametrize('over,refined', [
    ({'camera_resolution': 40}, True),
    ({'camera_resolution': 32, 'camera_th': '70.0', 'image_polarization': 'true'}, True),
    ({'camera_resolution': 32}, False),
])
def test_live_reference_block_interpolation(over, refined, gpu, tmp_path):
    """simulation_block_interp = true: trilinear anchors resolved across MeshBlocks of the same, coarser and
    finer refinement level (reference FindNearbyInds, simulation_sampling.cpp:1068-1321) on a two-level mock
    mesh (and on a single-level multi-block one), against the reference binary."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    mock = dict(AMR_MOCK)
    if refined:
        mock['refine'] = _amr_refine
    over = dict(over, simulation_block_interp='true')
    case = Case(tmp_path, 'simulation.input', over, mock=mock)
    pol = over.get('image_polarization') == 'true'
    ref = case.run_reference(checkpoints=not pol)
    cfg, ctx, image, _, _ = run_gpu_level0(case, taps=not pol)
    mine = image_arrays(case, image, cfg.resolution)
    if not pol:
        s = ctx.download_samples(0)
        S = s['pos'].shape[1]
        mask = np.arange(S)[None, :] < s['num'][:, None]
        t = ctx.download_sample_inds(0, interp=True)
        rs = ref['samp']
        assert np.array_equal(t['nan'][mask], rs['sample_nan'][mask])
        valid = mask & (t['cut'] == 0) & (t['nan'] == 0) & (t['fallback'] == 0)
        # first anchor of the eight (the reference stores all eight, (N,S,8,4))
        assert np.array_equal(t['inds'][valid], rs['sample_inds'][:, :, 0, :][valid])
        assert np.max(np.abs(t['fracs'][valid] - rs['sample_fracs'][valid])) < 1e-9
    check_images(mine, ref['npz'], str(over)) if not pol else None
    if pol:
        assert rel_err(mine['I_nu'], ref['npz']['I_nu']) <= PIXEL_TOL
        # This camera has pixels with |V| ~ 1e-4 I; the polarized kernel agrees with the reference to ~1e-8 I
        # per Stokes component (same with intra-block interpolation -- the sampling itself is exact, see the
        # unpolarized cases), so the components are measured against max(|ref|, 1e-2 I) here.
        errs = stokes_err(mine, ref['npz'], floor=1e-2)
        print('block-interp polarized errors', errs)
        for k, v in errs.items():
            assert v <= PIXEL_TOL, '%s %.3e' % (k, v)
    ctx.close()


@pytest.mark.parametrize('over', [
    {'camera_resolution': 32},
    {'camera_resolution': 32, 'simulation_block_interp': 'true'},
    {'camera_resolution': 28, 'simulation_interp': 'false'},
    {'camera_resolution': 24, 'image_polarization': 'true'},
])
def test_live_reference_cartesian_kerr_schild(over, gpu, tmp_path):
    """simulation_coord = cks: a uniform Cartesian Kerr-Schild box of 2x2x2 MeshBlocks (blacklight_b200/mock_snapshot.py:
    mock_fields_cks) -- the Cartesian branches of the sampling map, the simulation metric and the frame
    transformation (reference radiation_geometry.cpp:73-91,425-457), a = 0.5."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    over = dict(over, simulation_coord='cks', simulation_a='0.5')
    case = Case(tmp_path, 'simulation.input', over, mock=dict(blocks=(2, 2, 2), cks=dict(n=32)))
    pol = over.get('image_polarization') == 'true'
    ref = case.run_reference(checkpoints=not pol)
    cfg, ctx, image, _, _ = run_gpu_level0(case, taps=not pol)
    mine = image_arrays(case, image, cfg.resolution)
    if not pol:
        s = ctx.download_samples(0)
        S = s['pos'].shape[1]
        mask = np.arange(S)[None, :] < s['num'][:, None]
        interp = over.get('simulation_interp', 'true') == 'true'
        t = ctx.download_sample_inds(0, interp=interp)
        rs = ref['samp']
        assert np.array_equal(t['nan'][mask], rs['sample_nan'][mask])
        valid = mask & (t['cut'] == 0) & (t['nan'] == 0) & (t['fallback'] == 0)
        ri = rs['sample_inds'] if rs['sample_inds'].ndim == 3 else rs['sample_inds'][:, :, 0, :]
        assert np.array_equal(t['inds'][valid], ri[valid])
        check_images(mine, ref['npz'], str(over))
    else:
        # In this synthetic box a few left-edge pixels are 1e-8 ... 1e-10 of the peak brightness: their rays cross the
        # dense midplane in 3-unit cells with optical depths of tens per step, where the reference's closed-form step
        # amplifies round-off to 1e-6 ... 1e-4 of the pixel value (measured against an 80-bit evaluation of the same
        # recurrence in test_cks_polarized_faint_pixels_are_roundoff_limited, which checks EVERY pixel against that
        # per-pixel bound).  Here the plain 1e-6 bound is applied to the pixels above 1e-6 of the peak.
        I_ref = ref['npz']['I_nu']
        bright = I_ref >= 1e-6 * np.nanmax(I_ref)
        assert bright.sum() > 0.5 * bright.size
        m = {k: np.where(bright, v, 0.0) for k, v in mine.items() if k.endswith('_nu')}
        r = {k: np.where(bright, ref['npz'][k], 0.0) for k in m}
        assert rel_err(m['I_nu'], r['I_nu']) <= PIXEL_TOL
        errs = stokes_err(m, r, floor=1e-2)
        print('cks polarized errors', errs)
        for k, v in errs.items():
            assert v <= PIXEL_TOL, '%s %.3e' % (k, v)
    ctx.close()


@pytest.mark.parametrize('over', [
    {'camera_resolution': 32},
    {'camera_resolution': 28, 'simulation_interp': 'false'},
    {'camera_resolution': 24, 'image_polarization': 'true'},
])
def test_live_reference_code_kappa(over, gpu, tmp_path):
    """plasma_model = code_kappa: electron temperature from an electron-entropy variable of the snapshot
    (simulation_coefficients.cpp:351-358; the variable is gathered and interpolated like the other primitives,
    with its own non-positive fallback), through the .athdf reader of the drop-in executable."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    over = dict(over, plasma_model='code_kappa', simulation_kappa_name='r0')
    case = Case(tmp_path, 'simulation.input', over, mock=dict(blocks=(1, 2, 2), entropy=True))
    ref = case.run_reference(checkpoints=False)
    cfg, ctx, image, _, _ = run_gpu_level0(case)
    mine = image_arrays(case, image, cfg.resolution)
    assert float(np.nanmax(ref['npz']['I_nu'])) > 0.0
    assert rel_err(mine['I_nu'], ref['npz']['I_nu']) <= PIXEL_TOL
    assert flux_rel(mine['I_nu'], ref['npz']['I_nu']) <= FLUX_TOL
    if over.get('image_polarization') == 'true':
        for k, v in stokes_err(mine, ref['npz']).items():
            assert v <= PIXEL_TOL, '%s %.3e' % (k, v)
    ctx.close()
    npz, _ = case.run_gpu_file()   # the reader locates the variable by name
    assert rel_err(npz['I_nu'], ref['npz']['I_nu']) <= PIXEL_TOL


def test_waves_match_resident(gpu, tmp_path):
    """Tracing in waves (step buffer reused) must give the same image as a resident level, bit for bit."""
    case = Case(tmp_path, 'simulation.input', {'camera_resolution': 48})
    _, ctx, image, _, _ = run_gpu_level0(case)
    cfg = case.config(tile_rays=512)
    ctx2 = bl.Context(cfg)
    ctx2.upload_grid(case.grid_arrays())
    pos, dirs, fac = cfg.camera_root()
    ctx2.trace_level(0, pos, dirs, fac)
    image2, _, stats = ctx2.radiate_level(0)
    assert np.array_equal(image, image2, equal_nan=True)
    s1, s2 = ctx.download_samples(0, arrays=False), ctx2.download_samples(0, arrays=False)
    assert np.array_equal(s1['num'], s2['num']) and np.array_equal(s1['flags'], s2['flags'])
    ctx.close()
    ctx2.close()


def test_drop_in_executable_path(gpu, tmp_path):
    """blh_run_input_file (the re-hosted main: parser, athdf reader, npz writer) against the golden image."""
    base, over, mock = CASES['simulation_32']
    case = Case(tmp_path, base, over, mock=mock)
    npz, timings = case.run_gpu_file()
    gold = dict(np.load(os.path.join(GOLDEN, 'simulation_32.npz')))
    for k in ('mass_msun', 'width', 'frequency', 'adaptive_num_levels'):
        assert np.array_equal(npz[k], gold[k]), k
    assert npz['I_nu'].shape == gold['I_nu'].shape
    assert rel_err(npz['I_nu'], gold['I_nu']) <= PIXEL_TOL
    assert flux_rel(npz['I_nu'], gold['I_nu']) <= FLUX_TOL
    assert timings['rays'] == 32 * 32


def stokes_err(mine, ref, floor=1e-3):
    """Q, U, V can vanish where I does not: measure their error against max(|ref|, floor * I) per pixel."""
    out = {}
    I = ref['I_nu']
    for name in ('Q_nu', 'U_nu', 'V_nu'):
        a, b = mine[name], ref[name]
        ok = ~np.isnan(b)
        assert np.array_equal(np.isnan(a), np.isnan(b))
        scale = np.maximum(np.abs(b[ok]), floor * np.abs(I[ok]))
        scale = np.where(scale > 0, scale, 1.0)
        out[name] = float(np.max(np.abs(a[ok] - b[ok]) / scale)) if ok.any() else 0.0
    return out


@pytest.mark.parametrize('name', ['polarized_thermal_16', 'polarized_kappa_multi_12'])
def test_golden_polarized(name, gpu, tmp_path):
    base, over, mock = CASES[name]
    gold = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    case = Case(tmp_path, base, over, mock=mock)
    cfg, ctx, image, _, stats = run_gpu_level0(case)
    s = ctx.download_samples(0, arrays=False)
    assert np.array_equal(s['flags'], gold['sample_flags'])
    assert np.array_equal(s['num'], gold['sample_num'])
    mine = image_arrays(case, image, cfg.resolution)
    err_i = rel_err(mine['I_nu'], gold['I_nu'])
    assert err_i <= PIXEL_TOL, 'I_nu %.3e' % err_i
    assert flux_rel(mine['I_nu'], gold['I_nu']) <= FLUX_TOL
    errs = stokes_err(mine, gold)
    for k, v in errs.items():
        assert v <= PIXEL_TOL, '%s %.3e' % (k, v)
    ctx.close()


@pytest.mark.parametrize('over', [
    {'camera_resolution': 24, 'image_polarization': 'true', 'image_rotation_split': 'true'},
    {'camera_resolution': 20, 'image_polarization': 'true', 'plasma_power_frac': '0.5', 'plasma_p': '3.0',
     'plasma_gamma_min': '4.0', 'plasma_gamma_max': '1000.0', 'plasma_kappa_frac': '0.25', 'plasma_kappa': '3.7', 'plasma_w': '1.5',
     'image_tau': 'true', 'image_emission': 'true'},
    {'camera_resolution': 20, 'image_polarization': 'true', 'simulation_a': '0.9', 'camera_th': '20.0', 'camera_rotation': '30.0',
     'image_normalization': 'camera', 'camera_urn': '0.1'},
    {'camera_resolution': 16, 'image_polarization': 'true', 'image_num_frequencies': 2, 'image_frequency_start': '2.3e11',
     'image_frequency_end': '4.6e11', 'image_frequency_spacing': 'log', 'image_time': 'true', 'image_length': 'true',
     'image_lambda': 'true', 'image_emission': 'true', 'image_tau': 'true', 'image_lambda_ave': 'true',
     'image_emission_ave': 'true', 'image_tau_int': 'true', 'image_crossings': 'true'},
])
def test_live_reference_polarized(over, gpu, tmp_path):
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    case = Case(tmp_path, 'simulation.input', over)
    ref = case.run_reference(checkpoints=False)
    cfg, ctx, image, _, _ = run_gpu_level0(case)
    mine = image_arrays(case, image, cfg.resolution)
    assert rel_err(mine['I_nu'], ref['npz']['I_nu']) <= PIXEL_TOL
    assert flux_rel(mine['I_nu'], ref['npz']['I_nu']) <= FLUX_TOL
    errs = stokes_err(mine, ref['npz'])
    print('polarized errors', over, errs)
    for k, v in errs.items():
        assert v <= PIXEL_TOL, '%s %.3e' % (k, v)
    for k in mine:
        if not k.endswith('_nu'):
            assert rel_err(mine[k], ref['npz'][k]) <= PIXEL_TOL, k
    ctx.close()


@pytest.mark.parametrize('base,over', [
    ('true_color.input', {'camera_resolution': 16}),
    ('render.input', {'camera_resolution': 24}),
])
def test_time_series_true_color_and_render_against_reference(base, over, gpu, tmp_path):
    """BASELINE config 5: true-colour (10 unpolarized frequencies) and render (flat space, no light image) modes over
    a time series of mock snapshots -- simulation_multiple with a {05d} field in simulation_file and output_file
    (simulation_reader.cpp:870-904, output_writer.cpp:283-316), one geodesic pass reused by every frame -- through
    the drop-in executable against the reference binary, every array of every frame."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    import subprocess
    from blacklight_b200 import mock_snapshot as ms
    from harness import write_input
    d = str(tmp_path)
    case = Case(d, base, over)
    for n in range(3):
        grid = ms.to_blocks(ms.mock_fields(n_r=32, n_th=16, n_ph=32, pert_amp=0.1 + 0.2 * n, pert_n_ph=3 + n), (1, 1, 2))
        ms.write_athdf(os.path.join(d, 'data', 'mock.%05d.athdf' % (n + 4)), grid, time=10.0 * n)
    frames = {}
    for who in ('ref', 'gpu'):
        out = os.path.join(d, 'out_' + who)
        os.makedirs(out)
        kv = dict(case.kv)
        kv.update({'simulation_file': os.path.join(d, 'data', 'mock.{05d}.athdf'), 'simulation_multiple': 'true',
                   'simulation_start': '4', 'simulation_end': '6', 'output_file': os.path.join(out, 'frame.{05d}.npz')})
        path = os.path.join(d, who + '.input')
        write_input(path, kv)
        if who == 'ref':
            proc = subprocess.run([REF_BIN, path], cwd=d, capture_output=True, text=True, timeout=3600)
            assert proc.returncode == 0 and 'Calculation completed' in proc.stdout, proc.stdout + proc.stderr
        else:
            bl.run_input_file(path)
        frames[who] = [dict(np.load(os.path.join(out, 'frame.%05d.npz' % k))) for k in (4, 5, 6)]
    key = 'rendering' if base == 'render.input' else 'I_nu'
    for ref, mine in zip(frames['ref'], frames['gpu']):
        assert sorted(ref) == sorted(mine)
        for name in ref:
            if ref[name].dtype.kind == 'f' and ref[name].size > 16:
                assert mine[name].shape == ref[name].shape
                assert rel_err(mine[name], ref[name]) <= PIXEL_TOL, name
        assert float(np.nanmax(np.abs(ref[key]))) > 0.0
    assert rel_err(frames['gpu'][2][key], frames['gpu'][0][key]) > 1e-3      # the series evolves


def _format_case_names():
    sys.path.insert(0, GOLDEN)
    from make_golden_formats import FORMAT_CASES
    return sorted(FORMAT_CASES)


@pytest.mark.parametrize('name', _format_case_names())
def test_golden_snapshot_formats(name, gpu, tmp_path):
    """AthenaK, iharm3d (MKS and FMKS) and harm3d dumps against committed fixtures of the unmodified reference
    (tests/golden/make_golden_formats.py rebuilds the same deterministic mock dump): our reader + the drop-in path."""
    from make_golden_formats import write_case
    gold = dict(np.load(os.path.join(GOLDEN, 'formats_%s.npz' % name)))
    path, out = write_case(name, str(tmp_path))
    bl.run_input_file(path)
    mine = dict(np.load(out))
    bright = gold['I_nu'] >= 1e-6 * np.nanmax(gold['I_nu'])     # see test_live_reference_cartesian_kerr_schild
    bright &= gold['defined']      # FMKS: pixels where the reference reads past its arrays (make_golden_formats.py)
    assert bright.sum() > 0.5 * bright.size
    names = [k for k in gold if k.endswith('_nu')]
    m = {k: np.where(bright, mine[k], 0.0) for k in names}
    g = {k: np.where(bright, gold[k], 0.0) for k in names}
    assert rel_err(m['I_nu'], g['I_nu']) <= PIXEL_TOL
    assert flux_rel(m['I_nu'], g['I_nu']) <= FLUX_TOL
    if 'Q_nu' in gold:
        for k, v in stokes_err(m, g, floor=1e-2).items():
            assert v <= PIXEL_TOL, '%s %.3e' % (k, v)


def test_golden_render(gpu, tmp_path):
    base, over, mock = CASES['render_32']
    gold = dict(np.load(os.path.join(GOLDEN, 'render_32.npz')))
    case = Case(tmp_path, base, over, mock=mock)
    cfg, ctx, image, render, _ = run_gpu_level0(case)
    s = ctx.download_samples(0, arrays=False)
    assert np.array_equal(s['num'], gold['sample_num'])
    got = render.reshape(gold['rendering'].shape)
    assert rel_err(got, gold['rendering']) <= PIXEL_TOL
    ctx.close()


def test_golden_true_color(gpu, tmp_path):
    base, over, mock = CASES['true_color_16']
    gold = dict(np.load(os.path.join(GOLDEN, 'true_color_16.npz')))
    case = Case(tmp_path, base, over, mock=mock)
    cfg, ctx, image, _, _ = run_gpu_level0(case)
    mine = image_arrays(case, image, cfg.resolution)
    assert mine['I_nu'].shape == gold['I_nu'].shape
    assert rel_err(mine['I_nu'], gold['I_nu']) <= PIXEL_TOL
    assert np.array_equal(np.load(os.path.join(GOLDEN, 'true_color_16.npz'))['frequency'], gold['frequency'])
    ctx.close()


def test_adaptive_drop_in(gpu, tmp_path):
    """example_adaptive through the re-hosted main: refinement decisions, child block order, per-level images."""
    base, over, mock = CASES['adaptive_32']
    gold = dict(np.load(os.path.join(GOLDEN, 'adaptive_32.npz')))
    case = Case(tmp_path, base, over, mock=mock)
    npz, _ = case.run_gpu_file()
    assert int(npz['adaptive_num_levels'][0]) == int(gold['adaptive_num_levels'][0])
    assert np.array_equal(npz['adaptive_num_blocks'], gold['adaptive_num_blocks'])
    for k in gold:
        if k.startswith('adaptive_block_locs'):
            assert np.array_equal(npz[k], gold[k]), k
    for k in ('I_nu', 'tau', 'adaptive_I_nu_1', 'adaptive_tau_1'):
        assert npz[k].shape == gold[k].shape, k
        assert rel_err(npz[k], gold[k]) <= PIXEL_TOL, k
    I = {'I_nu': gold['I_nu'], 'Q_nu': gold['Q_nu'], 'U_nu': gold['U_nu'], 'V_nu': gold['V_nu']}
    for k, v in stokes_err({n: npz[n] for n in I}, I).items():
        assert v <= PIXEL_TOL, k


def _device_image_tensor(ctx, level):
    """torch view of a level's image in HBM (bl_device_image), as bench.py hands it to the sharded adaptive worker."""
    import torch
    if ctx._rays.get(level, 0) == 0:
        return torch.empty((ctx.num_quantities, 0), dtype=torch.float64, device='cuda:0')

    class View:
        pass
    ptr, shape = ctx.device_image(level)
    v = View()
    v.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': '<f8', 'data': (int(ptr), False), 'version': 2}
    return torch.as_tensor(v, device='cuda:0')


@pytest.mark.parametrize('world', [1, 3, -3])
def test_adaptive_sharded_over_ranks(world, gpu, tmp_path):
    """example_adaptive with the blocks of every level (root level included) dealt round-robin over `world` ranks
    (blacklight_b200/multigpu.py; the ranks are stepped in lock step inside this process, each with its own
    context): refinement flags, child block order and per-level images must not depend on the number of ranks
    and must match the reference."""
    from blacklight_b200 import multigpu
    on_device = world < 0   # -3: three ranks whose images stay in HBM until the assembled levels are downloaded
    world = abs(world)
    base, over, mock = CASES['adaptive_32']
    gold = dict(np.load(os.path.join(GOLDEN, 'adaptive_32.npz')))
    case = Case(tmp_path, base, over, mock=mock)
    max_level = int(case.kv['adaptive_max_level'])
    ctxs, workers = [], []
    for rank in range(world):
        cfg = case.config()
        cfg.set_level0_block_major(True)
        ctx = bl.Context(cfg)
        ctx.upload_grid(case.grid_arrays())
        ctxs.append(ctx)
        workers.append(multigpu.adaptive_worker(cfg, ctx, rank, world, max_level,
                                                device_images=_device_image_tensor if on_device else None))
    levels = multigpu.run_local(workers)[0]
    for ctx in ctxs:
        ctx.close()
    assert len(levels) - 1 == int(gold['adaptive_num_levels'][0])
    assert np.array_equal(levels[1]['locs'], gold['adaptive_block_locs_1'])
    res, bs = 32, int(case.kv['adaptive_block_size'])
    root = image_arrays(case, levels[0]['image'], res)
    assert rel_err(root['I_nu'], gold['I_nu']) <= PIXEL_TOL
    assert rel_err(root['tau'], gold['tau']) <= PIXEL_TOL
    nb1 = len(levels[1]['locs'])
    I1 = levels[1]['image'][0].reshape(nb1, bs, bs)
    assert rel_err(I1, gold['adaptive_I_nu_1']) <= PIXEL_TOL
    # bitwise independence of the rank count: compare against the single-rank run of the same code
    key = os.path.join(str(tmp_path.parent), 'adaptive_sharded_world1.npz')
    arrays = {'l%d_%s' % (i, k): L[k] for i, L in enumerate(levels) for k in ('locs', 'image') }
    if world == 1:
        np.savez(key, **arrays)
    elif os.path.exists(key):
        one = np.load(key)
        for k, v in arrays.items():
            assert np.array_equal(one[k], v, equal_nan=True), k


@pytest.mark.parametrize('over', [
    {'slow_interp': 'true'},
    {'slow_interp': 'false'},
    {'slow_interp': 'true', 'simulation_interp': 'false'},
    {'slow_interp': 'true', 'image_polarization': 'true', 'camera_resolution': 16},
])
def test_slow_light_against_reference(over, gpu, tmp_path):
    """slow_light_on: a time series of mock snapshots (perturbation amplitude growing with time), a sliding window
    of slow_chunk_size snapshots resident in HBM, time-slice choice / interpolation per sample
    (simulation_sampling.cpp:298-349) -- three consecutive images through the drop-in executable path against
    the reference binary, including the window shift between images."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    import subprocess
    from blacklight_b200 import mock_snapshot as ms
    from harness import write_input
    d = str(tmp_path)
    case = Case(d, 'simulation.input', dict({'camera_resolution': 24}, **over))
    n_files = 12
    for n in range(n_files):
        grid = ms.to_blocks(ms.mock_fields(n_r=32, n_th=16, n_ph=32, pert_amp=0.1 + 0.05 * n), (1, 1, 2))
        ms.write_athdf(os.path.join(d, 'data', 'mock.%05d.athdf' % n), grid, time=25.0 * n)
    images = {}
    for who in ('ref', 'gpu'):
        out = os.path.join(d, 'out_' + who)
        os.makedirs(out)
        kv = dict(case.kv)
        kv.update({'simulation_file': os.path.join(d, 'data', 'mock.{05d}.athdf'), 'simulation_multiple': 'true',
                   'simulation_start': '0', 'simulation_end': str(n_files - 1), 'slow_light_on': 'true',
                   'slow_chunk_size': '8', 'slow_t_start': '200.0', 'slow_dt': '20.0', 'slow_num_images': '3',
                   'slow_offset': '5', 'output_file': os.path.join(out, 'img.{03d}.npz')})
        path = os.path.join(d, who + '.input')
        write_input(path, kv)
        if who == 'ref':
            proc = subprocess.run([REF_BIN, path], cwd=d, capture_output=True, text=True, timeout=3600)
            assert proc.returncode == 0 and 'Calculation completed' in proc.stdout, proc.stdout + proc.stderr
        else:
            bl.run_input_file(path)
        images[who] = [dict(np.load(os.path.join(out, 'img.%03d.npz' % k))) for k in (5, 6, 7)]
    pol = over.get('image_polarization') == 'true'
    for ref, mine in zip(images['ref'], images['gpu']):
        assert rel_err(mine['I_nu'], ref['I_nu']) <= PIXEL_TOL
        assert flux_rel(mine['I_nu'], ref['I_nu']) <= FLUX_TOL
        if pol:
            for k, v in stokes_err(mine, ref).items():
                assert v <= PIXEL_TOL, '%s %.3e' % (k, v)
    # the series must actually evolve
    assert rel_err(images['gpu'][2]['I_nu'], images['gpu'][0]['I_nu']) > 1e-3


@pytest.mark.parametrize('over', [
    {'camera_resolution': 32},
    {'camera_resolution': 24, 'image_polarization': 'true', 'simulation_a': '0.0'},
    {'camera_resolution': 24, 'plasma_gamma': '1.5', 'plasma_use_p': 'false', 'plasma_gamma_i': '1.6666666666666667',
     'plasma_gamma_e': '1.3333333333333333'},
])
def test_harm3d_reader_against_reference(over, gpu, tmp_path):
    """simulation_format = harm3d: ascii header + float32 cell records in modified Kerr-Schild coordinates,
    converted on the host to spherical Kerr-Schild coordinates / normal-frame primitives exactly as the
    reference's reader does (simulation_reader.cpp:661-720,808-848, simulation_geometry.cpp:29-92,242-327);
    the device path is unchanged.  Through the drop-in executable, against the reference binary."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    import subprocess
    from blacklight_b200 import mock_snapshot as ms
    from harness import write_input
    d = str(tmp_path)
    case = Case(d, 'simulation.input', over)
    snap = os.path.join(d, 'data', 'mock.harm3d')
    ms.write_harm3d(snap, ms.mock_fields(n_r=48, n_th=32, n_ph=32), time=3.0)
    images = {}
    for who in ('ref', 'gpu'):
        kv = dict(case.kv)
        kv.update({'simulation_format': 'harm3d', 'simulation_file': snap, 'simulation_coord': 'sks',
                   'output_file': os.path.join(d, who + '.npz')})
        if 'plasma_gamma' not in over:
            kv.pop('plasma_gamma', None)     # taken from the file header
        kv.pop('simulation_block_interp', None)
        path = os.path.join(d, who + '.input')
        write_input(path, kv)
        if who == 'ref':
            proc = subprocess.run([REF_BIN, path], cwd=d, capture_output=True, text=True, timeout=3600)
            assert proc.returncode == 0 and 'Calculation completed' in proc.stdout, proc.stdout + proc.stderr
        else:
            bl.run_input_file(path)
        images[who] = dict(np.load(os.path.join(d, who + '.npz')))
    ref, mine = images['ref'], images['gpu']
    assert float(np.nanmax(ref['I_nu'])) > 0.0
    assert rel_err(mine['I_nu'], ref['I_nu']) <= PIXEL_TOL
    assert flux_rel(mine['I_nu'], ref['I_nu']) <= FLUX_TOL
    if over.get('image_polarization') == 'true':
        for k, v in stokes_err(mine, ref).items():
            assert v <= PIXEL_TOL, '%s %.3e' % (k, v)


@pytest.mark.parametrize('over,sizes', [
    ({'camera_resolution': 32}, (4, 4)),
    ({'camera_resolution': 28, 'simulation_interp': 'false'}, (8, 8)),
    ({'camera_resolution': 24, 'simulation_block_interp': 'true'}, (8, 4)),
    ({'camera_resolution': 24, 'image_polarization': 'true'}, (4, 8)),
    ({'camera_resolution': 24, 'plasma_model': 'code_kappa', 'simulation_kappa_name': 'r0'}, (4, 4)),
])
def test_athenak_reader_against_reference(over, sizes, gpu, tmp_path):
    """simulation_format = athenak: AthenaK binary dump of a Cartesian Kerr-Schild box of
    2x2x2 MeshBlocks, a = 0.5, read by our host reader (faces rebuilt from block edges, eint -> pressure, adiabatic
    index from the dump's <mhd> block; reference simulation_reader.cpp:434-588,915-1131) and rendered through the
    drop-in executable, against the reference binary on the same file."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    import subprocess
    from blacklight_b200 import mock_snapshot as ms
    from harness import write_input
    d = str(tmp_path)
    case = Case(d, 'simulation.input', over)
    grid = ms.to_blocks(ms.mock_fields_cks(n=32), (2, 2, 2))
    if 'simulation_kappa_name' in over:
        ms.add_entropy(grid)
    snap = os.path.join(d, 'data', 'mock.athenak.bin')
    ms.write_athenak(snap, grid, gamma_adi=13.0 / 9.0, time=1.0, location_size=sizes[0], variable_size=sizes[1], spin=0.5)
    images = {}
    for who in ('ref', 'gpu'):
        kv = dict(case.kv)
        kv.update({'simulation_format': 'athenak', 'simulation_file': snap, 'simulation_coord': 'cks', 'simulation_a': '0.5',
                   'output_file': os.path.join(d, who + '.npz')})
        kv.pop('plasma_gamma', None)     # taken from the dump
        path = os.path.join(d, who + '.input')
        write_input(path, kv)
        if who == 'ref':
            proc = subprocess.run([REF_BIN, path], cwd=d, capture_output=True, text=True, timeout=3600)
            assert proc.returncode == 0 and 'Calculation completed' in proc.stdout, proc.stdout + proc.stderr
        else:
            bl.run_input_file(path)
        images[who] = dict(np.load(os.path.join(d, who + '.npz')))
    ref, mine = images['ref'], images['gpu']
    assert float(np.nanmax(ref['I_nu'])) > 0.0
    if over.get('image_polarization') == 'true':
        # as in test_live_reference_cartesian_kerr_schild: pixels above 1e-6 of the peak
        bright = ref['I_nu'] >= 1e-6 * np.nanmax(ref['I_nu'])
        assert bright.sum() > 0.5 * bright.size
        m = {k: np.where(bright, v, 0.0) for k, v in mine.items() if k.endswith('_nu')}
        r = {k: np.where(bright, ref[k], 0.0) for k in m}
        assert rel_err(m['I_nu'], r['I_nu']) <= PIXEL_TOL
        for k, v in stokes_err(m, r, floor=1e-2).items():
            assert v <= PIXEL_TOL, '%s %.3e' % (k, v)
    else:
        assert rel_err(mine['I_nu'], ref['I_nu']) <= PIXEL_TOL
        assert flux_rel(mine['I_nu'], ref['I_nu']) <= FLUX_TOL


@pytest.mark.parametrize('over,fmks', [
    ({'camera_resolution': 32}, False),
    ({'camera_resolution': 28, 'simulation_interp': 'false'}, False),
    ({'camera_resolution': 24, 'image_polarization': 'true'}, False),
    ({'camera_resolution': 32}, True),
    ({'camera_resolution': 28, 'simulation_interp': 'false'}, True),
    ({'camera_resolution': 24, 'image_polarization': 'true'}, True),
    ({'camera_resolution': 24, 'plasma_use_p': 'false', 'plasma_gamma_i': '1.6666666666666667'}, True),
])
def test_iharm3d_reader_against_reference(over, fmks, gpu, tmp_path):
    """simulation_format = iharm3d, simulation_coord = sks (MKS dump, converted on the host) and
    fmks (native coordinates + the reader's (r, theta) -> (x1, x2) table; cells found by scaling on the device,
    reference simulation_sampling.cpp:190-198,397-452).  Through the drop-in executable against the reference
    binary on the same dump; unpolarized cases also compare the sampled cell indices and fractions."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    import subprocess
    import refio
    from blacklight_b200 import mock_snapshot as ms
    from harness import write_input
    d = str(tmp_path)
    case = Case(d, 'simulation.input', over)
    snap = os.path.join(d, 'data', 'mock.iharm3d.h5')
    ms.write_iharm3d(snap, n_r=48, n_th=32, n_ph=32, gamma_adi=13.0 / 9.0, time=2.0, hslope=0.3 if fmks else 0.7,
                     fmks=dict(poly_xt=0.82, poly_alpha=14.0, mks_smooth=0.5) if fmks else None)
    pol = over.get('image_polarization') == 'true'
    interp = over.get('simulation_interp', 'true') == 'true'
    images, paths = {}, {}
    for who in ('ref', 'gpu'):
        kv = dict(case.kv)
        kv.update({'simulation_format': 'iharm3d', 'simulation_file': snap, 'simulation_coord': 'fmks' if fmks else 'sks',
                   'simulation_a': '0.0', 'output_file': os.path.join(d, who + '.npz')})
        kv.pop('simulation_block_interp', None)
        kv.pop('plasma_gamma', None)           # header/gam of the dump
        if 'plasma_gamma_i' in over:
            kv['plasma_gamma_e'] = '1.3333333333333333'   # the mock dump carries neither gam_p nor gam_e
        if who == 'ref' and not pol and 'plasma_gamma_i' not in over:
            kv.update({'checkpoint_sample_save': 'true', 'checkpoint_sample_load': 'false',
                       'checkpoint_sample_file': os.path.join(d, 'samp.ckpt')})
        paths[who] = os.path.join(d, who + '.input')
        write_input(paths[who], kv)
        if who == 'ref':
            proc = subprocess.run([REF_BIN, paths[who]], cwd=d, capture_output=True, text=True, timeout=3600)
            assert proc.returncode == 0 and 'Calculation completed' in proc.stdout, proc.stdout + proc.stderr
        else:
            bl.run_input_file(paths[who])
        images[who] = dict(np.load(os.path.join(d, who + '.npz')))
    ref, mine = images['ref'], images['gpu']
    assert float(np.nanmax(ref['I_nu'])) > 0.0
    # sampled cells: our reader's arrays through the C ABI with the parity taps on (unpolarized kernel, same rays)
    kv = dict(bl.parse_input_text(open(paths['gpu']).read()), image_polarization='false')
    paths['taps'] = os.path.join(d, 'taps.input')
    write_input(paths['taps'], kv)
    cfg = bl.Config(paths['taps'])
    ctx = bl.Context(cfg)
    grid = bl.read_snapshot(cfg)
    ctx.upload_grid(grid)
    pos, dirs, fac = cfg.camera_root()
    ctx.trace_level(0, pos, dirs, fac)
    ctx.set_taps(True)
    ctx.radiate_level(0)
    s = ctx.download_samples(0)
    t = ctx.download_sample_inds(0, interp=interp)
    ctx.close()
    mask = np.arange(s['pos'].shape[1])[None, :] < s['num'][:, None]
    valid = mask & (t['cut'] == 0) & (t['nan'] == 0) & (t['fallback'] == 0)
    res = int(over['camera_resolution'])
    defined = np.ones((res, res), bool)
    if fmks:
        # The FMKS lookup uses zone (i, j) and (i + 1, j + 1) for every zone, the last ones included
        # (simulation_sampling.cpp:412-446): past a row that is the next row, past a variable's last cell the
        # first cells of the next variable (reproduced) -- and past the LAST variable's last cell whatever follows the
        # reference's array in memory.  Pixels whose rays take such a sample are not defined by the reference.
        nk, nj, ni = grid['n_k'], grid['n_j'], grid['n_i']
        k_m, j_m, i_m = (t['inds'][..., c].astype(np.int64) for c in (1, 2, 3))
        reach = nj * ni + ni + 1 if interp else 0          # the farthest corner of a trilinear stencil
        past = valid & ((k_m * nj + j_m) * ni + i_m + reach >= nk * nj * ni)
        defined = ~past.any(axis=1).reshape(res, res)
        assert defined.sum() > 0.9 * defined.size
    keep = lambda img: np.where(defined, img, 0.0)
    assert rel_err(keep(mine['I_nu']), keep(ref['I_nu'])) <= PIXEL_TOL
    assert flux_rel(keep(mine['I_nu']), keep(ref['I_nu'])) <= FLUX_TOL
    if pol:
        for k, v in stokes_err({q: keep(v) for q, v in mine.items() if q.endswith('_nu')},
                               {q: keep(v) for q, v in ref.items() if q.endswith('_nu')}).items():
            assert v <= PIXEL_TOL, '%s %.3e' % (k, v)
        return
    if 'plasma_gamma_i' in over:
        return
    rs = refio.read_sample_checkpoint(os.path.join(d, 'samp.ckpt'), interp=interp)
    assert np.array_equal(t['nan'][mask], rs['sample_nan'][mask])
    assert valid.sum() > 1000
    same = np.all(t['inds'][valid] == rs['sample_inds'][valid], axis=-1)
    print('iharm3d fmks=%s: %d of %d sampled cells differ' % (fmks, int((~same).sum()), int(valid.sum())))
    # theta = acos(z / r) comes from two different math libraries: a sample within an ulp of a zone boundary could
    # land on the other side (none observed)
    assert (~same).sum() <= 1e-5 * valid.sum()
    if interp:
        both = valid.copy()
        both[valid] = same
        assert np.max(np.abs(t['fracs'][both] - rs['sample_fracs'][both])) < 1e-6


def test_adaptive_two_levels_with_forced_region_against_reference(gpu, tmp_path):
    """Two refinement levels (a forced region plus the relative-Laplacian criterion), polarized, through the
    drop-in executable: block lists, block counts and every per-level image against the reference."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    over = {'camera_resolution': 32, 'adaptive_max_level': 2, 'adaptive_block_size': 8, 'adaptive_num_regions': 1,
            'adaptive_region_1_level': 2, 'adaptive_region_1_x_min': '-3.0', 'adaptive_region_1_x_max': '5.0',
            'adaptive_region_1_y_min': '-4.0', 'adaptive_region_1_y_max': '1.0', 'adaptive_rel_lapl_cut': '0.5',
            'adaptive_rel_lapl_frac': '0.1', 'adaptive_abs_grad_frac': '-1.0'}
    case = Case(tmp_path, 'adaptive.input', over)
    ref = case.run_reference(checkpoints=False)['npz']
    mine, _ = case.run_gpu_file()
    assert int(ref['adaptive_num_levels'][0]) == 2
    assert int(mine['adaptive_num_levels'][0]) == 2
    assert np.array_equal(mine['adaptive_num_blocks'], ref['adaptive_num_blocks'])
    for level in (1, 2):
        assert np.array_equal(mine['adaptive_block_locs_%d' % level], ref['adaptive_block_locs_%d' % level])
    for k in ref:
        if k.endswith('I_nu') or 'I_nu_' in k or k.endswith('tau') or 'tau_' in k:
            assert mine[k].shape == ref[k].shape, k
            assert rel_err(mine[k], ref[k]) <= PIXEL_TOL, k


def test_geodesic_checkpoint_exchange_with_reference(gpu, tmp_path):
    """checkpoint_geodesic_load: geodesics integrated by the REFERENCE (its checkpoint file, reference byte format)
    are loaded into the step buffer instead of tracing (bl_upload_samples).  Because the CUDA integrator is
    bit-identical to the reference's, the image from the loaded geodesics must equal the image from our own
    tracing bit for bit -- and the reference's image to tolerance.  The opposite direction (our checkpoint read
    back) is covered too."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    case = Case(tmp_path, 'simulation.input', {'camera_resolution': 40, 'simulation_a': '0.7'})
    ref = case.run_reference()
    ref_ckpt = os.path.join(case.dir, 'out_ref', 'geo.ckpt')
    own, _ = case.run_gpu_file()
    mine_ckpt = os.path.join(case.dir, 'mine.ckpt')
    saved, _ = case.run_gpu_file(extra={'checkpoint_geodesic_save': 'true', 'checkpoint_geodesic_file': mine_ckpt}, tag='save')
    assert open(mine_ckpt, 'rb').read() == open(ref_ckpt, 'rb').read(), 'checkpoint files differ from the reference byte for byte'
    for name, ckpt in (('reference checkpoint', ref_ckpt), ('own checkpoint', mine_ckpt)):
        loaded, _ = case.run_gpu_file(extra={'checkpoint_geodesic_load': 'true', 'checkpoint_geodesic_file': ckpt}, tag='load')
        assert np.array_equal(loaded['I_nu'], own['I_nu'], equal_nan=True), name
        assert rel_err(loaded['I_nu'], ref['npz']['I_nu']) <= PIXEL_TOL, name


@pytest.mark.parametrize('interp', ['true', 'false'])
def test_sample_checkpoint_written_in_reference_format(interp, gpu, tmp_path):
    """checkpoint_sample_save = true through the drop-in executable: the file has the reference's layout
    (sample_checkpoint.cpp:22-39; arrays as file_io.cpp:64-75 dumps them) and, wherever the reference defines
    them, its values: cell indices exactly, fractions to rounding, NaN / fallback flags exactly."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    import refio
    case = Case(tmp_path, 'simulation.input', {'camera_resolution': 24, 'simulation_interp': interp},
                mock=dict(blocks=(7, 2, 4)))
    ref = case.run_reference()
    path = os.path.join(case.dir, 'gpu_samp.ckpt')
    npz, _ = case.run_gpu_file(extra={'checkpoint_sample_save': 'true', 'checkpoint_sample_load': 'false',
                                      'checkpoint_sample_file': path})
    mine = refio.read_sample_checkpoint(path, interp=interp == 'true')
    rs, geo = ref['samp'], ref['geo']
    assert mine['sample_inds'].shape == rs['sample_inds'].shape
    S = rs['sample_nan'].shape[1]
    mask = np.arange(S)[None, :] < geo['sample_num'][:, None]
    assert np.array_equal(mine['sample_nan'][mask], rs['sample_nan'][mask])
    assert np.array_equal(mine['sample_fallback'][mask], rs['sample_fallback'][mask])
    valid = mask & (rs['sample_nan'] == 0) & (rs['sample_fallback'] == 0) & (mine['sample_inds'][..., 0] >= 0)
    assert valid.sum() > 1000
    assert np.array_equal(mine['sample_inds'][valid], rs['sample_inds'][valid])
    if interp == 'true':
        assert mine['sample_fracs'].shape == rs['sample_fracs'].shape
        assert np.max(np.abs(mine['sample_fracs'][valid] - rs['sample_fracs'][valid])) < 1e-10
    assert rel_err(npz['I_nu'], ref['npz']['I_nu']) <= PIXEL_TOL


def test_division_sqrt_sequences(gpu, tmp_path):
    """The geodesic kernel divides and takes square roots through branch-free instruction sequences with one
    refined reciprocal per shared denominator (csrc/glibc_math.cuh: div_by, sqrt_rn).  Their results must be
    the hardware IEEE results bit for bit -- the integrator's exact parity with the reference rests on that --
    so compare 2^31 operand pairs, a quarter of them hard cases for rounding (quotients adjacent to
    representable numbers and to midpoints)."""
    case = Case(str(tmp_path), 'formula.input', {'camera_resolution': 8})
    ctx = bl.Context(case.config())
    try:
        for seed in (1, 2):
            assert ctx.selftest_division(1 << 30, seed=seed) == 0
    finally:
        ctx.close()


@pytest.mark.parametrize('root,levels,pol,window', [
    (32, 2, False, ('-4.0', '7.0', '-5.0', '2.0')),
    (256, 2, False, ('-1.0', '4.5', '-3.5', '0.5')),      # the benchmark's 1024^2 frame
    (256, 4, True, ('1.0', '3.2', '-2.5', '-1.0')),       # the 4096^2 polarized target frame (traced in waves)
    (128, 5, 'c4', ('1.51', '2.99', '-2.99', '-1.51')),   # the same frame with the benchmark's physics (kappa = 4, 4 frequencies):
                                                          # one root block and all its descendants
])
def test_full_resolution_window_against_reference(root, levels, pol, window, gpu, tmp_path):
    """Parity at a resolution the reference cannot hold in memory as a full frame: a coarse
    root image whose forced refinement region is refined twice makes the reference trace EXACTLY the pixel rays
    that a 4x finer full frame has inside that window (same u_ind/v_ind expression, camera.cpp:393-396 vs
    :476-479).  The CUDA path renders the fine full frame in one piece; its window must match the reference's
    deepest-level blocks pixel by pixel.  32^2 root -> 128^2 frame, 256^2 root -> the benchmark's 1024^2 frame and,
    with four levels and polarization, the 4096^2 Stokes frame of the north-star target."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    bs = 8
    fine = root * 2 ** levels
    over = {'camera_resolution': root, 'image_polarization': 'true' if pol else 'false', 'image_tau': 'false',
            'adaptive_frequency_num': 1, 'adaptive_max_level': levels, 'adaptive_block_size': bs, 'adaptive_num_regions': 1,
            'adaptive_region_1_level': levels, 'adaptive_region_1_x_min': window[0], 'adaptive_region_1_x_max': window[1],
            'adaptive_region_1_y_min': window[2], 'adaptive_region_1_y_max': window[3],
            'adaptive_val_frac': '-1.0', 'adaptive_abs_grad_frac': '-1.0', 'adaptive_rel_grad_frac': '-1.0',
            'adaptive_abs_lapl_frac': '-1.0', 'adaptive_rel_lapl_frac': '-1.0'}
    F = 1
    if pol == 'c4':
        over.update(C4_PHYSICS)
        F = 4
    case = Case(tmp_path / 'ref', 'adaptive.input', over)
    ref = case.run_reference(checkpoints=False)['npz']
    assert int(ref['adaptive_num_levels'][0]) == levels
    locs, blocks = ref['adaptive_block_locs_%d' % levels], ref['adaptive_I_nu_%d' % levels]
    assert len(locs) > 0 and blocks.shape[-2:] == (bs, bs)
    if F > 1:
        # every frequency: image slots 4 l + s of the fine frame against the reference's (F, blocks, bs, bs) arrays
        full_over = {k: v for k, v in over.items() if not k.startswith('adaptive_')}
        full_over.update({'camera_resolution': fine, 'adaptive_max_level': 0})
        cfg, ctx, image, _, _ = run_gpu_level0(Case(tmp_path / 'gpu', 'adaptive.input', full_over))
        assert ctx.polarized_stage_ms(0)['slab'] > 0   # the slab pipeline rendered it
        cut = lambda plane: np.stack([plane.reshape(fine, fine)[v * bs:(v + 1) * bs, u * bs:(u + 1) * bs] for v, u in locs])
        for l in range(F):
            mine = {name: cut(image[4 * l + s_ind]) for s_ind, name in enumerate(('I_nu', 'Q_nu', 'U_nu', 'V_nu'))}
            theirs = {name: ref['adaptive_%s_%d' % (name, levels)][l] for name in mine}
            assert rel_err(mine['I_nu'], theirs['I_nu']) <= PIXEL_TOL, 'frequency %d' % l
            assert flux_rel(mine['I_nu'], theirs['I_nu']) <= FLUX_TOL
            for k, v in stokes_err(mine, theirs).items():
                assert v <= PIXEL_TOL, 'frequency %d %s %.3e' % (l, k, v)
        ctx.close()
        return
    full_over = {k: v for k, v in over.items() if not k.startswith('adaptive_')}
    full_over.update({'camera_resolution': fine, 'adaptive_max_level': 0})
    full_case = Case(tmp_path / 'gpu', 'adaptive.input', full_over)
    cfg, ctx, image, _, _ = run_gpu_level0(full_case)
    frame = image[0].reshape(fine, fine)
    window = np.stack([frame[v * bs:(v + 1) * bs, u * bs:(u + 1) * bs] for v, u in locs])
    assert rel_err(window, blocks) <= PIXEL_TOL
    if pol:
        mine = {'I_nu': window}
        theirs = {'I_nu': blocks}
        for s_ind, name in ((1, 'Q_nu'), (2, 'U_nu'), (3, 'V_nu')):
            plane = image[s_ind].reshape(fine, fine)
            mine[name] = np.stack([plane[v * bs:(v + 1) * bs, u * bs:(u + 1) * bs] for v, u in locs])
            theirs[name] = ref['adaptive_%s_%d' % (name, levels)]
        for k, v in stokes_err(mine, theirs).items():
            assert v <= PIXEL_TOL, '%s %.3e' % (k, v)
    ctx.close()


@pytest.mark.parametrize('name', ['adaptive_value_32', 'adaptive_abs_grad_32', 'adaptive_rel_grad_32', 'adaptive_abs_lapl_32',
                                  'adaptive_rel_lapl_region_32'])
def test_golden_adaptive_each_criterion(name, gpu, tmp_path):
    """Each refinement criterion of the device-side EvaluateBlock on its own (value, absolute / relative gradient,
    absolute / relative Laplacian; the last with a forced region), two levels deep, through the drop-in executable path:
    block counts, block lists and per-level images against the unmodified reference's (tests/golden/make_golden.py
    ADAPT_CASES; radiation_adaptive.cpp:163-312)."""
    from golden.make_golden import ADAPT_CASES
    gold = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    case = Case(tmp_path, 'adaptive.input', ADAPT_CASES[name])
    npz, _ = case.run_gpu_file()
    assert np.array_equal(npz['adaptive_num_blocks'], gold['adaptive_num_blocks'])
    for k in gold:
        if k.startswith('adaptive_block_locs'):
            assert np.array_equal(npz[k], gold[k]), k
    for k in gold:
        if k == 'I_nu' or k.startswith('adaptive_I_nu'):
            assert npz[k].shape == gold[k].shape, k
            assert rel_err(npz[k], gold[k]) <= PIXEL_TOL, k


@pytest.mark.parametrize('name', ['formula_photon_12', 'formula_additive_12', 'formula_rk4_max_steps_12', 'simulation_rk4_16',
                                  'simulation_rk2_kerr_16', 'simulation_amr_16'])
def test_golden_unpolarized_more(name, gpu, tmp_path):
    """Fixtures added with the restatement's later coverage: photon-orbit termination, camera-frame frequency
    normalisation, the fixed-step integrators (one with rays flagged by the step limit) and the block search on a
    two-level AMR mesh -- flags, counts, every stored sample and the sampled cell indices bit-identical, images to
    tolerance (same checks as test_golden_unpolarized)."""
    test_golden_unpolarized(name, gpu, tmp_path)


def _cpu_case_names():
    from golden.make_golden import CPU_CASES
    return sorted(CPU_CASES)


@pytest.mark.parametrize('name', _cpu_case_names())
def test_golden_simulation_options(name, gpu, tmp_path):
    """The image-only fixtures that pin the restatement's option coverage, through the CUDA path: non-thermal electron
    mixes, electron temperature from energies, code_kappa, every geometric and cell-value cut (each checked to bite),
    value fallbacks outside the grid (all auxiliary images) and inter-block interpolation on AMR / multi-block meshes.
    Images within the per-pixel and flux tolerances of the unmodified reference's."""
    from golden.make_golden import CPU_CASES
    over = dict(CPU_CASES[name])
    mock = over.pop('_mock', None) or (dict(entropy=True) if over.get('plasma_model') == 'code_kappa' else None)
    gold = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    case = Case(tmp_path, 'simulation.input', over, mock=mock)
    cfg, ctx, image, _, _ = run_gpu_level0(case)
    check_images(image_arrays(case, image, cfg.resolution), gold, name)
    ctx.close()


@pytest.mark.parametrize('name', ['cpu_slow_light_blend_12', 'cpu_slow_light_nearest_slice_12',
                                  'cpu_slow_light_blend_nearest_cell_12'])
def test_golden_slow_light(name, gpu, tmp_path):
    """First image of a slow-light run over a 12-file series (nearest slice, blended slices, blended slices of the
    nearest cell) through the drop-in executable path against the unmodified reference's image."""
    from golden.make_golden import SLOW_CASES, slow_light_setup
    from harness import write_input
    d = str(tmp_path)
    kv, _ = slow_light_setup(d, SLOW_CASES[name])
    path = os.path.join(d, 'gpu.input')
    write_input(path, kv)
    bl.run_input_file(path)
    mine = np.load(os.path.join(d, 'img.000.npz'))['I_nu']
    ref = np.load(os.path.join(GOLDEN, name + '.npz'))['I_nu']
    assert rel_err(mine, ref) <= PIXEL_TOL
    assert flux_rel(mine, ref) <= FLUX_TOL


def _render_polarized(case, env, tile_rays=0):
    """One level-0 polarized image with the given BL_POL_* environment (read by bl_create)."""
    saved = {k: os.environ.get(k) for k in ('BL_POL_FUSED', 'BL_POL_SLAB')}
    try:
        for k in saved:
            os.environ.pop(k, None)
        os.environ.update(env)
        cfg = case.config(tile_rays=tile_rays)
        ctx = bl.Context(cfg)
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    ctx.upload_grid(case.grid_arrays())
    pos, dirs, fac = cfg.camera_root()
    ctx.trace_level(0, pos, dirs, fac)
    image, _, stats = ctx.radiate_level(0)
    stages = ctx.polarized_stage_ms(0)
    ctx.close()
    return image, stats, stages


@pytest.mark.parametrize('over', [
    dict(C4_PHYSICS, camera_resolution=40),
    {'camera_resolution': 40, 'image_polarization': 'true', 'image_tau': 'true', 'image_lambda': 'true', 'image_emission': 'true'},
    {'camera_resolution': 32, 'image_polarization': 'true', 'image_rotation_split': 'true', 'image_num_frequencies': 2,
     'image_frequency_start': '2.3e11', 'image_frequency_end': '4.6e11', 'image_frequency_spacing': 'log',
     'plasma_power_frac': '0.5', 'plasma_p': '3.0', 'plasma_gamma_min': '4.0', 'plasma_gamma_max': '1000.0',
     'plasma_kappa_frac': '0.25', 'plasma_kappa': '3.7', 'plasma_w': '1.5', 'simulation_a': '0.9'},
])
def test_polarized_pipeline_matches_fused_kernel(over, gpu, tmp_path):
    """The polarized pipeline (sampling | geometry | coefficients | transfer over slabs, radiate_pol_split.cu) and the
    single fused kernel evaluate the same formulas: the images must agree to rounding, for every slab length (a slab
    boundary re-derives the previous sample's frame from a halo sample) and when the level is traced in waves."""
    case = Case(tmp_path, 'simulation.input', over)
    fused, st_f, stages_f = _render_polarized(case, {'BL_POL_FUSED': '1'})
    assert stages_f['slab'] == 0
    for env, tile in (({}, 0), ({'BL_POL_SLAB': '16'}, 0), ({'BL_POL_SLAB': '7'}, 0), ({'BL_POL_SLAB': '2000'}, 0),
                      ({'BL_POL_SLAB': '32'}, 512)):
        image, st, stages = _render_polarized(case, env, tile_rays=tile)
        assert stages['slab'] > 0, 'the pipeline did not run'
        assert st['num_samples'] == st_f['num_samples']
        assert np.array_equal(np.isnan(image), np.isnan(fused))
        light = 4 * int(over.get('image_num_frequencies', 1))
        scale = np.nanmax(np.abs(fused[0:light:4]))   # brightest Stokes I
        err = np.nanmax(np.abs(image[:light] - fused[:light])) / scale
        print('pipeline vs fused kernel, slab %s tile %d: %.3e of the peak' % (env, tile, err))
        # same formulas, but each kernel is contracted into FMAs its own way: rounding-level differences, which the
        # recurrence over ~10^3 samples carries along (the parity bound against the reference is 1e-6)
        assert err <= 1e-10, 'slab %s tile %d: Stokes images differ by %.3e of the peak' % (env, tile, err)
        if image.shape[0] > light:
            assert rel_err(image[light:], fused[light:]) <= 1e-10


@pytest.mark.parametrize('base,over,mock', [
    ('simulation.input', dict(C4_PHYSICS, camera_resolution=30), None),
    ('simulation.input', {'camera_resolution': 24, 'image_time': 'true', 'image_tau': 'true', 'image_crossings': 'true'}, {'blocks': (7, 2, 2)}),
    ('adaptive.input', {'camera_resolution': 32, 'adaptive_max_level': 2, 'adaptive_num_regions': 1, 'adaptive_region_1_level': 2,
                        'adaptive_region_1_x_min': '-4', 'adaptive_region_1_x_max': '4', 'adaptive_region_1_y_min': '-4',
                        'adaptive_region_1_y_max': '4'}, None),
    ('render.input', {'camera_resolution': 24}, None),
    ('formula.input', {'camera_resolution': 20}, None),
])
def test_multi_device_driver_is_bitwise_the_single_device(base, over, mock, gpu, tmp_path):
    """blh_run_input_file_devices: one context and one host thread per listed device, rows (adaptive: refinement blocks,
    the root level included) dealt round-robin, parts scattered into the frame.  Every array of the written npz must be
    bit for bit the single-device run's -- here with three contexts sharing the one GPU of the test box (the driver does
    not care whether the ordinals differ)."""
    case = Case(tmp_path, base, over, mock=mock)
    one, t1 = case.run_gpu_file(tag='one', devices=[0])
    three, t3 = case.run_gpu_file(tag='three', devices=[0, 0, 0])
    assert t3['devices'] == 3 and t1['rays'] == t3['rays'] and t1['samples'] == t3['samples']
    assert sorted(one) == sorted(three)
    for k in one:
        assert one[k].shape == three[k].shape, k
        assert np.array_equal(one[k], three[k], equal_nan=True), k


def test_sampled_cell_indices_exact_at_scale(gpu, tmp_path):
    """north_star asks for bit-exact sampled cell indices; the radiation kernels locate a sample with fused arithmetic
    (r from one rsqrt, CUDA acos / atan2), so exactness is an empirical property: an index can only differ where a
    coordinate rounds to the other side of a cell face.  Measured here on > 10^8 samples: three 256^2 frames (single
    block trilinear; 56 blocks, Kerr a = 0.9, inclined, nearest; two-level AMR mesh, tilted camera, trilinear) against
    the reference's sampling checkpoint.  Any mismatch fails the test and is printed with its count."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    configs = [
        ({'camera_resolution': 256}, None),
        ({'camera_resolution': 256, 'simulation_interp': 'false', 'simulation_a': '0.9', 'camera_th': '70.0', 'camera_ph': '25.0'},
         {'blocks': (7, 2, 4)}),
        ({'camera_resolution': 256, 'camera_th': '60.0', 'camera_rotation': '20.0'}, dict(AMR_MOCK, refine=_amr_refine)),
    ]
    total = bad_inds = bad_flags = 0
    worst_frac = 0.0
    for n, (over, mock) in enumerate(configs):
        case = Case(tmp_path / str(n), 'simulation.input', over, mock=mock)
        rs = case.run_reference(checkpoints='sample')['samp']
        cfg, ctx, image, _, stats = run_gpu_level0(case, taps=True)
        interp = case.kv['simulation_interp'] == 'true'
        t = ctx.download_sample_inds(0, interp=interp)
        num = ctx.download_samples(0, arrays=False)['num']
        ctx.close()
        S = t['nan'].shape[1]
        assert rs['sample_nan'].shape == t['nan'].shape
        mask = np.arange(S)[None, :] < num[:, None]
        bad_flags += int(np.count_nonzero((t['nan'] != rs['sample_nan'])[mask])) + int(np.count_nonzero((t['fallback'] != rs['sample_fallback'])[mask]))
        valid = mask & (t['cut'] == 0) & (t['nan'] == 0) & (t['fallback'] == 0)
        differ = np.any(t['inds'] != rs['sample_inds'], axis=2) & valid
        total += int(np.count_nonzero(valid))
        bad_inds += int(np.count_nonzero(differ))
        if interp:
            worst_frac = max(worst_frac, float(np.max(np.abs(t['fracs'] - rs['sample_fracs'])[valid & ~differ])))
        del t, rs
    print('sampled cell indices: %d samples compared, %d index mismatches, %d flag mismatches, worst fraction difference %.2e'
          % (total, bad_inds, bad_flags, worst_frac))
    assert total > 100_000_000
    assert bad_inds == 0 and bad_flags == 0
    assert worst_frac < 1e-9


def test_cks_polarized_faint_pixels_are_roundoff_limited(gpu, tmp_path):
    """On the Cartesian Kerr-Schild box the polarized images of the two codes differ by more than 1e-6 in pixels fainter
    than 1e-6 of the peak (test_live_reference_cartesian_kerr_schild compares the others).  Those rays cross 3-unit cells
    of the dense midplane with optical depths of tens per step, where the closed-form step (polarized.cpp:598-653, :656-779)
    subtracts exponentially large terms.  Shown here: the transfer recurrence is re-evaluated with numpy on the very
    per-sample inputs the CUDA pipeline used (transport matrix, step, coefficients), once in float64 and once in 80-bit
    long double (tests/stokes_transfer.py).  (1) The float64 evaluation reproduces the CUDA image wherever the formulas are
    well conditioned.  (2) kappa = |float64 - long double| / |long double| is the round-off the closed forms amplify for
    that pixel; every pixel -- faint ones included -- agrees with the reference within max(1e-6, 100 kappa), and wherever the
    two codes differ by more than 1e-6, kappa itself exceeds 1e-8: the difference is round-off of the reference's own
    formula, not a modelling difference."""
    if not os.path.exists(REF_BIN):
        pytest.skip('oracle/_ref/blacklight not present')
    import stokes_transfer
    over = {'camera_resolution': 24, 'image_polarization': 'true', 'simulation_coord': 'cks', 'simulation_a': '0.5'}
    case = Case(tmp_path, 'simulation.input', over, mock=dict(blocks=(2, 2, 2), cks=dict(n=32)))
    ref = case.run_reference(checkpoints=False)['npz']
    saved = os.environ.get('BL_POL_SLAB')
    os.environ['BL_POL_SLAB'] = '2000'          # one slab = the whole ray: the scratch then holds every sample
    try:
        cfg = case.config()
        ctx = bl.Context(cfg)
    finally:
        os.environ.pop('BL_POL_SLAB', None)
        if saved is not None:
            os.environ['BL_POL_SLAB'] = saved
    ctx.upload_grid(case.grid_arrays())
    pos, dirs, fac = cfg.camera_root()
    ctx.trace_level(0, pos, dirs, fac)
    image, _, _ = ctx.radiate_level(0)
    assert ctx.polarized_stage_ms(0)['slab'] >= 2000
    scratch, cam_map = ctx.polarized_scratch(0)
    num = ctx.download_samples(0, arrays=False)['num']
    ctx.close()
    nu = float(case.kv['image_frequency'])
    x_unit = 1.32712440018e26 * float(case.kv['simulation_m_msun']) / 2.99792458e10 ** 2
    dl_factor = x_unit / fac / nu
    img64 = stokes_transfer.transfer(scratch, cam_map, num, dl_factor, 0, nu, np.float64)
    imgld = stokes_transfer.transfer(scratch, cam_map, num, dl_factor, 0, nu, np.longdouble)
    I_gpu, I_ref = image[0], ref['I_nu'].ravel()
    peak = np.nanmax(I_ref)
    with np.errstate(all='ignore'):
        kappa = np.abs(img64[0] - imgld[0]).astype(np.float64) / np.maximum(np.abs(imgld[0]).astype(np.float64), 1e-300)
        err_restated = np.abs(I_gpu - img64[0]) / np.maximum(np.abs(img64[0]), 1e-12 * peak)
        err_ref = np.abs(I_gpu - I_ref) / np.maximum(np.abs(I_ref), 1e-12 * peak)
    lit = I_ref > 0
    off = lit & (err_ref > PIXEL_TOL)
    print('cks polarized: %d lit pixels, %d differ from the reference by > 1e-6 (I / peak of those: %.1e ... %.1e); '
          'kappa there %.1e ... %.1e, err / kappa %.1f ... %.1f; float64 restatement vs CUDA: max %.1e (well conditioned: %.1e)'
          % (lit.sum(), off.sum(), (I_ref[off] / peak).min() if off.any() else 0, (I_ref[off] / peak).max() if off.any() else 0,
             kappa[off].min() if off.any() else 0, kappa[off].max() if off.any() else 0,
             (err_ref[off] / kappa[off]).min() if off.any() else 0, (err_ref[off] / kappa[off]).max() if off.any() else 0,
             err_restated[lit].max(), err_restated[lit & (kappa < 1e-12)].max()))
    assert np.all(err_restated[lit] <= np.maximum(1e-10, 100.0 * kappa[lit]))
    assert np.all(err_ref[lit] <= np.maximum(PIXEL_TOL, 100.0 * kappa[lit]))
    assert np.all(kappa[off] > 1e-8)
    assert np.all(I_ref[off] < 1e-5 * peak)


def _bits(a):
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


DEVICE_CAMERAS = [
    ('formula.input', {'camera_resolution': 48}),                                                  # plane, a = 0.9, the bench's C1 camera
    ('formula.input', {'camera_resolution': 40, 'camera_type': 'pinhole', 'camera_th': '30.0', 'camera_ph': '50.0',
                       'camera_rotation': '25.0', 'camera_urn': '0.1', 'camera_uthn': '0.05', 'camera_uphn': '-0.02',
                       'camera_k_th': '0.1', 'camera_k_ph': '-0.05', 'image_normalization': 'infinity'}),
    ('formula.input', {'camera_resolution': 32, 'camera_th': '0.0'}),                              # pole camera, plane
    ('formula.input', {'camera_resolution': 32, 'camera_th': '180.0', 'camera_type': 'pinhole'}),  # south pole, pinhole
    ('formula.input', {'camera_resolution': 32, 'ray_flat': 'true', 'camera_rotation': '10.0'}),
    ('formula.input', {'camera_resolution': 32, 'formula_spin': '0.0', 'camera_th': '90.0'}),      # a = 0: hypot(x, 0); equatorial
    ('simulation.input', {'camera_resolution': 64}),                                               # the bench's C2 / C4 camera
]


@pytest.mark.parametrize('base,over', DEVICE_CAMERAS)
def test_device_camera_is_bitwise_the_host_camera(base, over, gpu, tmp_path):
    """bl_trace_level_pixels (csrc/camera_kernel.cu): positions, covariant momenta and frequency factors of the rays as the
    device generates them, against the host camera (csrc/host/camera.cpp, itself bit-identical to the reference's
    checkpoint: test_cpu_host.py::test_camera_bit_exact) -- every array element bit for bit, whole raster and a row
    subset, plane / pinhole / pole / flat-space / moving / rotated cameras, both frequency normalisations."""
    case = Case(tmp_path, base, over)
    cfg = case.config()
    ctx = bl.Context(cfg)
    res = cfg.resolution
    pos, dirs, fac = cfg.camera_root()
    st = ctx.trace_level_pixels(0)
    dpos, ddir, dfac = ctx.download_camera(0)
    assert np.array_equal(_bits(dpos), _bits(pos)) and np.array_equal(_bits(ddir), _bits(dirs)) and np.array_equal(_bits(dfac), _bits(fac))
    num_d = ctx.download_samples(0, arrays=False)
    # the geodesics integrated from them are the host path's
    st_h = ctx.trace_level(0, pos, dirs, fac)
    num_h = ctx.download_samples(0, arrays=False)
    assert st['num_samples'] == st_h['num_samples'] and np.array_equal(num_d['num'], num_h['num']) and np.array_equal(num_d['flags'], num_h['flags'])
    rows = np.array([res - 1, 3, 0, res // 2, 17 % res], np.int64)
    rpos, rdir, rfac = cfg.camera_rows(rows)
    ctx.trace_level_pixels(0, rows=rows)
    dpos, ddir, dfac = ctx.download_camera(0)
    assert np.array_equal(_bits(dpos), _bits(rpos)) and np.array_equal(_bits(ddir), _bits(rdir)) and np.array_equal(_bits(dfac), _bits(rfac))
    ctx.close()


def test_device_camera_blocks_and_large_raster(gpu, tmp_path):
    """Refined-level blocks (effective resolution res * 2^level, block-major) and a 1024^2 raster (10^6 pixels) bit for bit
    the host camera's; bad unit lists are refused."""
    over = {'camera_resolution': 32, 'adaptive_max_level': 3, 'adaptive_block_size': 8, 'camera_th': '70.0', 'camera_rotation': '5.0'}
    case = Case(tmp_path / 'blocks', 'adaptive.input', over)
    cfg = case.config()
    ctx = bl.Context(cfg)
    rng = np.random.default_rng(7)
    for level in (0, 1, 2, 3):
        nb = (32 << level) // 8
        locs = np.stack([rng.integers(0, nb, 23), rng.integers(0, nb, 23)], 1).astype(np.int32)
        locs[0], locs[1] = (0, 0), (nb - 1, nb - 1)
        pos, dirs, fac = cfg.camera_blocks(level, locs)
        ctx.trace_level_pixels(level, blocks=locs)
        dpos, ddir, dfac = ctx.download_camera(level)
        assert np.array_equal(_bits(dpos), _bits(pos)) and np.array_equal(_bits(ddir), _bits(dirs)) and np.array_equal(_bits(dfac), _bits(fac)), level
    with pytest.raises(bl.BlacklightError):
        ctx.trace_level_pixels(1, blocks=np.array([[0, 8]], np.int32))     # 8 blocks per side at level 1
    with pytest.raises(bl.BlacklightError):
        ctx.trace_level_pixels(0, rows=np.array([32], np.int32))
    ctx.close()
    big = Case(tmp_path / 'big', 'simulation.input', {'camera_resolution': 1024, 'camera_type': 'pinhole', 'ray_max_steps': 40})
    cfg = big.config()
    ctx = bl.Context(cfg)
    pos, dirs, fac = cfg.camera_root()
    ctx.trace_level_pixels(0)
    dpos, ddir, dfac = ctx.download_camera(0)
    assert np.array_equal(_bits(dpos), _bits(pos)) and np.array_equal(_bits(ddir), _bits(dirs)) and np.array_equal(_bits(dfac), _bits(fac))
    ctx.close()


@pytest.mark.parametrize('base,over', [
    ('simulation.input', {'camera_resolution': 24, 'image_polarization': 'true', 'output_camera': 'true'}),
    ('adaptive.input', {'camera_resolution': 32, 'adaptive_max_level': 2, 'output_camera': 'true', 'adaptive_num_regions': 1,
                        'adaptive_region_1_level': 2, 'adaptive_region_1_x_min': '-4', 'adaptive_region_1_x_max': '4',
                        'adaptive_region_1_y_min': '-4', 'adaptive_region_1_y_max': '4'}),
    ('formula.input', {'camera_resolution': 20, 'camera_type': 'pinhole', 'output_camera': 'true'}),
])
def test_drop_in_with_device_camera_matches_host_camera(base, over, gpu, tmp_path):
    """The drop-in driver generates the camera pixels on the device by default; BLACKLIGHT_HOST_CAMERA=1 keeps round 1's
    host arrays + upload.  Both runs (one device and three contexts) must write identical files, camera arrays included."""
    case = Case(tmp_path, base, over)
    dev, _ = case.run_gpu_file(tag='dev', devices=[0])
    dev3, _ = case.run_gpu_file(tag='dev3', devices=[0, 0, 0])
    os.environ['BLACKLIGHT_HOST_CAMERA'] = '1'
    try:
        host, _ = case.run_gpu_file(tag='host', devices=[0])
    finally:
        del os.environ['BLACKLIGHT_HOST_CAMERA']
    assert sorted(dev) == sorted(host) == sorted(dev3)
    for k in host:
        assert np.array_equal(dev[k], host[k], equal_nan=True), k
        assert np.array_equal(dev3[k], host[k], equal_nan=True), k


@pytest.mark.parametrize('base,over,tile', [
    ('simulation.input', {'camera_resolution': 48}, 0),
    ('simulation.input', dict(C4_PHYSICS, camera_resolution=40), 0),
    ('simulation.input', dict(C4_PHYSICS, camera_resolution=40), 512),                               # waves
    ('simulation.input', {'camera_resolution': 32, 'image_polarization': 'true', 'image_time': 'true', 'image_tau': 'true',
                          'image_crossings': 'true'}, 0),                                            # fused polarized kernel
    ('formula.input', {'camera_resolution': 40}, 0),
    ('true_color.input', {'camera_resolution': 24}, 384),
])
def test_ray_ordering_does_not_change_a_bit(base, over, tile, gpu, tmp_path):
    """The radiation kernels take their rays from the wave's list sorted by length (csrc/ray_order.cu) and the polarized
    pipeline launches each slab over the rays still alive in it, addressing its scratch by list position.  A ray's own
    arithmetic is untouched: every image array must equal, bit for bit, the one rendered in index order (BL_RAY_ORDER=0),
    for resident levels and waves, the slab pipeline (three slab lengths) and the fused kernels."""
    case = Case(tmp_path, base, over)

    def render(env):
        saved = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            cfg = case.config(tile_rays=tile)
            ctx = bl.Context(cfg)
        finally:
            for k, v in saved.items():
                os.environ.pop(k, None)
                if v is not None:
                    os.environ[k] = v
        if case.sim:
            ctx.upload_grid(case.grid_arrays())
        ctx.trace_level_pixels(0)
        image, _, st = ctx.radiate_level(0)
        ctx.close()
        return image, st

    plain, st0 = render({'BL_RAY_ORDER': '0'})
    for env in ({}, {'BL_POL_SLAB': '7'}, {'BL_POL_SLAB': '200'}):
        image, st = render(env)
        assert st['num_samples'] == st0['num_samples']
        assert np.array_equal(_bits(image), _bits(plain)), env
