#!/usr/bin/env python
"""Regenerate the golden fixtures in this directory by running the UNMODIFIED reference
(oracle/_ref/blacklight, built from /root/reference by oracle/Makefile) on small cases.

Run where the reference is built:  python tests/golden/make_golden.py
Each fixture holds what the parity tests compare: sample_flags, sample_num, the image arrays, a CRC-32 of
the masked sample_inds (simulation cases) and full samples of a few rays.
"""
import os
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from harness import Case  # noqa: E402

POL = {'image_polarization': 'true', 'image_rotation_split': 'false'}
KAPPA = {'plasma_kappa_frac': '1.0', 'plasma_kappa': '4.0', 'plasma_w': '1.0'}
MULTI = {'image_num_frequencies': 3, 'image_frequency_start': '8.6e10', 'image_frequency_end': '3.45e11',
         'image_frequency_spacing': 'log'}
AUX = {'image_time': 'true', 'image_length': 'true', 'image_lambda': 'true', 'image_emission': 'true',
       'image_tau': 'true', 'image_crossings': 'true'}
AUX_SIM = dict(AUX, image_lambda_ave='true', image_emission_ave='true', image_tau_int='true')

def amr_refine(bi, bj, bk):
    """Root blocks of the two-level mock mesh that are replaced by their eight children."""
    return bi == 0 and bj == 1


CASES = {
    'formula_16': ('formula.input', {'camera_resolution': 16}, None),
    'formula_aux_12': ('formula.input', dict(AUX, camera_resolution=12), None),
    'simulation_32': ('simulation.input', {'camera_resolution': 32}, None),
    'simulation_nearest_24': ('simulation.input', {'camera_resolution': 24, 'simulation_interp': 'false'}, None),
    'simulation_blocks_24': ('simulation.input', {'camera_resolution': 24}, {'blocks': (7, 2, 4)}),
    'simulation_aux_16': ('simulation.input', dict(AUX_SIM, camera_resolution=16), None),
    'simulation_kerr_24': ('simulation.input', {'camera_resolution': 24, 'simulation_a': '0.9', 'camera_th': '80.0'}, None),
    'polarized_thermal_16': ('simulation.input', dict(POL, camera_resolution=16), None),
    'polarized_kappa_multi_12': ('simulation.input', dict(POL, **KAPPA, **MULTI, camera_resolution=12), None),
    'adaptive_32': ('adaptive.input', {}, None),
    'render_32': ('render.input', {'camera_resolution': 32}, None),
    'true_color_16': ('true_color.input', {'camera_resolution': 16}, None),
    'simulation_rk4_16': ('simulation.input', {'camera_resolution': 16, 'ray_integrator': 'rk4', 'ray_step': '0.02'}, None),
    'simulation_rk2_kerr_16': ('simulation.input', {'camera_resolution': 16, 'ray_integrator': 'rk2', 'ray_step': '0.02',
                                                    'simulation_a': '0.9', 'camera_th': '80.0'}, None),
    'formula_rk4_max_steps_12': ('formula.input', {'camera_resolution': 12, 'ray_integrator': 'rk4', 'ray_step': '0.02',
                                                   'ray_max_steps': 600}, None),
    'formula_photon_12': ('formula.input', {'camera_resolution': 12, 'ray_terminate': 'photon', 'formula_spin': '0.7'}, None),
    'formula_additive_12': ('formula.input', {'camera_resolution': 12, 'ray_terminate': 'additive', 'ray_factor': '0.3',
                                              'image_normalization': 'camera'}, None),
    'simulation_amr_16': ('simulation.input', {'camera_resolution': 16},
                          {'blocks': (2, 2, 4), 'n_r': 32, 'n_th': 16, 'n_ph': 32, 'refine': amr_refine}),
    'formula_pinhole_pole_12': ('formula.input', {'camera_resolution': 12, 'camera_type': 'pinhole', 'camera_th': '180.0',
                                                  'camera_r': '100.0', 'camera_urn': '-0.05', 'camera_rotation': '25.0'}, None),
}

# image-only fixtures for the CPU-only checks of the restatement (tests/test_cpu_oracle.py)
PLAW = {'plasma_power_frac': '0.4', 'plasma_p': '3.0', 'plasma_gamma_min': '4.0', 'plasma_gamma_max': '1000.0'}
CPU_CASES = {
    'cpu_simulation_power_law_16': dict(PLAW, camera_resolution='16'),
    'cpu_simulation_mixed_electrons_16': dict(PLAW, plasma_power_frac='0.2', plasma_kappa_frac='0.5', plasma_kappa='4.0',
                                              plasma_w='1.0', camera_resolution='16'),
    'cpu_simulation_energy_temperature_16': {'plasma_use_p': 'false', 'plasma_gamma': '1.5',
                                             'plasma_gamma_i': '1.6666666666666667',
                                             'plasma_gamma_e': '1.3333333333333333', 'camera_resolution': '16'},
    'cpu_simulation_cuts_a_16': {'cut_omit_near': 'true', 'cut_omit_in': '3.0', 'cut_midplane_theta': '30.0',
                                 'cut_rho_min': '1.0e-18', 'camera_resolution': '16'},
    'cpu_simulation_cuts_b_16': {'cut_omit_far': 'true', 'cut_omit_out': '30.0', 'cut_midplane_z': '6.0', 'cut_plane': 'true',
                                 'cut_plane_origin': '1.0,0.0,0.5', 'cut_plane_normal': '0.2,1.0,0.1', 'cut_sigma_max': '-1.0',
                                 'cut_beta_inverse_max': '5.0', 'cut_theta_e_max': '50.0', 'camera_resolution': '16'},
    'cpu_simulation_cuts_c_16': {'cut_midplane_theta': '-5.0', 'cut_rho_min': '2.5e-17', 'cut_theta_e_max': '5.6',
                                 'cut_b_max': '42.0', 'camera_resolution': '16'},
    'cpu_simulation_cuts_d_16': {'cut_midplane_z': '-0.3', 'cut_n_e_max': '2.3e7', 'cut_p_gas_min': '450.0',
                                 'cut_sigma_min': '3.2e-3', 'cut_beta_inverse_max': '7.5e-2', 'camera_resolution': '16'},
    'cpu_simulation_cuts_e_16': {'cut_rho_max': '4.0e-17', 'cut_theta_e_min': '3.2', 'cut_b_min': '30.0', 'cut_n_e_min': '1.5e7',
                                 'cut_p_gas_max': '1000.0', 'cut_beta_inverse_min': '6.0e-2', 'cut_sigma_max': '4.5e-3',
                                 'camera_resolution': '16'},
    # all 27 auxiliary images kept: the fallback values only show in the cell-value averages
    'cpu_simulation_fallback_values_16': dict(AUX_SIM, fallback_nan='false', fallback_rho='1.0e-6', fallback_pgas='1.0e-8',
                                              camera_r='80.0', camera_width='60.0', camera_resolution='16'),
    # inter-block trilinear anchors on a two-level mesh, on a single-level multi-block one, and with nearest-cell fallbacks
    'cpu_simulation_block_interp_amr_16': {'simulation_block_interp': 'true', 'camera_resolution': '16',
                                           '_mock': {'blocks': (2, 2, 4), 'n_r': 32, 'n_th': 16, 'n_ph': 32, 'refine': amr_refine}},
    'cpu_simulation_block_interp_amr_tilted_16': {'simulation_block_interp': 'true', 'camera_resolution': '16', 'camera_th': '35.0',
                                                  'camera_ph': '100.0',
                                                  '_mock': {'blocks': (2, 2, 4), 'n_r': 32, 'n_th': 16, 'n_ph': 32, 'refine': amr_refine}},
    'cpu_simulation_block_interp_blocks_16': {'simulation_block_interp': 'true', 'camera_resolution': '16',
                                              '_mock': {'blocks': (2, 2, 4), 'n_r': 32, 'n_th': 16, 'n_ph': 32}},
    'cpu_simulation_fallback_entropy_16': dict(AUX_SIM, fallback_nan='false', fallback_rho='1.0e-6', fallback_pgas='1.0e-8',
                                               fallback_kappa='1.0e8', plasma_model='code_kappa', simulation_kappa_name='r0',
                                               camera_r='80.0', camera_width='60.0', camera_resolution='16'),
    'cpu_simulation_code_kappa_16': {'plasma_model': 'code_kappa', 'simulation_kappa_name': 'r0', 'camera_resolution': '16'},
    'cpu_simulation_code_kappa_nearest_16': {'plasma_model': 'code_kappa', 'simulation_kappa_name': 'r0',
                                             'simulation_interp': 'false', 'camera_resolution': '16'},
}

# slow light: a 12-file time series of small mock snapshots; first image only (snapshot time = slow_t_start)
SLOW_CASES = {
    'cpu_slow_light_blend_12': {'slow_interp': 'true'},
    'cpu_slow_light_nearest_slice_12': {'slow_interp': 'false'},
    'cpu_slow_light_blend_nearest_cell_12': {'slow_interp': 'true', 'simulation_interp': 'false'},
}
SLOW_FILES, SLOW_DT_FILE = 12, 25.0


def slow_light_setup(d, over):
    """Writes the series under d/data and returns (input keys, list of per-file grids).  Shared with the CPU test."""
    from harness import load_input
    from blacklight_b200 import mock_snapshot as ms
    os.makedirs(os.path.join(d, 'data'), exist_ok=True)
    grids = []
    for n in range(SLOW_FILES):
        grid = ms.to_blocks(ms.mock_fields(n_r=32, n_th=16, n_ph=32, pert_amp=0.1 + 0.05 * n), (1, 1, 2))
        ms.write_athdf(os.path.join(d, 'data', 'mock.%05d.athdf' % n), grid, time=SLOW_DT_FILE * n)
        grids.append(grid)
    kv = load_input('simulation.input')
    kv.update({'camera_resolution': '12', 'simulation_file': os.path.join(d, 'data', 'mock.{05d}.athdf'),
               'simulation_multiple': 'true', 'simulation_start': '0', 'simulation_end': str(SLOW_FILES - 1),
               'slow_light_on': 'true', 'slow_chunk_size': '8', 'slow_t_start': '200.0', 'slow_dt': '20.0',
               'slow_num_images': '1', 'slow_offset': '0', 'output_file': os.path.join(d, 'img.{03d}.npz'),
               'num_threads': '8'})
    kv.update(over)
    return kv, grids


def make_slow(only):
    import subprocess
    from harness import REF_BIN, write_input
    for name, over in SLOW_CASES.items():
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as d:
            kv, _ = slow_light_setup(d, over)
            path = os.path.join(d, 'ref.input')
            write_input(path, kv)
            proc = subprocess.run([REF_BIN, path], cwd=d, capture_output=True, text=True, timeout=3600)
            assert proc.returncode == 0 and 'Calculation completed' in proc.stdout, proc.stdout + proc.stderr
            image = np.load(os.path.join(d, 'img.000.npz'))['I_nu']
            np.savez_compressed(os.path.join(os.environ.get('GOLDEN_OUT', HERE), name + '.npz'), I_nu=image)
            print(name, image.shape, float(np.nanmax(image)))

# adaptive refinement: one criterion per case, two levels, unpolarized; block lists and per-level images only
ADAPT_OFF = {'adaptive_rel_lapl_frac': '-1.0', 'adaptive_max_level': '2', 'image_polarization': 'false'}
ADAPT_CASES = {
    'adaptive_value_32': dict(ADAPT_OFF, adaptive_val_cut='4.0e-5', adaptive_val_frac='0.3'),
    'adaptive_abs_grad_32': dict(ADAPT_OFF, adaptive_abs_grad_cut='2.0e-5', adaptive_abs_grad_frac='0.2'),
    'adaptive_rel_grad_32': dict(ADAPT_OFF, adaptive_rel_grad_cut='0.5', adaptive_rel_grad_frac='0.25'),
    'adaptive_abs_lapl_32': dict(ADAPT_OFF, adaptive_abs_lapl_cut='1.0e-5', adaptive_abs_lapl_frac='0.2'),
    'adaptive_rel_lapl_region_32': dict(ADAPT_OFF, adaptive_rel_lapl_frac='0.25', adaptive_num_regions='1',
                                        adaptive_region_1_level='2', adaptive_region_1_x_min='-11.0',
                                        adaptive_region_1_x_max='-7.0', adaptive_region_1_y_min='2.0',
                                        adaptive_region_1_y_max='9.0'),
}


def make_adaptive(only):
    for name, over in ADAPT_CASES.items():
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as d:
            ref = Case(d, 'adaptive.input', over, threads=8).run_reference(checkpoints=False)['npz']
            keep = {k: v for k, v in ref.items() if k == 'I_nu' or k.startswith('adaptive_num') or
                    k.startswith('adaptive_block_locs') or k.startswith('adaptive_I_nu')}
            np.savez_compressed(os.path.join(os.environ.get('GOLDEN_OUT', HERE), name + '.npz'), **keep)
            print(name, list(ref['adaptive_num_blocks']))


def masked_inds_crc(geo, samp, sim_interp):
    num = geo['sample_num']
    S = samp['sample_inds'].shape[1]
    valid = (np.arange(S)[None, :] < num[:, None]) & (samp['sample_nan'] == 0) & (samp['sample_fallback'] == 0)
    # geometric cuts are not in the checkpoint: recompute r > camera_r from positions is done by the test,
    # here we only keep samples whose stored index is inside the grid (unset entries are arbitrary)
    return valid


def main():
    only = sys.argv[1:]
    make_slow(only)
    make_adaptive(only)
    for name, over in CPU_CASES.items():
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as d:
            over = dict(over)
            mock = over.pop('_mock', None) or (dict(entropy=True) if over.get('plasma_model') == 'code_kappa' else None)
            ref = Case(d, 'simulation.input', over, mock=mock, threads=8).run_reference(checkpoints=False)
            keep = ref['npz'] if 'image_tau_int' in over else {'I_nu': ref['npz']['I_nu']}
            np.savez_compressed(os.path.join(os.environ.get('GOLDEN_OUT', HERE), name + '.npz'), **keep)
            print(name, ref['npz']['I_nu'].shape)
    for name, (base, over, mock) in CASES.items():
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as d:
            case = Case(d, base, over, mock=mock, threads=8)
            ref = case.run_reference(checkpoints=True)
            out = {k: v for k, v in ref['npz'].items()}
            geo = ref['geo']
            out['sample_flags'] = geo['sample_flags']
            out['sample_num'] = geo['sample_num']
            out['geodesic_num_steps'] = np.int32(geo['geodesic_num_steps'])
            for k in ('cam_x', 'u_con', 'u_cov', 'norm_con', 'norm_con_c', 'hor_con_c', 'vert_con_c', 'camera_pos',
                      'camera_dir', 'momentum_factors', 'image_frequencies'):
                out['geo_' + k] = geo[k]
            rays = np.unique(np.linspace(0, len(geo['sample_num']) - 1, 5).astype(int))
            out['probe_rays'] = rays
            for r in rays:
                n = geo['sample_num'][r]
                out['probe_pos_%d' % r] = geo['sample_pos'][r, :n]
                out['probe_dir_%d' % r] = geo['sample_dir'][r, :n]
                out['probe_len_%d' % r] = geo['sample_len'][r, :n]
            # checksum over every sample of every ray (positions, momenta, lengths) within sample_num
            S = geo['sample_pos'].shape[1]
            mask = np.arange(S)[None, :] < geo['sample_num'][:, None]
            out['samples_crc'] = np.uint32(zlib.crc32(geo['sample_pos'][mask].tobytes()
                                                       + geo['sample_dir'][mask].tobytes()
                                                       + geo['sample_len'][mask].tobytes()))
            if 'samp' in ref:
                samp = ref['samp']
                out['sample_nan_count'] = np.int64(samp['sample_nan'][mask].sum())
                # cut samples (r > camera_r for these configs) leave sample_inds unset: mask them by radius
                x = geo['sample_pos']
                a = float(case.kv['simulation_a'])
                rr2 = (x[..., 1:] ** 2).sum(-1)
                r2 = 0.5 * (rr2 - a * a + np.hypot(rr2 - a * a, 2.0 * a * x[..., 3]))
                cut = np.sqrt(r2) > float(case.kv['camera_r'])
                valid = mask & ~cut & (samp['sample_nan'] == 0) & (samp['sample_fallback'] == 0)
                out['valid_count'] = np.int64(valid.sum())
                out['inds_crc'] = np.uint32(zlib.crc32(np.ascontiguousarray(samp['sample_inds'][valid]).tobytes()))
            np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
            print(name, 'rays', len(geo['sample_num']), 'S', S, 'keys', len(out), ref['timers'])


if __name__ == '__main__':
    main()
