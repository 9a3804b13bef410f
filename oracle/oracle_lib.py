"""ctypes face of the plain-C restatement (oracle/blacklight_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_ref', 'libblacklight_oracle.so')


class Geo(ctypes.Structure):
    _fields_ = [('a', ctypes.c_double), ('flat', ctypes.c_int), ('camera_r', ctypes.c_double),
                ('r_terminate', ctypes.c_double), ('ray_step', ctypes.c_double), ('tol_abs', ctypes.c_double),
                ('tol_rel', ctypes.c_double), ('max_steps', ctypes.c_int), ('max_retries', ctypes.c_int)]


class Formula(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in ('a', 'camera_r', 'x_unit', 'r0', 'h', 'l0', 'q', 'nup', 'cn0', 'alpha',
                                               'abs_a', 'beta')] + [('fallback_nan', ctypes.c_int)]


class Sim(ctypes.Structure):
    _fields_ = [('a', ctypes.c_double), ('camera_r', ctypes.c_double), ('x_unit', ctypes.c_double)] + \
               [(n, ctypes.c_int) for n in ('n_b', 'n_k', 'n_j', 'n_i', 'interp', 'fallback_nan')] + \
               [(n, ctypes.c_double) for n in ('d_unit', 'mu', 'ne_ni', 'rat_low', 'rat_high', 'cut_sigma_max')] + \
               [('coord', ctypes.c_int)] + \
               [(n, ctypes.c_double) for n in ('power_frac', 'power_p', 'power_gamma_min', 'power_gamma_max',
                                               'kappa_frac', 'kappa', 'kappa_w')] + \
               [('flat', ctypes.c_int), ('cut_omit_in', ctypes.c_double), ('cut_omit_out', ctypes.c_double)] + \
               [('use_energy', ctypes.c_int), ('gamma', ctypes.c_double), ('gamma_i', ctypes.c_double), ('gamma_e', ctypes.c_double)] + \
               [('code_kappa', ctypes.c_int)] + \
               [('cut_omit_near', ctypes.c_int), ('cut_omit_far', ctypes.c_int), ('cut_plane', ctypes.c_int),
                ('cut_cam', ctypes.c_double * 3), ('cut_midplane_theta', ctypes.c_double), ('cut_midplane_z', ctypes.c_double),
                ('cut_plane_origin', ctypes.c_double * 3), ('cut_plane_normal', ctypes.c_double * 3),
                ('cut_val_min', ctypes.c_double * 7), ('cut_val_max', ctypes.c_double * 7)] + \
               [(n, ctypes.c_double) for n in ('fallback_rho', 'fallback_pgas', 'fallback_kappa')] + \
               [('n_t', ctypes.c_int), ('slow_interp', ctypes.c_int), ('snapshot_time', ctypes.c_double),
                ('times', ctypes.c_double * 64)] + \
               [('block_interp', ctypes.c_int), ('n_3_root', ctypes.c_int), ('levels', ctypes.c_void_p),
                ('locations', ctypes.c_void_p)]

CUT_VALUES = ('rho', 'n_e', 'p_gas', 'theta_e', 'b', 'sigma', 'beta_inverse')


class Feature(ctypes.Structure):
    _fields_ = [('image', ctypes.c_int), ('quantity', ctypes.c_int), ('type', ctypes.c_int)] + \
               [(n, ctypes.c_double) for n in ('min', 'max', 'tau_scale', 'thresh', 'opacity')] + [('xyz', ctypes.c_double * 3)]


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, 'blacklight_oracle.c')
        if os.path.exists(LIB) and os.path.getmtime(src) > os.path.getmtime(LIB):   # struct layouts must match this file
            import subprocess
            subprocess.run(['make', '-C', HERE, 'restatement'], check=True, capture_output=True)
        _lib = ctypes.CDLL(LIB)
        _lib.orc_trace_dp.restype = ctypes.c_int
        _lib.orc_trace_rk.restype = ctypes.c_int
    return _lib


def r_terminate(kv, a):
    r_h = 1.0 + np.sqrt(1.0 - a * a)
    mode = kv['ray_terminate']
    if mode == 'photon':
        return 2.0 * (1.0 + np.cos(2.0 / 3.0 * np.arccos(-abs(a))))
    return r_h * float(kv['ray_factor']) if mode == 'multiplicative' else r_h + float(kv['ray_factor'])


def trace(kv, a, cam_pos, cam_dir):
    g = Geo(a=a, flat=int(kv['ray_flat'] == 'true'), camera_r=float(kv['camera_r']), r_terminate=r_terminate(kv, a),
            ray_step=float(kv['ray_step']), tol_abs=float(kv['ray_tol_abs']), tol_rel=float(kv['ray_tol_rel']),
            max_steps=int(kv['ray_max_steps']), max_retries=int(kv['ray_max_retries']))
    n, cap = len(cam_pos), g.max_steps
    num, flags = np.zeros(n, np.int32), np.zeros(n, np.uint8)
    pos, dirs, length = np.zeros((n, cap, 4)), np.zeros((n, cap, 4)), np.zeros((n, cap))
    args = (ctypes.c_long(n), _p(np.ascontiguousarray(cam_pos)), _p(np.ascontiguousarray(cam_dir)), cap, _p(num),
            _p(flags), _p(pos), _p(dirs), _p(length))
    kind = kv.get('ray_integrator', 'dp')
    if kind == 'dp':
        steps = lib().orc_trace_dp(ctypes.byref(g), *args)
    else:
        steps = lib().orc_trace_rk(ctypes.byref(g), {'rk4': 4, 'rk2': 2}[kind], *args)
    return dict(num=num, flags=flags, pos=pos, dir=dirs, len=length, steps=steps, cap=cap)


C, GG_MSUN = 2.99792458e10, 1.32712440018e26


def formula_image(kv, s, mom, freqs):
    a = float(kv['formula_spin'])
    mass_msun = float(kv['formula_mass']) * C * C / GG_MSUN
    P = Formula(a=a, camera_r=float(kv['camera_r']), x_unit=GG_MSUN * mass_msun / (C * C), r0=float(kv['formula_r0']),
                h=float(kv['formula_h']), l0=float(kv['formula_l0']), q=float(kv['formula_q']), nup=float(kv['formula_nup']),
                cn0=float(kv['formula_cn0']), alpha=float(kv['formula_alpha']), abs_a=float(kv['formula_a']),
                beta=float(kv['formula_beta']), fallback_nan=int(kv['fallback_nan'] == 'true'))
    n, F = len(mom), len(freqs)
    image = np.zeros((F, n))
    freqs = np.ascontiguousarray(freqs, np.float64)
    lib().orc_formula_image(ctypes.byref(P), ctypes.c_long(n), s['cap'], _p(s['num']), _p(s['flags']), _p(s['pos']),
                            _p(s['dir']), _p(s['len']), _p(np.ascontiguousarray(mom)), F, _p(freqs), _p(image))
    return image


def formula_aux(kv, s, mom, freqs, camera_x):
    """Auxiliary images of the formula model: dict(time, length, crossings: (n); lambda, emission, tau: (F, n))."""
    a = float(kv['formula_spin'])
    mass_msun = float(kv['formula_mass']) * C * C / GG_MSUN
    P = Formula(a=a, camera_r=float(kv['camera_r']), x_unit=GG_MSUN * mass_msun / (C * C), r0=float(kv['formula_r0']),
                h=float(kv['formula_h']), l0=float(kv['formula_l0']), q=float(kv['formula_q']), nup=float(kv['formula_nup']),
                cn0=float(kv['formula_cn0']), alpha=float(kv['formula_alpha']), abs_a=float(kv['formula_a']),
                beta=float(kv['formula_beta']), fallback_nan=int(kv['fallback_nan'] == 'true'))
    n, F = len(mom), len(freqs)
    out = {k: np.zeros(n) for k in ('time', 'length', 'crossings')}
    out.update({k: np.zeros((F, n)) for k in ('lambda', 'emission', 'tau')})
    freqs = np.ascontiguousarray(freqs, np.float64)
    cam = np.ascontiguousarray(camera_x, np.float64)
    lib().orc_formula_aux(ctypes.byref(P), ctypes.c_long(n), s['cap'], _p(s['num']), _p(s['flags']), _p(s['pos']),
                          _p(s['dir']), _p(s['len']), _p(np.ascontiguousarray(mom)), F, _p(freqs), _p(cam),
                          _p(out['time']), _p(out['length']), _p(out['lambda']), _p(out['emission']), _p(out['tau']),
                          _p(out['crossings']))
    return out


AUX_NAMES = ['time', 'length', 'lambda', 'emission', 'tau', 'crossings'] + \
            [p + c for p in ('lambda_ave_', 'emission_ave_', 'tau_int_')
             for c in ('rho', 'n_e', 'p_gas', 'Theta_e', 'B', 'sigma', 'beta_inverse')]


def rgb_to_xyz(r, g, b):
    """sRGB (0-255) to CIE XYZ as the input reader does for render_*_rgb keys (utils/colors.cpp:24-36)."""
    lin = [c / 255.0 for c in (r, g, b)]
    lin = [c / 12.92 if c <= 0.040449936 else ((c + 0.055) / 1.055) ** 2.4 for c in lin]
    return (0.4123955889674142 * lin[0] + 0.3575834307637148 * lin[1] + 0.18049264738170154 * lin[2],
            0.21258623078559552 * lin[0] + 0.715170303703411 * lin[1] + 0.0722004986433362 * lin[2],
            0.019297215491746938 * lin[0] + 0.11918386458084851 * lin[1] + 0.9504971251315798 * lin[2])


def render_features(kv):
    """The render_<i>_<f>_* keys of an input file as a Feature array (render_reader.cpp)."""
    quantities = ['rho', 'n_e', 'p_gas', 'Theta_e', 'B', 'sigma', 'beta_inverse']
    types = {'fill': 0, 'thresh': 1, 'rise': 2, 'fall': 3}
    feats = []
    for i in range(1, int(kv.get('render_num_images', 0)) + 1):
        for f in range(1, int(kv['render_%d_num_features' % i]) + 1):
            key = lambda name: kv.get('render_%d_%d_%s' % (i, f, name))
            ft = Feature(image=i - 1, quantity=quantities.index(key('quantity')), type=types[key('type')])
            for name in ('min', 'max', 'tau_scale', 'thresh', 'opacity'):
                setattr(ft, name, float(key(name)) if key(name) is not None else 0.0)
            xyz = [float(v) for v in key('xyz').split(',')] if key('xyz') else rgb_to_xyz(*[float(v) for v in key('rgb').split(',')])
            ft.xyz[:] = xyz
            feats.append(ft)
    return (Feature * len(feats))(*feats), len(feats)


def simulation_image(kv, s, mom, grid, want_inds=True, camera_x=None, render=False, cut_camera_x=None, slow=None):
    """camera_x given: also the 27 auxiliary images, returned as a dict name -> (n) array in place of the indices.
    render: also the false-colour images (render_num_images, 3, n) of the input file's render_* features, as a third value."""
    a = float(kv['simulation_a'])
    P = Sim(a=a, camera_r=float(kv['camera_r']), x_unit=GG_MSUN * float(kv['simulation_m_msun']) / (C * C),
            n_b=grid['n_b'], n_k=grid['n_k'], n_j=grid['n_j'], n_i=grid['n_i'], interp=int(kv['simulation_interp'] == 'true'),
            fallback_nan=int(kv['fallback_nan'] == 'true'), d_unit=float(kv['simulation_rho_cgs']), mu=float(kv['plasma_mu']),
            ne_ni=float(kv['plasma_ne_ni']), rat_low=float(kv['plasma_rat_low']), rat_high=float(kv['plasma_rat_high']),
            cut_sigma_max=float(kv['cut_sigma_max']), coord=int(kv.get('simulation_coord', 'sks') == 'cks'),
            power_frac=float(kv.get('plasma_power_frac', 0.0)), power_p=float(kv.get('plasma_p', 0.0)),
            power_gamma_min=float(kv.get('plasma_gamma_min', 0.0)), power_gamma_max=float(kv.get('plasma_gamma_max', 0.0)),
            kappa_frac=float(kv.get('plasma_kappa_frac', 0.0)), kappa=float(kv.get('plasma_kappa', 0.0)),
            kappa_w=float(kv.get('plasma_w', 0.0)), flat=int(kv.get('ray_flat', 'false') == 'true'),
            cut_omit_in=float(kv.get('cut_omit_in', -1.0)), cut_omit_out=float(kv.get('cut_omit_out', -1.0)),
            use_energy=int(kv.get('plasma_use_p', 'true') == 'false'), gamma=float(kv.get('plasma_gamma', 0.0)),
            gamma_i=float(kv.get('plasma_gamma_i', 0.0)), gamma_e=float(kv.get('plasma_gamma_e', 0.0)),
            code_kappa=int(kv.get('plasma_model', 'ti_te_beta') == 'code_kappa'),
            cut_omit_near=int(kv.get('cut_omit_near', 'false') == 'true'), cut_omit_far=int(kv.get('cut_omit_far', 'false') == 'true'),
            cut_plane=int(kv.get('cut_plane', 'false') == 'true'),
            cut_midplane_theta=float(kv.get('cut_midplane_theta', 0.0)) * np.pi / 180.0,
            cut_midplane_z=float(kv.get('cut_midplane_z', 0.0)), fallback_rho=float(kv.get('fallback_rho', 0.0)),
            fallback_pgas=float(kv.get('fallback_pgas', 0.0)), fallback_kappa=float(kv.get('fallback_kappa', 0.0)))
    triple = lambda key: (ctypes.c_double * 3)(*[float(v) for v in kv.get(key, '0,0,0').split(',')])
    P.cut_plane_origin, P.cut_plane_normal = triple('cut_plane_origin'), triple('cut_plane_normal')
    # the sigma maximum is the struct's own cut_sigma_max
    P.cut_val_min = (ctypes.c_double * 7)(*[float(kv.get('cut_%s_min' % v, -1.0)) for v in CUT_VALUES])
    P.cut_val_max = (ctypes.c_double * 7)(*[-1.0 if v == 'sigma' else float(kv.get('cut_%s_max' % v, -1.0)) for v in CUT_VALUES])
    if P.cut_omit_near or P.cut_omit_far:
        P.cut_cam = (ctypes.c_double * 3)(*[float(v) for v in cut_camera_x[1:4]])
    if slow is not None:   # dict(times=descending slice times, snapshot_time=...); grid['prim'] is (n_t, n_var, ...)
        P.n_t, P.slow_interp, P.snapshot_time = len(slow['times']), int(kv['slow_interp'] == 'true'), float(slow['snapshot_time'])
        for q, t in enumerate(slow['times']):
            P.times[q] = float(t)
    n = len(mom)
    image = np.zeros(n)
    inds = np.full((n, s['cap'], 4), -1, np.int32) if want_inds else None
    aux = np.zeros((27, n)) if camera_x is not None else None
    feats, n_feat = render_features(kv) if render else (None, 0)
    n_render = int(kv.get('render_num_images', 0)) if render else 0
    rendering = np.zeros((n_render, 3, n)) if render else None
    keep = [np.ascontiguousarray(grid[k]) for k in ('x1f', 'x2f', 'x3f', 'x1v', 'x2v', 'x3v', 'prim')]
    if kv.get('simulation_block_interp', 'false') == 'true' and kv['simulation_interp'] == 'true':
        # the reference reads one element past a block's cell centres at its upper edge: pad the last block
        for q in (3, 4, 5):
            keep[q] = np.ascontiguousarray(np.concatenate([keep[q].ravel(), [0.0]]))
        mesh = [np.ascontiguousarray(grid['levels'], np.int32), np.ascontiguousarray(grid['locations'], np.int32)]
        P.block_interp, P.n_3_root = 1, int(grid['n_3_root'])
        P.levels, P.locations = mesh[0].ctypes.data, mesh[1].ctypes.data
    lib().orc_simulation_image(ctypes.byref(P), ctypes.c_long(n), s['cap'], _p(s['num']), _p(s['flags']), _p(s['pos']),
                               _p(s['dir']), _p(s['len']), _p(np.ascontiguousarray(mom)), ctypes.c_double(float(kv['image_frequency'])),
                               *[_p(k) for k in keep], _p(image), _p(inds),
                               _p(None if camera_x is None else np.ascontiguousarray(camera_x, np.float64)), _p(aux),
                               n_feat, feats, n_render, _p(rendering))
    if render:
        return image, (dict(zip(AUX_NAMES, aux)) if aux is not None else inds), rendering
    if aux is not None:
        return image, dict(zip(AUX_NAMES, aux))
    return image, inds


def slow_window(file_times, chunk, snapshot_time, first=0):
    """Files resident for the first image of a slow-light run (simulation_reader.cpp:211-262): reading starts with
    file first + chunk - 1 and advances until a file's time reaches snapshot_time (or the series ends); the window is
    that file and the chunk - 1 before it, newest first.  Returns the file numbers."""
    latest = first + chunk - 2
    latest_time = -np.inf
    while latest_time < snapshot_time and latest < first + len(file_times) - 1:
        latest += 1
        latest_time = file_times[latest - first]
    return [latest - q for q in range(chunk)]
