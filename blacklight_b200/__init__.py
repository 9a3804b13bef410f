"""blacklight_b200 -- Python face of the B200-native ray-tracing hot path.

Thin ctypes layer over the C ABI in include/blacklight_b200.h / blacklight_b200_host.h.  The product
is the shared library built from blacklight_b200/csrc (CUDA sm_100a kernels + C++ host layer); this
module exists so tests and bench.py can drive it with numpy/torch host buffers.  There is no Python or
CPU implementation of the path: importing works anywhere, but any compute call raises unless the
library is built and a CUDA device is present.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BLACKLIGHT_B200_LIB points at an alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get('BLACKLIGHT_B200_LIB') or os.path.join(_HERE, 'libblacklight_b200.so')
EXE_PATH = os.path.join(_HERE, 'bin', 'blacklight_b200')


class BlacklightError(RuntimeError):
    pass


class GridView(ctypes.Structure):
    """bl_grid_view"""
    _fields_ = [('n_b', ctypes.c_int32), ('n_k', ctypes.c_int32), ('n_j', ctypes.c_int32), ('n_i', ctypes.c_int32),
                ('n_var', ctypes.c_int32), ('levels', ctypes.c_void_p), ('locations', ctypes.c_void_p),
                ('x1f', ctypes.c_void_p), ('x2f', ctypes.c_void_p), ('x3f', ctypes.c_void_p),
                ('x1v', ctypes.c_void_p), ('x2v', ctypes.c_void_p), ('x3v', ctypes.c_void_p),
                ('prim', ctypes.c_void_p)] + [(n, ctypes.c_int32) for n in (
                    'ind_rho', 'ind_pgas', 'ind_kappa', 'ind_uu1', 'ind_uu2', 'ind_uu3', 'ind_bb1', 'ind_bb2', 'ind_bb3',
                    'n_3_root')] + [('sks_map', ctypes.c_void_p), ('sks_map_n1', ctypes.c_int32), ('sks_map_n2', ctypes.c_int32),
                                    ('sks_map_r_in', ctypes.c_double), ('sks_map_dr', ctypes.c_double),
                                    ('sks_map_dtheta', ctypes.c_double), ('simulation_bounds', ctypes.c_double * 6)]


class LevelStats(ctypes.Structure):
    """bl_level_stats"""
    _fields_ = [('num_rays', ctypes.c_int64), ('geodesic_num_steps', ctypes.c_int32),
                ('num_bad_geodesics', ctypes.c_int64), ('num_samples', ctypes.c_int64),
                ('num_attempts', ctypes.c_int64), ('num_accepted', ctypes.c_int64),
                ('ms_geodesic', ctypes.c_double), ('ms_radiation', ctypes.c_double), ('ms_refine', ctypes.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Camera(ctypes.Structure):
    """bl_camera"""
    _fields_ = [('type', ctypes.c_int32), ('normalization', ctypes.c_int32), ('width', ctypes.c_double), ('r', ctypes.c_double)] + \
               [(n, ctypes.c_double * 4) for n in ('x', 'u_con', 'u_cov', 'norm_con', 'norm_con_c', 'hor_con_c', 'vert_con_c')]


PIXELS_ROWS, PIXELS_BLOCKS = 0, 1   # bl_trace_level_pixels unit kinds

_lib = None


def load_library():
    """Load libblacklight_b200.so; raises BlacklightError (never falls back) if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BlacklightError('%s is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                              '(or make -C blacklight_b200/csrc); there is no CPU fallback' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    sig = {
        'bl_create': (i32, [vp, ctypes.POINTER(vp)]), 'bl_destroy': (None, [vp]),
        'bl_last_error': (ctypes.c_char_p, [vp]), 'bl_image_num_quantities': (i32, [vp]),
        'bl_upload_grid': (i32, [vp, ctypes.POINTER(GridView)]),
        'bl_trace_level': (i32, [vp, i32, vp, vp, vp, i64, ctypes.POINTER(LevelStats)]),
        'bl_radiate_level': (i32, [vp, i32, i32, vp, vp, ctypes.POINTER(LevelStats)]),
        'bl_refine_level': (i32, [vp, i32, vp, i64, vp, ctypes.POINTER(i64)]),
        'bl_set_taps': (i32, [vp, i32]),
        'bl_retrace_level': (i32, [vp, i32, ctypes.POINTER(LevelStats)]),
        'bl_upload_samples': (i32, [vp, i32, vp, vp, vp, i64, i32, vp, vp, vp, vp, vp, ctypes.POINTER(LevelStats)]),
        'bl_launch_count': (ctypes.c_longlong, [vp]), 'bl_cuda_stream': (vp, [vp]),
        'bl_device_image': (i32, [vp, i32, ctypes.POINTER(vp), ctypes.POINTER(i64)]),
        'bl_download_polarized_scratch': (i32, [vp, i32, vp, vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64)]),
        'bl_polarized_stage_ms': (i32, [vp, i32, ctypes.POINTER(dbl * 3), ctypes.POINTER(ctypes.c_int32)]),
        'bl_polarized_sampling_ms': (i32, [vp, i32, ctypes.POINTER(dbl)]),
        'bl_download_samples': (i32, [vp, i32, vp, vp, vp, vp, vp]),
        'bl_download_sample_inds': (i32, [vp, i32, vp, vp, vp, vp, vp]),
        'bl_device_info': (i32, [vp, ctypes.c_char_p, i32, ctypes.POINTER(i32), ctypes.POINTER(dbl)]),
        'bl_measure_fp64_peak': (i32, [vp, ctypes.POINTER(dbl)]),
        'bl_selftest_division': (i32, [vp, ctypes.c_uint64, i64, ctypes.POINTER(i64)]),
        'blh_last_error': (ctypes.c_char_p, []), 'blh_config_from_input': (i32, [ctypes.c_char_p, ctypes.POINTER(vp)]),
        'blh_config_free': (None, [vp]), 'blh_config_params': (vp, [vp]), 'blh_config_num_runs': (i32, [vp]),
        'blh_config_set_device': (None, [vp, i32, i64]), 'blh_config_set_level0_block_major': (None, [vp, i32]), 'blh_camera_frame': (i32, [vp, vp]),
        'blh_camera_root': (i64, [vp, vp, vp, vp]),
        'blh_camera_refined': (i64, [vp, i32, vp, vp, i64, vp, vp, vp, vp]),
        'blh_run_input_file': (i32, [ctypes.c_char_p, i32, i32, vp]),
        'blh_crc32': (ctypes.c_uint32, [vp, ctypes.c_uint64]),
        'blh_camera_rows': (i64, [vp, vp, i64, vp, vp, vp]),
        'blh_run_input_file_devices': (i32, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), i32, i32, vp]),
        'bl_device_count': (i32, []),
        'bl_set_camera': (i32, [vp, ctypes.POINTER(Camera)]),
        'bl_trace_level_pixels': (i32, [vp, i32, i32, vp, i64, ctypes.POINTER(LevelStats)]),
        'bl_download_camera': (i32, [vp, i32, vp, vp, vp]),
        'blh_camera_struct': (i32, [vp, ctypes.POINTER(Camera)]),
        'blh_camera_blocks': (i64, [vp, i32, vp, i64, vp, vp, vp]),
        'blh_snapshot_read': (i32, [vp, ctypes.c_char_p, ctypes.POINTER(vp)]),
        'blh_snapshot_view': (i32, [vp, ctypes.POINTER(GridView), ctypes.POINTER(dbl), ctypes.POINTER(dbl)]),
        'blh_snapshot_free': (None, [vp]), 'blh_snapshot_reread': (i32, [vp, ctypes.c_char_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def parse_input_text(text):
    """key -> value of a parameter file (same rules as the reference: whitespace removed, # comments)."""
    out = {}
    for line in text.splitlines():
        line = ''.join(line.split()).split('#')[0]
        if line:
            k, v = line.split('=', 1)
            out[k] = v
    return out


class Config:
    """Parsed parameter file + camera frame (host side; blh_config)."""

    def __init__(self, input_path, device=0, tile_rays=0):
        lib = load_library()
        self._h = ctypes.c_void_p()
        if lib.blh_config_from_input(os.fsencode(input_path), ctypes.byref(self._h)) != 0:
            raise BlacklightError(lib.blh_last_error().decode())
        lib.blh_config_set_device(self._h, device, tile_rays)
        with open(input_path) as f:
            self.keys = parse_input_text(f.read())

    def __del__(self):
        if getattr(self, '_h', None) and _lib is not None:
            _lib.blh_config_free(self._h)
            self._h = None

    def set_level0_block_major(self, on=True):
        """Level-0 rays will be handed over block by block (sharded adaptive runs); set before Context()."""
        _lib.blh_config_set_level0_block_major(self._h, 1 if on else 0)
        self.level0_block_major = bool(on)

    @property
    def params_ptr(self):
        return _lib.blh_config_params(self._h)

    @property
    def resolution(self):
        return int(self.keys['camera_resolution'])

    @property
    def block_size(self):
        return int(self.keys.get('adaptive_block_size', 0))

    def camera_frame(self):
        out = np.empty((7, 4))
        _lib.blh_camera_frame(self._h, _ptr(out))
        return dict(zip(('cam_x', 'u_con', 'u_cov', 'norm_con', 'norm_con_c', 'hor_con_c', 'vert_con_c'), out))

    def camera_root(self, pinned=False):
        n = self.resolution ** 2
        pos, dirs, fac = _host_array((n, 4), pinned), _host_array((n, 4), pinned), _host_array((n,), pinned)
        if _lib.blh_camera_root(self._h, _ptr(pos), _ptr(dirs), _ptr(fac)) != n:
            raise BlacklightError(_lib.blh_last_error().decode())
        return pos, dirs, fac

    def camera_refined(self, level, parent_locs, flags):
        parent_locs = np.ascontiguousarray(parent_locs, np.int32)
        flags = np.ascontiguousarray(flags, np.uint8)
        nb = 4 * int(np.count_nonzero(flags))
        npix = nb * self.block_size ** 2
        locs, pos, dirs, fac = np.empty((nb, 2), np.int32), np.empty((npix, 4)), np.empty((npix, 4)), np.empty(npix)
        got = _lib.blh_camera_refined(self._h, level, _ptr(parent_locs), _ptr(flags), len(flags), _ptr(locs),
                                      _ptr(pos), _ptr(dirs), _ptr(fac))
        if got != nb:
            raise BlacklightError(_lib.blh_last_error().decode())
        return locs, pos, dirs, fac


    def camera_rows(self, rows, pinned=False):
        """Camera arrays of the given level-0 image rows only: blh_camera_rows."""
        rows = np.ascontiguousarray(rows, np.int64)
        n = len(rows) * self.resolution
        pos, dirs, fac = _host_array((n, 4), pinned), _host_array((n, 4), pinned), _host_array((n,), pinned)
        if _lib.blh_camera_rows(self._h, _ptr(rows), len(rows), _ptr(pos), _ptr(dirs), _ptr(fac)) != n:
            raise BlacklightError(_lib.blh_last_error().decode())
        return pos, dirs, fac

    def camera_blocks(self, level, locs):
        """Camera arrays of the given blocks of a level only ((n,2) block locations): blh_camera_blocks."""
        locs = np.ascontiguousarray(locs, np.int32).reshape(-1, 2)
        npix = len(locs) * self.block_size ** 2
        pos, dirs, fac = np.empty((npix, 4)), np.empty((npix, 4)), np.empty(npix)
        if _lib.blh_camera_blocks(self._h, level, _ptr(locs), len(locs), _ptr(pos), _ptr(dirs), _ptr(fac)) != len(locs):
            raise BlacklightError(_lib.blh_last_error().decode())
        return pos, dirs, fac


def _host_array(shape, pinned=False, dtype=np.float64):
    if pinned:
        import torch
        t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
        a = t.numpy()
        a_base = a  # keep tensor alive through the array's base chain
        a_base.flags.writeable = True
        _PINNED_KEEPALIVE.append(t)
        return a
    return np.empty(shape, dtype)


_PINNED_KEEPALIVE = []


class Context:
    """One GPU context (bl_ctx): grid residency, geodesic step buffers, radiation kernels."""

    def __init__(self, config):
        lib = load_library()
        self.config = config
        self.level0_block_major = bool(getattr(config, 'level0_block_major', False))   # as bl_create saw it
        self._h = ctypes.c_void_p()
        if lib.bl_create(config.params_ptr, ctypes.byref(self._h)) != 0:
            raise BlacklightError(lib.bl_last_error(None).decode())
        self.num_quantities = lib.bl_image_num_quantities(self._h)
        self._rays = {}
        self._steps = {}

    def close(self):
        if getattr(self, '_h', None) and _lib is not None:
            _lib.bl_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != 0:
            raise BlacklightError(_lib.bl_last_error(self._h).decode())

    def device_info(self):
        name = ctypes.create_string_buffer(256)
        sm, free = ctypes.c_int(), ctypes.c_double()
        self._check(_lib.bl_device_info(self._h, name, 256, ctypes.byref(sm), ctypes.byref(free)))
        return {'name': name.value.decode(), 'sm_count': sm.value, 'hbm_free_gb': free.value}

    def measure_fp64_peak(self):
        out = ctypes.c_double()
        self._check(_lib.bl_measure_fp64_peak(self._h, ctypes.byref(out)))
        return out.value

    def selftest_division(self, num_pairs, seed=1):
        """Mismatches between the shared-reciprocal division and the hardware IEEE division (must be 0)."""
        out = ctypes.c_int64()
        self._check(_lib.bl_selftest_division(self._h, seed, num_pairs, ctypes.byref(out)))
        return out.value

    def upload_grid(self, g):
        """g: dict from blacklight_b200.mock_snapshot.grid_view_arrays (or any reader) -- host numpy arrays."""
        keep = {k: np.ascontiguousarray(g[k]) for k in ('levels', 'locations', 'x1f', 'x2f', 'x3f', 'x1v', 'x2v', 'x3v', 'prim')}
        assert keep['prim'].dtype == np.float32 and keep['x1f'].dtype == np.float64
        v = GridView()
        for k in ('n_b', 'n_k', 'n_j', 'n_i', 'n_var', 'ind_rho', 'ind_pgas', 'ind_kappa', 'ind_uu1', 'ind_uu2', 'ind_uu3',
                  'ind_bb1', 'ind_bb2', 'ind_bb3', 'n_3_root'):
            setattr(v, k, int(g[k]))
        for k, a in keep.items():
            setattr(v, k, a.ctypes.data)
        if 'sks_map' in g:   # simulation_coord = fmks (blacklight_b200.read_snapshot of an iharm3d dump)
            keep['sks_map'] = np.ascontiguousarray(g['sks_map'], np.float64)
            v.sks_map = keep['sks_map'].ctypes.data
            v.sks_map_n2, v.sks_map_n1 = keep['sks_map'].shape[1:]
            v.sks_map_r_in, v.sks_map_dr, v.sks_map_dtheta = g['sks_map_r_in'], g['sks_map_dr'], g['sks_map_dtheta']
            for d in range(6):
                v.simulation_bounds[d] = float(g['simulation_bounds'][d])
        self._check(_lib.bl_upload_grid(self._h, ctypes.byref(v)))

    def set_camera(self, config=None):
        """Hand the camera frame of `config` (default: the context's own) to the device: bl_set_camera."""
        cam = Camera()
        cfg = config or self.config
        if _lib.blh_camera_struct(cfg._h, ctypes.byref(cam)) != 0:
            raise BlacklightError(_lib.blh_last_error().decode())
        self._check(_lib.bl_set_camera(self._h, ctypes.byref(cam)))
        self._have_camera = True

    def trace_level_pixels(self, level, rows=None, blocks=None):
        """Trace a level whose camera pixels are generated on the device (bl_trace_level_pixels): `rows` = image rows of
        the level's raster (None with blocks None: the whole raster), or `blocks` = (B, 2) int32 (v, u) block locations."""
        if not getattr(self, '_have_camera', False):
            self.set_camera()
        res = self.config.resolution << level
        st = LevelStats()
        if blocks is not None:
            units = np.ascontiguousarray(blocks, np.int32).reshape(-1, 2)
            n_units, per, kind = len(units), self.config.block_size ** 2, PIXELS_BLOCKS
        elif rows is not None:
            units = np.ascontiguousarray(rows, np.int32).ravel()
            n_units, per, kind = len(units), res, PIXELS_ROWS
        else:
            units, n_units, per, kind = None, res, res, PIXELS_ROWS
        self._check(_lib.bl_trace_level_pixels(self._h, level, kind, _ptr(units), n_units, ctypes.byref(st)))
        self._rays[level] = n_units * per
        self._steps[level] = st.geodesic_num_steps
        return st.as_dict()

    def download_camera(self, level=0):
        """(pos (N,4), dir (N,4), factor (N)) of the level's rays as they sit in HBM: bl_download_camera."""
        n = self._rays[level]
        pos, dirs, fac = np.empty((n, 4)), np.empty((n, 4)), np.empty(n)
        self._check(_lib.bl_download_camera(self._h, level, _ptr(pos), _ptr(dirs), _ptr(fac)))
        return pos, dirs, fac

    def trace_level(self, level, pos, dirs, fac):
        pos, dirs, fac = (np.ascontiguousarray(a, np.float64) for a in (pos, dirs, fac))
        st = LevelStats()
        self._check(_lib.bl_trace_level(self._h, level, _ptr(pos), _ptr(dirs), _ptr(fac), len(fac), ctypes.byref(st)))
        self._rays[level] = len(fac)
        self._steps[level] = st.geodesic_num_steps
        return st.as_dict()

    def upload_samples(self, level, pos, dirs, fac, flags, num, sample_pos, sample_dir, sample_len):
        """Load geodesics integrated elsewhere (reference checkpoint layouts) instead of tracing the level."""
        pos, dirs, fac = (np.ascontiguousarray(a, np.float64) for a in (pos, dirs, fac))
        flags = np.ascontiguousarray(flags, np.uint8)
        num = np.ascontiguousarray(num, np.int32)
        sp, sd, sl = (np.ascontiguousarray(a, np.float64) for a in (sample_pos, sample_dir, sample_len))
        st = LevelStats()
        self._check(_lib.bl_upload_samples(self._h, level, _ptr(pos), _ptr(dirs), _ptr(fac), len(fac), sl.shape[1], _ptr(flags),
                                           _ptr(num), _ptr(sp), _ptr(sd), _ptr(sl), ctypes.byref(st)))
        self._rays[level] = len(fac)
        self._steps[level] = st.geodesic_num_steps
        return st.as_dict()

    def retrace_level(self, level):
        st = LevelStats()
        self._check(_lib.bl_retrace_level(self._h, level, ctypes.byref(st)))
        return st.as_dict()

    def launch_count(self):
        return int(_lib.bl_launch_count(self._h))

    def device_image(self, level=0):
        """(device pointer, (Q, rays)) of the level's image in HBM: bl_device_image."""
        ptr, n = ctypes.c_void_p(), ctypes.c_int64()
        self._check(_lib.bl_device_image(self._h, level, ctypes.byref(ptr), ctypes.byref(n)))
        return ptr.value, (self.num_quantities, n.value)

    def polarized_scratch(self, level=0):
        """(fields, slab, rays) scratch of the last slab of the polarized pipeline and the (10, rays) camera
        half-step map: bl_download_polarized_scratch."""
        nf, slab, rays = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        self._check(_lib.bl_download_polarized_scratch(self._h, level, None, None, ctypes.byref(nf), ctypes.byref(slab), ctypes.byref(rays)))
        out, cam = np.empty((nf.value, slab.value, rays.value)), np.empty((10, rays.value))
        self._check(_lib.bl_download_polarized_scratch(self._h, level, _ptr(out), _ptr(cam), None, None, None))
        return out, cam

    def polarized_stage_ms(self, level=0):
        """Device ms of the four polarized stages (sampling, geometry, coefficients, transfer) in the last radiate_level
        and the slab length; slab 0 means the fused kernel ran."""
        ms, slab = (ctypes.c_double * 3)(), ctypes.c_int32()
        self._check(_lib.bl_polarized_stage_ms(self._h, level, ctypes.byref(ms), ctypes.byref(slab)))
        smp = ctypes.c_double()
        self._check(_lib.bl_polarized_sampling_ms(self._h, level, ctypes.byref(smp)))
        return {'sampling_ms': smp.value, 'geometry_ms': ms[0], 'coefficients_ms': ms[1], 'transfer_ms': ms[2], 'slab': slab.value}

    def cuda_stream(self):
        """cudaStream_t (as an integer) that this context's kernels and copies are issued on."""
        return int(_lib.bl_cuda_stream(self._h) or 0)

    def radiate_level(self, level, snapshot=0, image=None, render=None, num_render=0, download=True):
        n = self._rays[level]
        if image is None and download:
            image = np.empty((self.num_quantities, n))
        if render is None and num_render > 0:
            render = np.empty((num_render, 3, n))
        st = LevelStats()
        self._check(_lib.bl_radiate_level(self._h, level, snapshot, _ptr(image), _ptr(render), ctypes.byref(st)))
        self._steps[level] = st.geodesic_num_steps
        return image, render, st.as_dict()

    def refine_level(self, level, block_locs):
        block_locs = np.ascontiguousarray(block_locs, np.int32)
        flags = np.zeros(len(block_locs), np.uint8)
        cnt = ctypes.c_int64()
        self._check(_lib.bl_refine_level(self._h, level, _ptr(block_locs), len(block_locs), _ptr(flags), ctypes.byref(cnt)))
        return flags, cnt.value

    def set_taps(self, enabled=True):
        self._check(_lib.bl_set_taps(self._h, 1 if enabled else 0))

    def download_samples(self, level, arrays=True):
        n, s = self._rays[level], max(self._steps[level], 1)
        flags, num = np.empty(n, np.uint8), np.empty(n, np.int32)
        pos = np.empty((n, s, 4)) if arrays else None
        dirs = np.empty((n, s, 4)) if arrays else None
        length = np.empty((n, s)) if arrays else None
        self._check(_lib.bl_download_samples(self._h, level, _ptr(flags), _ptr(num), _ptr(pos), _ptr(dirs), _ptr(length)))
        return dict(flags=flags, num=num, pos=pos, dir=dirs, len=length)

    def download_sample_inds(self, level, interp=True):
        n, s = self._rays[level], max(self._steps[level], 1)
        inds = np.empty((n, s, 4), np.int32)
        fracs = np.empty((n, s, 3)) if interp else None
        nan_, cut, fb = (np.empty((n, s), np.uint8) for _ in range(3))
        self._check(_lib.bl_download_sample_inds(self._h, level, _ptr(inds), _ptr(fracs), _ptr(nan_), _ptr(cut), _ptr(fb)))
        return dict(inds=inds, fracs=fracs, nan=nan_, cut=cut, fallback=fb)


def read_snapshot(config, path=None, then=None):
    """Read one snapshot with the reader the input file selects (simulation_format = athena, athenak, iharm3d or
    harm3d): blh_snapshot_read.  Returns the arrays of bl_grid_view as numpy copies (what Context.upload_grid takes)
    plus 'time' and 'plasma_gamma'.  then: a later file of the same series, read on top of the first one the way the
    driver does for time series (blh_snapshot_reread: layout kept, cell data and time refreshed)."""
    lib = load_library()
    h = ctypes.c_void_p()
    if lib.blh_snapshot_read(config._h, os.fsencode(path) if path else None, ctypes.byref(h)) != 0:
        raise BlacklightError(lib.blh_last_error().decode())
    try:
        if then is not None and lib.blh_snapshot_reread(h, os.fsencode(then)) != 0:
            raise BlacklightError(lib.blh_last_error().decode())
        v, t, g = GridView(), ctypes.c_double(), ctypes.c_double()
        lib.blh_snapshot_view(h, ctypes.byref(v), ctypes.byref(t), ctypes.byref(g))
        out = {n: getattr(v, n) for n, c in GridView._fields_ if c is ctypes.c_int32 and not n.startswith('sks_map')}

        def arr(ptr, shape, dtype):
            n = int(np.prod(shape))
            buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()
        nb, nk, nj, ni = v.n_b, v.n_k, v.n_j, v.n_i
        out['levels'] = arr(v.levels, (nb,), np.int32)
        out['locations'] = arr(v.locations, (nb, 3), np.int32)
        for name, n in (('x1', ni), ('x2', nj), ('x3', nk)):
            out[name + 'f'] = arr(getattr(v, name + 'f'), (nb, n + 1), np.float64)
            out[name + 'v'] = arr(getattr(v, name + 'v'), (nb, n), np.float64)
        out['prim'] = arr(v.prim, (v.n_var, nb, nk, nj, ni), np.float32)
        out['time'], out['plasma_gamma'] = t.value, g.value
        if v.sks_map:   # simulation_coord = fmks
            out['sks_map'] = arr(v.sks_map, (2, v.sks_map_n2, v.sks_map_n1), np.float64)
            out['simulation_bounds'] = np.array(list(v.simulation_bounds))
            for k in ('sks_map_r_in', 'sks_map_dr', 'sks_map_dtheta'):
                out[k] = getattr(v, k)
        return out
    finally:
        lib.blh_snapshot_free(h)


def run_input_file(path, device=-1, quiet=True, devices=None):
    """Full drop-in run (read input, trace, radiate, write output): blh_run_input_file, or on a list of CUDA devices
    blh_run_input_file_devices (one context per device inside this process)."""
    lib = load_library()
    t = np.zeros(12)
    if devices is not None:
        arr = (ctypes.c_int * len(devices))(*devices)
        rc = lib.blh_run_input_file_devices(os.fsencode(path), arr, len(devices), 1 if quiet else 0, _ptr(t))
    else:
        rc = lib.blh_run_input_file(os.fsencode(path), device, 1 if quiet else 0, _ptr(t))
    if rc != 0:
        raise BlacklightError(lib.blh_last_error().decode())
    names = ('total_s', 'geodesic_s', 'read_s', 'sample_s', 'image_s', 'render_s', 'gpu_geodesic_ms',
             'gpu_radiation_ms', 'gpu_refine_ms', 'rays', 'samples', 'devices')
    return dict(zip(names, t))
