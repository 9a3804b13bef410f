"""Readers for the reference's binary dumps (TEST INFRASTRUCTURE).

Format: reference src/utils/file_io.cpp:64-75 -- an Array<> is five int32 extents n1..n5 (fastest
first) followed by raw little-endian data; scalars and fixed-size vectors are raw.
"""
import numpy as np


class _Reader:
    def __init__(self, path):
        self.f = open(path, 'rb')

    def raw(self, dtype, count):
        return np.fromfile(self.f, dtype=dtype, count=count)

    def array(self, dtype):
        n = self.raw(np.int32, 5)
        shape = tuple(int(v) for v in n[::-1] if True)
        data = self.raw(dtype, int(np.prod(n, dtype=np.int64)))
        shape = tuple(int(v) for v in n[::-1])
        # drop leading singleton extents (n5..): keep the trailing significant ones
        while len(shape) > 1 and shape[0] == 1:
            shape = shape[1:]
        return data.reshape(shape)


def read_geodesic_checkpoint(path):
    """Level-0 geodesic checkpoint (reference geodesic_checkpoint.cpp:28-59)."""
    r = _Reader(path)
    out = {}
    for name in ('cam_x', 'u_con', 'u_cov', 'norm_con', 'norm_con_c', 'hor_con_c', 'vert_con_c'):
        out[name] = r.raw(np.float64, 4)
    out['camera_pos'] = r.array(np.float64)
    out['camera_dir'] = r.array(np.float64)
    out['image_frequencies'] = r.array(np.float64)
    out['momentum_factors'] = r.array(np.float64)
    out['geodesic_num_steps'] = int(r.raw(np.int32, 1)[0])
    out['sample_flags'] = r.array(np.uint8)
    out['sample_num'] = r.array(np.int32)
    out['sample_pos'] = r.array(np.float64)
    out['sample_dir'] = r.array(np.float64)
    out['sample_len'] = r.array(np.float64)
    return out


def read_sample_checkpoint(path, interp=True):
    """Sampling checkpoint (reference sample_checkpoint.cpp:22-39)."""
    r = _Reader(path)
    out = {'sample_inds': r.array(np.int32)}
    if interp:
        out['sample_fracs'] = r.array(np.float64)
    out['sample_nan'] = r.array(np.uint8)
    out['sample_fallback'] = r.array(np.uint8)
    return out
