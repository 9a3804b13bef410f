"""Synthetic GRMHD snapshots: the input generator of bench.py and of the parity tests (not on the product path).

Restates the closed-form disk model of the reference's benchmark-input generator
(reference scripts/generate_mock_simulation.py:24-77, defaults :346-425) -- a power-law torus in
spherical Kerr-Schild coordinates with an m=4 perturbation, Keplerian-like u^phi, vertical + toroidal
field -- and writes it as an Athena++ `.athdf` file with a numpy-only HDF5 writer (no h5py in this
image).  The file layout is the subset the reference's own parser accepts (SURVEY.md Appendix C:
superblock v0, symbol-table root group, v1 object headers, contiguous float32/int datasets).

The same arrays are returned in the host layout the C ABI's bl_grid_view expects, so the CUDA path and
the reference binary see identical inputs.  `blocks=(nb_r, nb_th, nb_ph)` re-partitions the same cells
into several MeshBlocks (all level 0) to exercise the block search.
"""
import struct

import numpy as np


def mock_fields(n_r=77, n_th=64, n_ph=128, pert_amp=0.1, pert_n_r=3.0, pert_n_th=2.0, pert_n_ph=4,
                r_min=None, r_max=None):
    """Face/centre coordinates and primitives of the mock torus (float64, (ph, th, r) order)."""
    if r_min is None:
        r_min = 2.0 * 25.0 ** (-1.0 / 75.0)
    if r_max is None:
        r_max = 2.0 * 25.0 ** (76.0 / 75.0)
    rho_amp, rho_r_power, rho_th_scale, rho_floor = 1.0, 0.5, np.pi / 8.0, 1.0e-8
    pgas_amp, pgas_r_power, pgas_th_scale, pgas_floor = 0.1, 1.25, np.pi / 8.0, 1.0e-9
    r_isco = 6.0
    omega_isco = r_isco ** -1.5
    gamma_isco = (1.0 - 2.0 / r_isco - r_isco ** 2 * omega_isco ** 2) ** -0.5
    uph_r_power = 1.5
    uph_amp = gamma_isco * omega_isco * r_isco ** uph_r_power
    uph_th_scale = np.pi / 8.0
    bph_amp, bph_r_power, bph_th_scale = 0.2, 1.75, np.pi / 8.0
    bz_amp, bz_rr_power = 0.02, 0.625
    cut_r_min, cut_r_max, cut_th_min = 2.0, 50.0, np.pi / 16.0

    rf = np.exp(np.linspace(np.log(r_min), np.log(r_max), n_r + 1))
    thf = np.linspace(0.0, np.pi, n_th + 1)
    phf = np.linspace(0.0, 2.0 * np.pi, n_ph + 1)
    r = 0.5 * (rf[:-1] + rf[1:])
    th = 0.5 * (thf[:-1] + thf[1:])
    ph = 0.5 * (phf[:-1] + phf[1:])
    R, TH, PH = r[None, None, :], th[None, :, None], ph[:, None, None]

    inside = (np.where((r < cut_r_min) | (r > cut_r_max), 0.0, 1.0)[None, None, :]
              * np.where((th < cut_th_min) | (th > np.pi - cut_th_min), 0.0, 1.0)[None, :, None]
              * np.ones_like(ph)[:, None, None])
    wave_r = np.cos(2.0 * np.pi * pert_n_r * np.log(r / cut_r_min) / np.log(cut_r_max / cut_r_min))
    wave_th = -np.cos(2.0 * np.pi * pert_n_th * (th - cut_th_min) / (np.pi - 2.0 * cut_th_min))
    wave_ph = np.cos(pert_n_ph * ph)
    pert = 1.0 + pert_amp * wave_r[None, None, :] * wave_th[None, :, None] * wave_ph[:, None, None]
    off = np.abs(TH - np.pi / 2.0)

    rho = np.maximum(rho_amp * R ** -rho_r_power * np.exp(-off / rho_th_scale) * pert * inside, rho_floor)
    pgas = np.maximum(pgas_amp * R ** -pgas_r_power * np.exp(-off / pgas_th_scale) * pert ** 2 * inside, pgas_floor)
    uur = np.zeros_like(rho)
    uuth = np.zeros_like(rho)
    uuph = uph_amp * R ** -uph_r_power * np.exp(-off / uph_th_scale) * inside
    cyl = np.maximum(R * np.sin(TH), cut_r_min)
    bbz = bz_amp * cyl ** -bz_rr_power
    bbr = np.cos(TH) * bbz * np.ones_like(PH)
    bbth = -np.sin(TH) / R * bbz * np.ones_like(PH)
    bbph = bph_amp * R ** -bph_r_power * np.exp(-off / bph_th_scale) * np.ones_like(PH)
    bbph = bbph * np.where(th > np.pi / 2.0, -1.0, 1.0)[None, :, None]
    prim = np.stack([rho, pgas, uur, uuth, uuph, bbr, bbth, bbph]).astype(np.float32)  # (8, ph, th, r)
    return dict(rf=rf, thf=thf, phf=phf, r=r, th=th, ph=ph, prim=prim)


def mock_fields_at(R, TH, PH, pert_amp=0.1, pert_n_r=3.0, pert_n_th=2.0, pert_n_ph=4, inflow=0.05):
    """The closed-form torus of mock_fields evaluated at arbitrary (broadcastable) spherical Kerr-Schild points,
    plus a small poloidal velocity (`inflow`) so that every component of the vector transformations of the Harm
    readers carries signal.  Returns 8 float64 arrays (rho, pgas, normal-frame uu^r, uu^th, uu^ph, B^r, B^th, B^ph)."""
    r_isco = 6.0
    omega_isco = r_isco ** -1.5
    gamma_isco = (1.0 - 2.0 / r_isco - r_isco ** 2 * omega_isco ** 2) ** -0.5
    uph_amp = gamma_isco * omega_isco * r_isco ** 1.5
    cut_r_min, cut_r_max, cut_th_min, scale = 2.0, 50.0, np.pi / 16.0, np.pi / 8.0
    R, TH, PH = np.broadcast_arrays(R, TH, PH)
    inside = np.where((R < cut_r_min) | (R > cut_r_max) | (TH < cut_th_min) | (TH > np.pi - cut_th_min), 0.0, 1.0)
    pert = 1.0 + pert_amp * (np.cos(2.0 * np.pi * pert_n_r * np.log(R / cut_r_min) / np.log(cut_r_max / cut_r_min))
                             * -np.cos(2.0 * np.pi * pert_n_th * (TH - cut_th_min) / (np.pi - 2.0 * cut_th_min))
                             * np.cos(pert_n_ph * PH))
    off = np.abs(TH - np.pi / 2.0)
    rho = np.maximum(R ** -0.5 * np.exp(-off / scale) * pert * inside, 1.0e-8)
    pgas = np.maximum(0.1 * R ** -1.25 * np.exp(-off / scale) * pert ** 2 * inside, 1.0e-9)
    uur = -inflow * R ** -0.5 * np.exp(-off / scale) * inside
    uuth = 0.2 * inflow * np.cos(TH) / R * inside
    uuph = uph_amp * R ** -1.5 * np.exp(-off / scale) * inside
    bbz = 0.02 * np.maximum(R * np.sin(TH), cut_r_min) ** -0.625
    bbr = np.cos(TH) * bbz
    bbth = -np.sin(TH) / R * bbz
    bbph = 0.2 * R ** -1.75 * np.exp(-off / scale) * np.where(TH > np.pi / 2.0, -1.0, 1.0)
    return [np.array(a, np.float64) for a in (rho, pgas, uur, uuth, uuph, bbr, bbth, bbph)]


def mock_fields_cks(n=48, half_width=52.0):
    """Smooth torus-like fields on a uniform Cartesian Kerr-Schild box [-w, w]^3 (simulation_coord = cks):
    x1, x2, x3 = x, y, z; velocity and field components are Cartesian.  Not a physical solution -- a smooth,
    positive, non-symmetric test field for the parity of the cks code paths (reference
    radiation_geometry.cpp:73-91,425-457; the reference ships no Cartesian generator)."""
    xf = np.linspace(-half_width, half_width, n + 1)
    xc = 0.5 * (xf[:-1] + xf[1:])
    Z, Y, X = np.meshgrid(xc, xc, xc, indexing='ij')
    R = np.sqrt(X * X + Y * Y + Z * Z) + 1.0
    cyl = np.sqrt(X * X + Y * Y) + 1.0
    off = np.abs(np.arctan2(Z, cyl))
    wave = 1.0 + 0.1 * np.cos(4.0 * np.arctan2(Y, X)) * np.cos(0.3 * R)
    rho = np.maximum(R ** -0.5 * np.exp(-off / (np.pi / 8.0)) * wave, 1.0e-8)
    pgas = np.maximum(0.1 * R ** -1.25 * np.exp(-off / (np.pi / 8.0)) * wave ** 2, 1.0e-9)
    omega = 0.3 * cyl ** -1.5 * np.exp(-off / (np.pi / 8.0))
    uu1, uu2, uu3 = -Y * omega, X * omega, 0.02 * Z / R
    bz = 0.02 * cyl ** -0.625
    bph = 0.2 * R ** -1.75 * np.where(Z > 0.0, 1.0, -1.0)
    bb1, bb2, bb3 = -Y / cyl * bph + 0.1 * bz * X / R, X / cyl * bph + 0.1 * bz * Y / R, bz
    prim = np.stack([rho, pgas, uu1, uu2, uu3, bb1, bb2, bb3]).astype(np.float32)   # (8, z, y, x)
    return dict(rf=xf, thf=xf.copy(), phf=xf.copy(), r=xc, th=xc.copy(), ph=xc.copy(), prim=prim)


def to_blocks(fields, blocks=(1, 1, 1)):
    """Partition the single-block mock into nb_r x nb_th x nb_ph MeshBlocks (level 0).

    Returns the arrays the Athena++ reader produces: float32 coordinates, prim (8, n_b, n_k, n_j, n_i).
    Block order: x3 slowest, then x2, then x1 (Athena++ Z-order is not needed by either code)."""
    nb_r, nb_th, nb_ph = blocks
    prim = fields['prim']
    n_ph, n_th, n_r = prim.shape[1:]
    assert n_r % nb_r == 0 and n_th % nb_th == 0 and n_ph % nb_ph == 0
    n_i, n_j, n_k = n_r // nb_r, n_th // nb_th, n_ph // nb_ph
    n_b = nb_r * nb_th * nb_ph
    out = dict(n_b=n_b, n_i=n_i, n_j=n_j, n_k=n_k)
    out['x1f'] = np.empty((n_b, n_i + 1), np.float32)
    out['x2f'] = np.empty((n_b, n_j + 1), np.float32)
    out['x3f'] = np.empty((n_b, n_k + 1), np.float32)
    out['x1v'] = np.empty((n_b, n_i), np.float32)
    out['x2v'] = np.empty((n_b, n_j), np.float32)
    out['x3v'] = np.empty((n_b, n_k), np.float32)
    out['prim'] = np.empty((8, n_b, n_k, n_j, n_i), np.float32)
    out['levels'] = np.zeros(n_b, np.int32)
    out['locations'] = np.zeros((n_b, 3), np.int64)
    b = 0
    for bk in range(nb_ph):
        for bj in range(nb_th):
            for bi in range(nb_r):
                si, sj, sk = slice(bi * n_i, (bi + 1) * n_i), slice(bj * n_j, (bj + 1) * n_j), slice(bk * n_k, (bk + 1) * n_k)
                out['x1f'][b] = fields['rf'][bi * n_i:(bi + 1) * n_i + 1]
                out['x2f'][b] = fields['thf'][bj * n_j:(bj + 1) * n_j + 1]
                out['x3f'][b] = fields['phf'][bk * n_k:(bk + 1) * n_k + 1]
                out['x1v'][b] = fields['r'][si]
                out['x2v'][b] = fields['th'][sj]
                out['x3v'][b] = fields['ph'][sk]
                out['prim'][:, b] = prim[:, sk, sj, si]
                out['locations'][b] = (bi, bj, bk)
                b += 1
    out['root_size'] = (n_r, n_th, n_ph)
    return out


# ---------------------------------------------------------------------------------------------
# numpy-only HDF5 writer (superblock v0, one root group with a symbol table, contiguous datasets)

_UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b):
    return b + b'\0' * (-len(b) % 8)


def _datatype(dtype):
    dtype = np.dtype(dtype)
    if dtype.kind == 'i':
        return struct.pack('<BBBBI', 0x10, 0x08, 0, 0, dtype.itemsize) + struct.pack('<HH', 0, 8 * dtype.itemsize)
    if dtype.kind == 'f' and dtype.itemsize == 4:
        return struct.pack('<BBBBI', 0x11, 0x20, 31, 0, 4) + struct.pack('<HHBBBBI', 0, 32, 23, 8, 0, 23, 127)
    if dtype.kind == 'f' and dtype.itemsize == 8:
        return struct.pack('<BBBBI', 0x11, 0x20, 63, 0, 8) + struct.pack('<HHBBBBI', 0, 64, 52, 11, 0, 52, 1023)
    if dtype.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0, 0, 0, dtype.itemsize)
    raise ValueError(dtype)


def _dataspace(shape):
    return struct.pack('<BBB5x', 1, len(shape), 0) + b''.join(struct.pack('<Q', n) for n in shape)


def _message(mtype, body):
    body = _pad8(body)
    return struct.pack('<HHB3x', mtype, len(body), 0) + body


def _object_header(messages):
    body = b''.join(messages)
    return struct.pack('<BBHII4x', 1, 0, len(messages), 1, len(body)) + body


def _attribute(name, value):
    value = np.asarray(value)
    nm = name.encode() + b'\0'
    dt, ds = _datatype(value.dtype), _dataspace(value.shape)
    body = struct.pack('<BBHHH', 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + value.tobytes()
    return _message(12, body)


def write_athdf(path, grid, time=0.0):
    """Write `grid` (output of to_blocks) as an Athena++ .athdf file."""
    n_b = grid['n_b']
    # optional electron-entropy variable (plasma_model = code_kappa): a sixth hydro variable named 'r0'
    hydro = grid['prim'][0:5]
    names = [b'rho', b'press', b'vel1', b'vel2', b'vel3']
    if 'kappa' in grid:
        hydro = np.concatenate([hydro, grid['kappa'][None]], axis=0)
        names.append(b'r0')
    names += [b'Bcc1', b'Bcc2', b'Bcc3']
    datasets = [
        ('B', grid['prim'][5:8]), ('Levels', grid['levels']), ('LogicalLocations', grid['locations']),
        ('prim', hydro), ('x1f', grid['x1f']), ('x1v', grid['x1v']), ('x2f', grid['x2f']),
        ('x2v', grid['x2v']), ('x3f', grid['x3f']), ('x3v', grid['x3v'])]
    n_r, n_th, n_ph = grid['root_size']
    attrs = [
        _attribute('NumCycles', np.int32(0)), _attribute('Time', np.float32(time)),
        _attribute('Coordinates', np.array(b'kerr-schild', dtype='S11')),
        _attribute('RootGridSize', np.array([n_r, n_th, n_ph], np.int32)),
        _attribute('NumMeshBlocks', np.int32(n_b)),
        _attribute('MeshBlockSize', np.array([grid['n_i'], grid['n_j'], grid['n_k']], np.int32)),
        _attribute('MaxLevel', np.int32(int(np.max(grid['levels'])))), _attribute('NumVariables', np.array([len(names) - 3, 3], np.int32)),
        _attribute('DatasetNames', np.array([b'prim', b'B'], dtype='S21')),
        _attribute('VariableNames', np.array(names, dtype='S21'))]

    # layout: [superblock 96][root header][heap header 32][heap data][TREE][SNOD][dataset headers][data]
    heap_data = b'\0' * 8
    name_offsets = []
    for name, _ in datasets:
        name_offsets.append(len(heap_data))
        heap_data += _pad8(name.encode() + b'\0')
    root_header_size = len(_object_header(attrs + [_message(17, struct.pack('<QQ', 0, 0))]))
    root_addr = 96
    heap_addr = root_addr + root_header_size
    heap_data_addr = heap_addr + 32
    tree_addr = heap_data_addr + len(heap_data)
    tree = b'TREE' + struct.pack('<BBHQQ', 0, 0, 1, _UNDEF, _UNDEF)
    snod_addr = tree_addr + len(tree) + 24
    tree += struct.pack('<QQQ', 0, snod_addr, name_offsets[-1])
    snod_size = 8 + 40 * len(datasets)
    header_addr = snod_addr + snod_size
    headers, header_addrs = [], []
    # dataset headers have a fixed size given rank, so compute them twice (addresses, then final)
    sizes = []
    for name, arr in datasets:
        arr = np.ascontiguousarray(arr)
        h = _object_header([_message(1, _dataspace(arr.shape)), _message(3, _datatype(arr.dtype)),
                            _message(8, struct.pack('<BBQQ', 3, 1, 0, arr.nbytes))])
        sizes.append(len(h))
    pos = header_addr
    for s in sizes:
        header_addrs.append(pos)
        pos += s
    data_addrs = []
    for name, arr in datasets:
        pos += -pos % 8
        data_addrs.append(pos)
        pos += np.ascontiguousarray(arr).nbytes
    eof = pos
    for (name, arr), daddr in zip(datasets, data_addrs):
        arr = np.ascontiguousarray(arr)
        headers.append(_object_header([_message(1, _dataspace(arr.shape)), _message(3, _datatype(arr.dtype)),
                                       _message(8, struct.pack('<BBQQ', 3, 1, daddr, arr.nbytes))]))
    snod = b'SNOD' + struct.pack('<BBH', 1, 0, len(datasets))
    for off, haddr in zip(name_offsets, header_addrs):
        snod += struct.pack('<QQII16x', off, haddr, 0, 0)
    root_header = _object_header(attrs + [_message(17, struct.pack('<QQ', tree_addr, heap_addr))])
    assert len(root_header) == root_header_size
    heap = b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), _UNDEF, heap_data_addr)
    superblock = (b'\x89HDF\r\n\x1a\n' + struct.pack('<BBBBBBBB', 0, 0, 0, 0, 0, 8, 8, 0)
                  + struct.pack('<HHI', 4, 16, 0) + struct.pack('<QQQQ', 0, _UNDEF, eof, _UNDEF)
                  + struct.pack('<QQII', 0, root_addr, 1, 0) + struct.pack('<QQ', tree_addr, heap_addr))
    assert len(superblock) == 96
    with open(path, 'wb') as f:
        f.write(superblock + root_header + heap + heap_data + tree + snod + b''.join(headers))
        for (name, arr), daddr in zip(datasets, data_addrs):
            f.seek(daddr)
            f.write(np.ascontiguousarray(arr).tobytes())
        f.truncate(eof)


def write_h5_tree(path, tree):
    """Write nested dicts {name: ndarray | dict} as an HDF5 file of old-style groups (symbol table + local heap +
    one B-tree leaf per group, v1 object headers whose first message is the symbol-table message, contiguous
    datasets; 0-d arrays become scalar dataspaces) -- the subset the reference's parser walks for 'a/b/c' paths
    (hdf5_format_metadata.cpp:29-160)."""
    out = bytearray(96)

    def put(b):
        out.extend(b'\0' * (-len(out) % 8))
        addr = len(out)
        out.extend(b)
        return addr

    def dataset(arr):
        arr = np.ascontiguousarray(arr).reshape(np.shape(arr))     # ascontiguousarray promotes 0-d to 1-d
        daddr = put(arr.tobytes() if arr.nbytes else b'\0' * 8)
        return put(_object_header([_message(1, _dataspace(arr.shape)), _message(3, _datatype(arr.dtype)),
                                   _message(8, struct.pack('<BBQQ', 3, 1, daddr, arr.nbytes))]))

    def group(node):
        names = sorted(node)
        headers = [group(node[n]) if isinstance(node[n], dict) else (dataset(np.asarray(node[n])), None, None) for n in names]
        heap_data, offs = b'\0' * 8, []
        for n in names:
            offs.append(len(heap_data))
            heap_data += _pad8(n.encode() + b'\0')
        heap_data_addr = put(heap_data)
        heap_addr = put(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), _UNDEF, heap_data_addr))
        snod = b'SNOD' + struct.pack('<BBH', 1, 0, len(names))
        for off, (haddr, bt, hp) in zip(offs, headers):
            snod += struct.pack('<QQII', off, haddr, 0 if bt is None else 1, 0)
            snod += struct.pack('<QQ', bt, hp) if bt is not None else b'\0' * 16
        snod_addr = put(snod)
        tree_addr = put(b'TREE' + struct.pack('<BBHQQ', 0, 0, 1, _UNDEF, _UNDEF) + struct.pack('<QQQ', 0, snod_addr, offs[-1]))
        header_addr = put(_object_header([_message(17, struct.pack('<QQ', tree_addr, heap_addr))]))
        return header_addr, tree_addr, heap_addr

    root_addr, tree_addr, heap_addr = group(tree)
    eof = len(out)
    out[0:96] = (b'\x89HDF\r\n\x1a\n' + struct.pack('<BBBBBBBB', 0, 0, 0, 0, 0, 8, 8, 0)
                 + struct.pack('<HHI', 4, 16, 0) + struct.pack('<QQQQ', 0, _UNDEF, eof, _UNDEF)
                 + struct.pack('<QQII', 0, root_addr, 1, 0) + struct.pack('<QQ', tree_addr, heap_addr))
    with open(path, 'wb') as f:
        f.write(bytes(out))


def fmks_theta(x1, x2, hslope, r_in, poly_xt, poly_alpha, mks_smooth):
    """theta(x1, x2) of the FMKS coordinates (reference simulation_geometry.cpp:416-434)."""
    poly_norm = (poly_alpha + 1.0) * poly_xt ** poly_alpha
    poly_norm = 0.5 * np.pi * poly_norm / (poly_norm + 1.0)
    y = 2.0 * x2 - 1.0
    theta_g = np.pi * x2 + (1.0 - hslope) / 2.0 * np.sin(2.0 * np.pi * x2)
    theta_j = 0.5 * np.pi + poly_norm * y * (1.0 + (y / poly_xt) ** poly_alpha / (poly_alpha + 1.0))
    return theta_g + np.exp(mks_smooth * (np.log(r_in) - x1)) * (theta_j - theta_g)


def write_iharm3d(path, n_r=48, n_th=32, n_ph=32, gamma_adi=4.0 / 3.0, time=0.0, spin=0.0, hslope=1.0, fmks=None,
                  r_min=None, r_max=None, **model):
    """iharm3d-format HDF5 dump of the mock torus (reference scripts/generate_mock_simulation.py:79-157,200-243 for
    the MKS case): header/{n1,n2,n3,gam,metric,n_prim,prim_names,geom/...}, t, and prims (n1, n2, n3, 8) float32 =
    RHO, UU, normal-frame velocity U1-3 and lab-frame field B1-3 in the modified coordinates.
    fmks = dict(poly_xt, poly_alpha, mks_smooth) writes FMKS coordinates instead (theta depends on x1 and x2); the
    closed-form model is then evaluated at each cell's own (r, theta).  The reference ships no FMKS generator.
    Spin 0 only (the generator's metric is Schwarzschild in Kerr-Schild form)."""
    assert spin == 0.0
    if r_min is None:
        r_min = 2.0 * 25.0 ** (-1.0 / 75.0)
    if r_max is None:
        r_max = 2.0 * 25.0 ** (76.0 / 75.0)
    lrf = np.linspace(np.log(r_min), np.log(r_max), n_r + 1)
    x2f = np.linspace(0.0, 1.0, n_th + 1)
    phf = np.linspace(0.0, 2.0 * np.pi, n_ph + 1)
    lr, x2, ph = (0.5 * (f[:-1] + f[1:]) for f in (lrf, x2f, phf))
    X1, X2 = np.meshgrid(lr, x2, indexing='xy')            # (th, r)
    R = np.exp(X1)
    if fmks is None:
        TH = np.pi * X2 + (1.0 - hslope) / 2.0 * np.sin(2.0 * np.pi * X2)
        dth_dx1 = np.zeros_like(TH)
        dth_dx2 = np.pi + (1.0 - hslope) * np.pi * np.cos(2.0 * np.pi * X2)
    else:
        args = (hslope, r_min, fmks['poly_xt'], fmks['poly_alpha'], fmks['mks_smooth'])
        TH = fmks_theta(X1, X2, *args)
        eps = 1.0e-6
        dth_dx1 = (fmks_theta(X1 + eps, X2, *args) - fmks_theta(X1 - eps, X2, *args)) / (2.0 * eps)
        dth_dx2 = (fmks_theta(X1, X2 + eps, *args) - fmks_theta(X1, X2 - eps, *args)) / (2.0 * eps)
    prim = mock_fields_at(R[None], TH[None], ph[:, None, None], **model)       # 8 x (ph, th, r), float64
    rho, pgas, uur, uuth, uuph, bbr, bbth, bbph = prim
    R3, TH3 = R[None], TH[None]
    f = 2.0 / R3
    g_tr, g_rr, g_thth, g_phph = f, 1.0 + f, R3 ** 2, R3 ** 2 * np.sin(TH3) ** 2
    gtt, gtr = -(1.0 + f), f
    alpha = 1.0 / np.sqrt(-gtt)
    # standard normal frame -> standard coordinate frame
    uut = np.sqrt(1.0 + g_rr * uur ** 2 + g_thth * uuth ** 2 + g_phph * uuph ** 2)
    ut, ur, uth, uph = uut / alpha, uur - alpha * uut * gtr, uuth, uuph
    u_r, u_th, u_ph = g_tr * ut + g_rr * ur, g_thth * uth, g_phph * uph
    bt = u_r * bbr + u_th * bbth + u_ph * bbph
    br, bth, bph = (bbr + bt * ur) / ut, (bbth + bt * uth) / ut, (bbph + bt * uph) / ut
    # standard -> modified coordinate frame: dr = r dx1, dth = dth_dx1 dx1 + dth_dx2 dx2
    a1, a2 = dth_dx1[None], dth_dx2[None]
    u1, u3 = ur / R3, uph
    u2 = (uth - a1 * u1) / a2
    b1, b3 = br / R3, bph
    b2 = (bth - a1 * b1) / a2
    # modified coordinate frame -> modified normal frame primitives: g^{0i} of the modified coordinates
    g01 = gtr / R3
    g02 = -a1 * gtr / (R3 * a2)
    uu1, uu2, uu3 = u1 + alpha ** 2 * g01 * ut, u2 + alpha ** 2 * g02 * ut, u3
    bb1, bb2, bb3 = b1 * ut - bt * u1, b2 * ut - bt * u2, b3 * ut - bt * u3
    prims = np.array([rho, pgas / (gamma_adi - 1.0), uu1, uu2, uu3, bb1, bb2, bb3], dtype=np.float32).transpose()
    name = 'MKS' if fmks is None else 'FMKS'
    S20 = lambda *v: np.array(v, dtype='S20')
    f64, i32 = (lambda v: np.array(v, np.float64)), (lambda v: np.array(v, np.int32))
    params = {'a': f64(spin), 'hslope': f64(hslope), 'r_eh': f64(2.0), 'r_in': f64(r_min), 'r_out': f64(r_max)}
    if fmks is not None:
        params.update({k: f64(fmks[k]) for k in ('poly_xt', 'poly_alpha', 'mks_smooth')})
    tree = {'header': {'version': S20(b'iharm-blacklight'), 'gam': f64(gamma_adi), 'tf': f64(0.0), 'n1': i32(n_r), 'n2': i32(n_th),
                       'n3': i32(n_ph), 'metric': S20(name.encode()), 'n_prim': i32(8),
                       'prim_names': S20(b'RHO', b'UU', b'U1', b'U2', b'U3', b'B1', b'B2', b'B3'), 'has_electrons': i32(0),
                       'geom': {'dx1': f64(lrf[1] - lrf[0]), 'dx2': f64(x2f[1] - x2f[0]), 'dx3': f64(phf[1] - phf[0]),
                                'startx1': f64(lrf[0]), 'startx2': f64(x2f[0]), 'startx3': f64(phf[0]), 'n_dim': i32(4),
                                name.lower(): params}},
            't': f64(time), 'prims': prims}
    write_h5_tree(path, tree)
    return dict(lrf=lrf, x2f=x2f, phf=phf, r=R, th=TH, prims=prims)


def write_harm3d(path, fields, gamma_adi=4.0 / 3.0, time=0.0):
    """Write the single-block mock (output of mock_fields) in the 'harm3d' ascii-header + binary format
    (reference scripts/generate_mock_simulation.py:79-157,245-279): modified Kerr-Schild coordinates
    x1 = ln r, x2 = theta / pi (h = 1), x3 = phi; per cell 6 coordinates, rho, u_gas and the coordinate-frame
    four-velocity and magnetic four-vector, float32, with the variable index fastest."""
    rf, thf, phf, r, th, ph = (fields[k] for k in ('rf', 'thf', 'phf', 'r', 'th', 'ph'))
    rho, pgas, uur, uuth, uuph, bbr, bbth, bbph = (fields['prim'][q].astype(np.float64) for q in range(8))
    R, TH = r[None, None, :], th[None, :, None]
    sigma = R ** 2
    f = 2.0 * R / sigma
    g_tt, g_tr, g_rr, g_thth, g_phph = -(1.0 - f), f, 1.0 + f, sigma, R ** 2 * np.sin(TH) ** 2
    gtt, gtr = -(1.0 + f), f
    alpha = 1.0 / np.sqrt(-gtt)
    ugas = pgas / (gamma_adi - 1.0)
    uut = np.sqrt(1.0 + g_rr * uur ** 2 + g_thth * uuth ** 2 + g_phph * uuph ** 2)
    ut = uut / alpha
    ur = uur - alpha * uut * gtr
    uth, uph = uuth, uuph
    u_r = g_tr * ut + g_rr * ur
    u_th = g_thth * uth
    u_ph = g_phph * uph
    u0, u1, u2, u3 = ut, ur / R, uth / np.pi, uph
    bt = u_r * bbr + u_th * bbth + u_ph * bbph
    br = (bbr + bt * ur) / ut
    bth = (bbth + bt * uth) / ut
    bph = (bbph + bt * uph) / ut
    b0, b1, b2, b3 = bt, br / R, bth / np.pi, bph
    lrf, lr = np.log(rf), np.log(r)
    x2f, x2 = thf / np.pi, th / np.pi
    dlr, dx2, dph = lrf[1] - lrf[0], x2f[1] - x2f[0], phf[1] - phf[0]
    shape = rho.shape
    data = [np.broadcast_to(lr[None, None, :], shape), np.broadcast_to(x2[None, :, None], shape),
            np.broadcast_to(ph[:, None, None], shape), np.broadcast_to(r[None, None, :], shape),
            np.broadcast_to(th[None, :, None], shape), np.broadcast_to(ph[:, None, None], shape),
            rho, ugas, u0 * np.ones(shape), u1 * np.ones(shape), u2 * np.ones(shape), u3 * np.ones(shape),
            b0 * np.ones(shape), b1 * np.ones(shape), b2 * np.ones(shape), b3 * np.ones(shape)]
    with open(path, 'w') as f_out:
        f_out.write('{0:24.16e} '.format(time))
        f_out.write('{0} {1} {2} '.format(len(r), len(th), len(ph)))
        f_out.write('{0:24.16e} {1:24.16e} {2:24.16e} '.format(lrf[0], x2f[0], phf[0]))
        f_out.write('{0:24.16e} {1:24.16e} {2:24.16e} '.format(dlr, dx2, dph))
        f_out.write('0.0 ')
        f_out.write('{0:24.16e} '.format(gamma_adi))
        f_out.write('{0:24.16e} '.format(rf[0]))
        f_out.write('1.0 ')
        f_out.write('8\n')
        np.array(data, dtype=np.float32).transpose().tofile(f_out)


def write_athenak(path, grid, gamma_adi=4.0 / 3.0, time=0.0, location_size=4, variable_size=4, spin=0.0,
                  extra_first=('spare',)):
    """Write a uniform-block Cartesian grid (output of to_blocks on mock_fields_cks) as an AthenaK binary dump,
    version 1.1, the way the reference's reader takes it apart (simulation_reader.cpp:434-588,915-1131): ascii
    pre-header (time, sizes, variable names, header offset), the run's input parameters as text (<coord> a,
    <mhd> gamma are looked at), then per MeshBlock 6 int32 cell index bounds, 3 int32 logical location + int32
    level, 6 face positions (x1min, x1max, x2min, ...) of location_size bytes and the cell data variable by
    variable.  Pressure is stored as internal energy `eint`; `extra_first` puts unrelated variables in front so
    that the by-name lookup is exercised.  The reference ships no AthenaK generator."""
    names = list(extra_first) + ['dens', 'velx', 'vely', 'velz', 'eint', 'bcc1', 'bcc2', 'bcc3']
    prim = grid['prim'].astype(np.float64)
    src = {'dens': prim[0], 'eint': prim[1] / (gamma_adi - 1.0), 'velx': prim[2], 'vely': prim[3], 'velz': prim[4],
           'bcc1': prim[5], 'bcc2': prim[6], 'bcc3': prim[7]}
    if 'kappa' in grid:
        names.append('r0')
        src['r0'] = grid['kappa'].astype(np.float64)
    params = ('# mock AthenaK parameter dump\n<coord>\ngeneral_rel = true\na = %.17g\n<mhd>\neos = ideal\n'
              'gamma = %.17g\n<problem>\nuser_hist = false\n' % (spin, gamma_adi)).encode()
    head = ('Athena binary output version=1.1\n  size of preheader=5\n  time=%.16e\n  cycle=0\n'
            '  size of location=%d\n  size of variable=%d\n  number of variables=%d\n  variables:  %s  \n'
            '  header offset=%d\n' % (time, location_size, variable_size, len(names), ' '.join(names), len(params))).encode()
    loc_t = np.float32 if location_size == 4 else np.float64
    var_t = np.float32 if variable_size == 4 else np.float64
    n_i, n_j, n_k = grid['n_i'], grid['n_j'], grid['n_k']
    with open(path, 'wb') as f:
        f.write(head + params)
        for b in range(grid['n_b']):
            li, lj, lk = (int(v) for v in grid['locations'][b])
            np.array([li * n_i, (li + 1) * n_i - 1, lj * n_j, (lj + 1) * n_j - 1, lk * n_k, (lk + 1) * n_k - 1,
                      li, lj, lk, int(grid['levels'][b])], np.int32).tofile(f)
            np.array([grid['x1f'][b, 0], grid['x1f'][b, -1], grid['x2f'][b, 0], grid['x2f'][b, -1],
                      grid['x3f'][b, 0], grid['x3f'][b, -1]], loc_t).tofile(f)
            for name in names:
                data = src[name][b] if name in src else np.full((n_k, n_j, n_i), 1.0e30)
                np.ascontiguousarray(data, var_t).tofile(f)


def grid_view_arrays(grid):
    """Arrays in the layout SimulationReader hands to the integrator (float32 coords widened to f64)."""
    if 'kappa' in grid:   # the reader stacks hydro (with the entropy variable last) before the field
        prim = np.concatenate([grid['prim'][0:5], grid['kappa'][None], grid['prim'][5:8]], axis=0)
        out = grid_view_arrays({k: v for k, v in grid.items() if k != 'kappa'})
        out.update(n_var=9, prim=np.ascontiguousarray(prim, np.float32), ind_kappa=5, ind_bb1=6, ind_bb2=7, ind_bb3=8)
        return out
    return dict(
        n_b=grid['n_b'], n_k=grid['n_k'], n_j=grid['n_j'], n_i=grid['n_i'], n_var=8,
        levels=np.ascontiguousarray(grid['levels'], np.int32),
        locations=np.ascontiguousarray(grid['locations'], np.int64).astype(np.int32),
        x1f=grid['x1f'].astype(np.float64), x2f=grid['x2f'].astype(np.float64), x3f=grid['x3f'].astype(np.float64),
        x1v=grid['x1v'].astype(np.float64), x2v=grid['x2v'].astype(np.float64), x3v=grid['x3v'].astype(np.float64),
        prim=np.ascontiguousarray(grid['prim'], np.float32),
        ind_rho=0, ind_pgas=1, ind_kappa=-1, ind_uu1=2, ind_uu2=3, ind_uu3=4, ind_bb1=5, ind_bb2=6, ind_bb3=7,
        n_3_root=grid['root_size'][2])


def to_blocks_amr(n_r, n_th, n_ph, blocks, refine, **kwargs):
    """Two-level mesh: the root grid of `blocks` MeshBlocks at level 0, with the root blocks for which
    refine(bi, bj, bk) is true replaced by their 8 children at level 1 (same cells per block, half the spacing).
    Both levels are independent evaluations of the same closed-form model (mock_fields) at their own cell
    centres -- what an AMR code would produce for a smooth field.  Exercises the inter-block interpolation
    (reference simulation_sampling.cpp:1068-1321) across same-level, coarser and finer neighbours."""
    coarse = to_blocks(mock_fields(n_r=n_r, n_th=n_th, n_ph=n_ph, **kwargs), blocks)
    fine = to_blocks(mock_fields(n_r=2 * n_r, n_th=2 * n_th, n_ph=2 * n_ph, **kwargs),
                     tuple(2 * b for b in blocks))
    fine_id = {tuple(int(v) for v in loc): b for b, loc in enumerate(fine['locations'])}
    keys = ('x1f', 'x2f', 'x3f', 'x1v', 'x2v', 'x3v')
    rows = {k: [] for k in keys}
    prims, levels, locs = [], [], []
    for b, loc in enumerate(coarse['locations']):
        bi, bj, bk = (int(v) for v in loc)
        if refine(bi, bj, bk):
            for dk in (0, 1):
                for dj in (0, 1):
                    for di in (0, 1):
                        child = (2 * bi + di, 2 * bj + dj, 2 * bk + dk)
                        fb = fine_id[child]
                        for k in keys:
                            rows[k].append(fine[k][fb])
                        prims.append(fine['prim'][:, fb])
                        levels.append(1)
                        locs.append(child)
        else:
            for k in keys:
                rows[k].append(coarse[k][b])
            prims.append(coarse['prim'][:, b])
            levels.append(0)
            locs.append((bi, bj, bk))
    out = dict(n_b=len(levels), n_i=coarse['n_i'], n_j=coarse['n_j'], n_k=coarse['n_k'])
    for k in keys:
        out[k] = np.stack(rows[k]).astype(np.float32)
    out['prim'] = np.stack(prims, axis=1).astype(np.float32)
    out['levels'] = np.array(levels, np.int32)
    out['locations'] = np.array(locs, np.int64)
    out['root_size'] = coarse['root_size']
    return out


def add_entropy(grid, scale=2.0e7):
    """Electron-entropy variable kappa_e ~ p / rho^(4/3), a smooth positive field for plasma_model = code_kappa
    (theta_e = (sqrt(1 + 25 (rho_e kappa)^(2/3)) - 1) / 5).  The default scale gives theta_e of order 1-100 in the
    torus; much colder electrons (scale <= 1e5) make the plasma Faraday-thick by many orders of magnitude per
    step, where the polarized solution of ANY implementation is limited by the rounding of sin/cos of 1e9."""
    rho, pgas = grid['prim'][0].astype(np.float64), grid['prim'][1].astype(np.float64)
    grid['kappa'] = (scale * pgas / rho ** (4.0 / 3.0)).astype(np.float32)
    return grid


def make_mock(path=None, blocks=(1, 1, 1), refine=None, cks=None, entropy=False, **kwargs):
    if cks is not None:
        grid = to_blocks(mock_fields_cks(**cks), blocks)
    elif refine is not None:
        n_r, n_th, n_ph = kwargs.pop('n_r', 77), kwargs.pop('n_th', 64), kwargs.pop('n_ph', 128)
        grid = to_blocks_amr(n_r, n_th, n_ph, blocks, refine, **kwargs)
    else:
        grid = to_blocks(mock_fields(**kwargs), blocks)
    if entropy:
        add_entropy(grid)
    if path is not None:
        write_athdf(path, grid)
    return grid


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('filename')
    ap.add_argument('--n_r', type=int, default=77)
    ap.add_argument('--n_th', type=int, default=64)
    ap.add_argument('--n_ph', type=int, default=128)
    ap.add_argument('--blocks', type=int, nargs=3, default=(1, 1, 1))
    ap.add_argument('--pert_amp', type=float, default=0.1)
    ap.add_argument('--pert_n_ph', type=int, default=4)
    a = ap.parse_args()
    make_mock(a.filename, tuple(a.blocks), n_r=a.n_r, n_th=a.n_th, n_ph=a.n_ph, pert_amp=a.pert_amp, pert_n_ph=a.pert_n_ph)
