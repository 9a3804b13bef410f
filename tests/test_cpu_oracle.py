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


@pytest.mark.parametrize('name', ['formula_16', 'formula_pinhole_pole_12'])
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


@pytest.mark.parametrize('name', ['simulation_32', 'simulation_nearest_24', 'simulation_blocks_24', 'simulation_kerr_24'])
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
