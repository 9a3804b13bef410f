import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session')
def ref_bin():
    """The unmodified reference built by oracle/Makefile (travels to the GPU box, not in git)."""
    path = os.path.join(ROOT, 'oracle', '_ref', 'blacklight')
    if not os.path.exists(path):
        pytest.skip('oracle/_ref/blacklight not built (run __graft_entry__.build() where /root/reference exists)')
    return path


@pytest.fixture(scope='session')
def lib():
    import blacklight_b200
    return blacklight_b200.load_library()


@pytest.fixture(scope='session')
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail('GPU test selected but no CUDA device is visible (blacklight_b200 has no CPU fallback)')
    return torch.device('cuda:0')
