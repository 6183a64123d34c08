import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def reference_root():
    """Directory holding the reference `raleigh` package for tests that drive the
    reference's own solver: baseline/_ref travels to the GPU box, /root/reference
    exists only in the build container (CPU tests only)."""
    for cand in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
        if os.path.isdir(os.path.join(cand, 'raleigh', 'core')):
            return cand
    return None


@pytest.fixture(scope='session')
def ref_root():
    r = reference_root()
    if r is None:
        pytest.skip('reference package not available on this box')
    return r


@pytest.fixture(scope='session')
def gpu_backend():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import raleigh_b200
    return raleigh_b200
