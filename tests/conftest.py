import os
import sys

import pytest

# compute_dtype=float32 parity tests compare against the oracle at fp32 tolerances: exact FFMA products.  The
# library default ('tf32', like XLA:GPU) is covered by tests/test_tf32_gpu.py, which sets the precision itself.
os.environ.setdefault('MLB_MATMUL_PRECISION', 'highest')

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope='session')
def mlb():
    import madrona_learn_b200 as m
    return m
