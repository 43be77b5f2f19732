import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden', 'gd_golden.npz')


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    """Golden vectors produced by the UNMODIFIED reference (oracle/make_golden.py)."""
    z = np.load(GOLDEN)
    manifest = json.loads(bytes(z['manifest']).decode())
    return z, manifest
