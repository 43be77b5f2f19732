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


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without CUDA: gpu-marked tests are skipped, not errors.  An
    explicit `-m gpu` selection is left alone: there a missing device must FAIL, not skip."""
    expr = config.getoption('-m') or ''
    if 'gpu' in expr and 'not gpu' not in expr:
        return
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason='no CUDA device (gpu-marked test)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    """Golden vectors produced by the UNMODIFIED reference (oracle/make_golden.py)."""
    z = np.load(GOLDEN)
    manifest = json.loads(bytes(z['manifest']).decode())
    return z, manifest
