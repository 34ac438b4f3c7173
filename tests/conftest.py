import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def golden(name):
    return np.load(os.path.join(GOLD, name + '.npz'))


def relerr(a, b):
    a, b = np.asarray(a), np.asarray(b)
    s = np.max(np.abs(b)) if b.size else 1.0
    d = np.max(np.abs(a - b)) if a.size else 0.0
    return d / s if s > 0 else d


@pytest.fixture(scope='session')
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda', 0)
