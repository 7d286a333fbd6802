import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope='session')
def have_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def epe(a, b):
    d = np.asarray(a, np.float64) - np.asarray(b, np.float64)
    return np.sqrt((d * d).sum(-1))
