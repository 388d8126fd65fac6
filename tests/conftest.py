import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """Fixtures frozen from the real reference by oracle/make_golden.py (tests/golden/)."""

    def __init__(self):
        d = os.path.join(ROOT, 'tests', 'golden')
        self.arrays = np.load(os.path.join(d, 'golden.npz'))
        with open(os.path.join(d, 'golden.json')) as f:
            self.meta = json.load(f)

    def cases(self, kind):
        return sorted(k for k, v in self.meta['cases'].items() if v['kind'] == kind)

    def params(self, name):
        return self.meta['cases'][name]['params']

    def get(self, name, arr):
        return self.arrays['%s/%s' % (name, arr)]


_GOLDEN = None


def golden_data():
    global _GOLDEN
    if _GOLDEN is None:
        _GOLDEN = Golden()
    return _GOLDEN


@pytest.fixture(scope='session')
def golden():
    return golden_data()


def seeded(seed, *shape):
    """Same recipe as oracle/make_golden.py::seeded."""
    g = np.random.default_rng(seed)
    return g.standard_normal(shape) + 1j * g.standard_normal(shape)


def seeded_typed(seed, dt, *shape):
    """Same recipe as oracle/make_golden.py::seeded_typed."""
    g = np.random.default_rng(seed)
    if dt.startswith('int'):
        info = np.iinfo(dt)
        return g.integers(info.min, info.max, size=shape, dtype=np.int64, endpoint=True).astype(dt)
    if dt.startswith('float'):
        return g.standard_normal(shape).astype(dt)
    return (g.standard_normal(shape) + 1j * g.standard_normal(shape)).astype(dt)


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))
