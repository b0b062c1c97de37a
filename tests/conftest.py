"""pytest configuration: the `gpu` marker, repo root on sys.path, golden-fixture loaders."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope='session')
def gold():
    return golden


def keymap(keys):
    return {str(k): i for i, k in enumerate(keys)}


def relerr(a, r):
    """Elementwise relative error; equal values (incl. 0 == 0) and NaN == NaN count as 0."""
    a, r = np.asarray(a, dtype=np.float64), np.asarray(r, dtype=np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        e = np.abs(a - r) / np.abs(r)
    e[a == r] = 0.0
    e[np.isnan(a) & np.isnan(r)] = 0.0
    return e


def formalisms_of(atm_npz):
    return [(str(c), str(f)) for c, f in zip(atm_npz['alpha_constituents'], atm_npz['alpha_formalisms']) if str(f) != 'none']


TRUNC = {'h2s': 1e-22, 'ph3': 1e-22}
