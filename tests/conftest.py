import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def load_golden(name: str) -> np.ndarray:
    with np.load(os.path.join(GOLDEN_DIR, f'{name}.npz')) as z:
        return z['out']


@pytest.fixture(scope='session')
def ns():
    from oracle import cases
    return cases.b200_namespace()


@pytest.fixture(scope='session')
def engine():
    from signals_b200 import engine as engine_mod
    return engine_mod.Engine()


def max_abs_err(got: np.ndarray, want: np.ndarray) -> float:
    """max |got - want| treating NaN==NaN as equal and mismatched NaNs as inf."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    if (nan_g != nan_w).any():
        return float('inf')
    d = np.abs(np.where(nan_w, 0.0, got - np.where(nan_w, 0.0, want)))
    return float(d.max()) if d.size else 0.0
