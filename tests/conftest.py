"""pytest configuration: markers and import path."""

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def regression_data():
    """The data set the reference's QNN / QRC golden tests use (tests/qnn/test_qnnr.py:20-27)."""
    from sklearn.datasets import make_regression
    from sklearn.preprocessing import MinMaxScaler

    X, y = make_regression(n_samples=6, n_features=1, random_state=42)
    X = MinMaxScaler((0.1, 0.9)).fit_transform(X, y)
    return np.asarray(X, dtype=np.float64), np.asarray(y, dtype=np.float64)
