import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "large-velocity-power-spectrum_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))


@pytest.fixture(scope="session")
def orc():
    import vpower_oracle
    return vpower_oracle


@pytest.fixture(scope="session")
def golden_fold():
    """Reference outputs of the folding stage (tests/golden/make_golden_fold.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden_fold.npz"))


def fold_case_field(seed, N):
    """The velocity / mass arrays of a fold golden case, from its seed (same recipe as make_golden_fold.py)."""
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(N, N, N, 3)).astype(np.float32)
    x = (np.arange(N) + 0.5) / N
    v[..., 0] += (2.0 * np.sin(2 * np.pi * 3 * x)[:, None, None]).astype(np.float32)
    v[..., 2] += (1.5 * np.cos(2 * np.pi * 5 * x)[None, :, None]).astype(np.float32)
    mass = (1.0 + rng.random((N, N, N))).astype(np.float32)
    return v, mass
