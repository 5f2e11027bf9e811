import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(autouse=True)
def _set_sample_rate():
    # same autouse fixture as the reference's tests/conftest.py:5-9
    import pygmu2_b200 as pg
    pg.set_sample_rate(44100)
    yield
    pg.set_sample_rate(44100)


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name))


def rel_err(y, ref):
    """north_star tolerance metric: max-abs error over full scale of the reference."""
    import numpy as np
    y = np.asarray(y, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(y - ref)) / max(np.max(np.abs(ref)), 1e-30))
