import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
if ORACLE_DIR not in sys.path:
    sys.path.insert(0, ORACLE_DIR)

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(GOLDEN)


MULTIBAND = os.path.join(ROOT, "tests", "golden", "multiband_golden.npz")


@pytest.fixture(scope="session")
def multiband():
    """Outputs of the REAL reference on its own 4-band Hamiltonians (oracle/make_golden_multiband.py)."""
    import numpy as np
    return np.load(MULTIBAND)


CONFIG0 = os.path.join(ROOT, "tests", "golden", "config0_golden.npz")


@pytest.fixture(scope="session")
def config0():
    """BASELINE config 0 (1-D well, 1024 points, Gauss-Seidel V-cycles) run by the REAL reference
    (oracle/make_golden_config0.py)."""
    import numpy as np
    return np.load(CONFIG0)
