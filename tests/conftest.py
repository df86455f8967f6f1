import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


# Tolerances -------------------------------------------------------------------------------
# The reference's own golden-vector test (numba_osc_tests.py:82) uses
#   atol = 10 * finfo.resolution (1e-14 f8 / 1e-5 f4), rtol = 100 * ALLCLOSE_KW.rtol (1e-10 / 1e-3)
AC_KW_F8 = dict(atol=1e-14, rtol=1e-10)
AC_KW_F4 = dict(atol=1e-5, rtol=1e-3)
