"""pisa_b200 -- B200-native oscillation reweighting + histogramming behind the PISA stage API.

Process-wide globals mirror ``pisa/__init__.py`` (reference :152-273): the float type is chosen
once from ``PISA_FTYPE``; the only compute target is the hand-written sm_100a CUDA library
(``pisa_b200/libpisa_b200.so``, C ABI in ``include/pisa_b200.h``).  There is no CPU fallback.
"""
import os

import numpy as np

__version__ = "0.1.0"

_FTYPE_NAMES = {
    "float32": np.float32, "fp32": np.float32, "single": np.float32, "32": np.float32,
    "float64": np.float64, "fp64": np.float64, "double": np.float64, "64": np.float64,
}
FTYPE = _FTYPE_NAMES.get(os.environ.get("PISA_FTYPE", "fp64").strip().lower())
if FTYPE is None:
    raise ValueError("PISA_FTYPE=%r not understood; use fp32 or fp64" % os.environ.get("PISA_FTYPE"))
CTYPE = np.complex64 if FTYPE == np.float32 else np.complex128
ITYPE = np.int32 if FTYPE == np.float32 else np.int64
HASH_SIGFIGS = 12
TARGET = "cuda"  # the only target of this implementation (reference: cpu / parallel / cuda)

RESOURCES_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "resources")

# `from pisa import ureg, Q_` (reference pisa/__init__.py:59-64): the unit registry of this package
from pisa_b200.utils.units import Quantity as Q_, ureg  # noqa: E402  pylint: disable=wrong-import-position
