"""Load the *unmodified* reference hot-path modules from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by ``make_golden.py`` (fixture generation) and
by the optional ``-m "not gpu"`` cross-checks that run in the build container.
Nothing on the product path, in ``smoke()`` or in ``bench.py`` imports this:
``/root/reference`` does not exist on the GPU box.

``import pisa`` fails here (pint / uncertainties / fast_histogram / h5py are
not installed), but the hot-path arithmetic only needs a handful of names from
the package root, so a ~30-line stub package is generated in a temp dir and the
reference *files* are symlinked into it (never copied into this repo):

    pisa/utils/numba_tools.py
    pisa/stages/osc/prob3numba/numba_osc_kernels.py, numba_osc_hostfuncs.py
    pisa/stages/osc/layers.py, osc_params.py, nsi_params.py
    pisa/core/translation.py, bin_indexing.py   (binning / fast_histogram stubbed)
    pisa/stages/flux/barr_simple.py, pisa/utils/barr_parameterization.py   (Stage / Param stubbed)

The float type is fixed at import time by the reference (``PISA_FTYPE``), so
FP32 fixtures are produced in a separate process.
"""
import importlib
import os
import sys
import tempfile
import textwrap

REFERENCE_ROOT = os.environ.get("PISA_REFERENCE_ROOT", "/root/reference")

_STUB_INIT = '''
import os
import numpy as np
FTYPE = np.float32 if os.environ.get("PISA_FTYPE", "fp64") in ("fp32", "float32", "single") else np.float64
CTYPE = np.complex64 if FTYPE == np.float32 else np.complex128
ITYPE = np.int32 if FTYPE == np.float32 else np.int64
HASH_SIGFIGS = 12
TARGET = os.environ.get("PISA_TARGET", "cpu")
PISA_NUM_THREADS = int(os.environ.get("PISA_NUM_THREADS", os.cpu_count() if TARGET == "parallel" else 1))
PISA_HIST_THREADING = "off"
EPSILON = 1e-9
class _U:
    dimensionless = 1.0
    def __call__(self, *a, **k): return 1.0
    def __getattr__(self, k): return 1.0
ureg = _U()
'''

_STUBS = {
    "pisa/utils/__init__.py": "",
    "pisa/core/__init__.py": "",
    "pisa/stages/__init__.py": "",
    "pisa/stages/osc/__init__.py": "",
    "pisa/stages/osc/prob3numba/__init__.py": "",
    "pisa/utils/comparisons.py": '''
import numpy as np
from pisa import FTYPE, HASH_SIGFIGS
FTYPE_PREC = np.finfo(FTYPE).eps
FTYPE_SIGFIGS = int(np.abs(np.ceil(np.log10(FTYPE_PREC))))
EQUALITY_SIGFIGS = min(HASH_SIGFIGS, FTYPE_SIGFIGS)
EQUALITY_PREC = 10**-EQUALITY_SIGFIGS
ALLCLOSE_KW = dict(rtol=EQUALITY_PREC, atol=FTYPE_PREC, equal_nan=True)
def recursiveEquality(a, b): return np.allclose(a, b, **ALLCLOSE_KW)
def isscalar(x): return np.isscalar(x)
''',
    "pisa/utils/log.py": '''
import logging
logging.trace = logging.debug
class Levels: DEBUG=2; INFO=1; WARN=0; TRACE=3
def set_verbosity(v): pass
''',
    "pisa/utils/fileio.py": '''
import numpy as np
def from_file(fname, as_array=False, **kw):
    return np.loadtxt(fname)
''',
    "pisa/utils/profiler.py": "def profile(f): return f\n",
    "pisa/utils/resources.py": '''
import os
RES = os.path.join(os.environ.get("PISA_REFERENCE_ROOT", "/root/reference"), "pisa_examples", "resources")
def find_resource(name, fail=True):
    return name if os.path.exists(name) else os.path.join(RES, name)
def open_resource(name, mode="r"):
    return open(find_resource(name), mode)
''',
    "pisa/core/binning.py": "class OneDimBinning: pass\nclass MultiDimBinning: pass\n",
    "pisa/stages/flux/__init__.py": "",
    "pisa/core/param.py": "class Param:\n    def __init__(self, **k): pass\nclass ParamSet(list):\n    pass\n",
    "pisa/core/stage.py": "class Stage:\n    def __init__(self, **k): pass\n",
    "fast_histogram/__init__.py": "",
}

_LINKS = [
    "pisa/utils/numba_tools.py",
    "pisa/stages/osc/prob3numba/numba_osc_kernels.py",
    "pisa/stages/osc/prob3numba/numba_osc_hostfuncs.py",
    "pisa/stages/osc/layers.py",
    "pisa/stages/osc/osc_params.py",
    "pisa/stages/osc/nsi_params.py",
    "pisa/core/translation.py",
    "pisa/core/bin_indexing.py",
    "pisa/utils/barr_parameterization.py",
    "pisa/stages/flux/barr_simple.py",
    "pisa/utils/flux_weights.py",
]

_loaded = None


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pisa"))


def load():
    """Return a namespace with the reference modules (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    root = tempfile.mkdtemp(prefix="pisa_ref_stub_")
    for rel, txt in _STUBS.items():
        p = os.path.join(root, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(textwrap.dedent(txt))
    with open(os.path.join(root, "pisa/__init__.py"), "w") as f:
        f.write(textwrap.dedent(_STUB_INIT))
    for rel in _LINKS:
        dst = os.path.join(root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        os.symlink(os.path.join(REFERENCE_ROOT, rel), dst)
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "pisa_ref_numba_cache"))
    sys.path.insert(0, root)

    class NS:
        pass

    ns = NS()
    ns.root = root
    ns.pisa = importlib.import_module("pisa")
    ns.kernels = importlib.import_module("pisa.stages.osc.prob3numba.numba_osc_kernels")
    ns.hostfuncs = importlib.import_module("pisa.stages.osc.prob3numba.numba_osc_hostfuncs")
    ns.layers = importlib.import_module("pisa.stages.osc.layers")
    ns.osc_params = importlib.import_module("pisa.stages.osc.osc_params")
    ns.nsi_params = importlib.import_module("pisa.stages.osc.nsi_params")
    ns.translation = importlib.import_module("pisa.core.translation")
    ns.bin_indexing = importlib.import_module("pisa.core.bin_indexing")
    ns.barr_simple = importlib.import_module("pisa.stages.flux.barr_simple")
    ns.flux_weights = importlib.import_module("pisa.utils.flux_weights")
    ns.resources = os.path.join(REFERENCE_ROOT, "pisa_examples", "resources")
    _loaded = ns
    return ns
