"""Materialise and load the UNMODIFIED reference hot-path modules (numba CPU path).

Two users:
  * ``tests/golden/ref_loader.py`` (build container only): symlinks the reference files from
    ``/root/reference`` into a temp dir to generate / cross-check the golden fixtures;
  * ``bench.py --impl reference`` and its ``cpu_baseline_numba`` leg (GPU box): loads ``baseline/_ref``, a
    git-ignored directory that ``__graft_entry__.build()`` fills HERE with byte-identical copies of the handful of
    reference files below (it travels to the GPU box with the snapshot; ``/root/reference`` does not exist there).

``import pisa`` of the full package fails in this image (pint / uncertainties / fast_histogram / h5py are not
installed and cannot be fetched), but the hot-path arithmetic needs only a few names from the package root, so a
small stub ``pisa/__init__.py`` (FTYPE / TARGET from ``PISA_FTYPE`` / ``PISA_TARGET`` exactly as the reference reads
them, pisa/__init__.py:152-215) plus stubs of ``log`` / ``comparisons`` / ``fileio`` stand in; every file that
carries arithmetic is the reference's own, unmodified.  Nothing on the product path imports this module.
"""
import importlib
import os
import shutil
import sys
import textwrap

REFERENCE_ROOT = os.environ.get("PISA_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

STUB_INIT = '''
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

STUBS = {
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

# every reference file the fixtures need (tests/golden/ref_loader.py)
ALL_FILES = [
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
# the subset the timed CPU baseline needs: the numba kernels, their host wrappers and the Earth layers
BENCH_FILES = ALL_FILES[:4]


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pisa"))


def materialize(dst, files=ALL_FILES, mode="copy"):
    """Write the stub package into ``dst`` and add the reference files (``copy`` or ``symlink``)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for rel, txt in STUBS.items():
        p = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(textwrap.dedent(txt))
    with open(os.path.join(dst, "pisa/__init__.py"), "w") as f:
        f.write(textwrap.dedent(STUB_INIT))
    for rel in files:
        src, out = os.path.join(REFERENCE_ROOT, rel), os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        if os.path.lexists(out):
            os.remove(out)
        if mode == "symlink":
            os.symlink(src, out)
        else:
            shutil.copyfile(src, out)
    return dst


def build_ref(force=False):
    """``baseline/_ref`` (git-ignored): the stub package + copies of BENCH_FILES.  Called by
    ``__graft_entry__.build()`` in the build container; a no-op where the reference tree is absent."""
    marker = os.path.join(REF_DIR, "pisa", "stages", "osc", "prob3numba", "numba_osc_kernels.py")
    if not reference_available():
        return REF_DIR if os.path.exists(marker) else None
    if force or not os.path.exists(marker):
        materialize(REF_DIR, BENCH_FILES, mode="copy")
    return REF_DIR


def ref_built():
    return os.path.exists(os.path.join(REF_DIR, "pisa", "stages", "osc", "prob3numba", "numba_osc_kernels.py"))


def load_modules(root, names):
    """Import ``names`` (dotted module paths) from the stub package at ``root``; returns a dict."""
    if "pisa" in sys.modules and not getattr(sys.modules["pisa"], "__file__", "").startswith(root):
        raise RuntimeError("another `pisa` package is already imported in this process")
    if root not in sys.path:
        sys.path.insert(0, root)
    return {n: importlib.import_module(n) for n in names}
