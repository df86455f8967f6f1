"""Load the *unmodified* reference hot-path modules from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by ``make_golden.py`` (fixture generation) and
by the optional ``-m "not gpu"`` cross-checks that run in the build container.
Nothing on the product path or in ``smoke()`` imports this: ``/root/reference``
does not exist on the GPU box.

The stub-package recipe (which reference files, which names are stubbed) lives in
``baseline/ref_pkg.py`` and is shared with the timed CPU baseline of ``bench.py``;
here the reference files are symlinked into a temp dir (never copied into this repo).

The float type is fixed at import time by the reference (``PISA_FTYPE``), so
FP32 fixtures are produced in a separate process.
"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from baseline import ref_pkg  # noqa: E402

REFERENCE_ROOT = ref_pkg.REFERENCE_ROOT
_LINKS = ref_pkg.ALL_FILES
_loaded = None


def available():
    return ref_pkg.reference_available()


def load():
    """Return a namespace with the reference modules (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    root = tempfile.mkdtemp(prefix="pisa_ref_stub_")
    ref_pkg.materialize(root, ref_pkg.ALL_FILES, mode="symlink")
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "pisa_ref_numba_cache"))
    mods = ref_pkg.load_modules(root, [
        "pisa", "pisa.stages.osc.prob3numba.numba_osc_kernels", "pisa.stages.osc.prob3numba.numba_osc_hostfuncs",
        "pisa.stages.osc.layers", "pisa.stages.osc.osc_params", "pisa.stages.osc.nsi_params",
        "pisa.core.translation", "pisa.core.bin_indexing", "pisa.stages.flux.barr_simple", "pisa.utils.flux_weights"])

    class NS:
        pass

    ns = NS()
    ns.root = root
    ns.pisa = mods["pisa"]
    ns.kernels = mods["pisa.stages.osc.prob3numba.numba_osc_kernels"]
    ns.hostfuncs = mods["pisa.stages.osc.prob3numba.numba_osc_hostfuncs"]
    ns.layers = mods["pisa.stages.osc.layers"]
    ns.osc_params = mods["pisa.stages.osc.osc_params"]
    ns.nsi_params = mods["pisa.stages.osc.nsi_params"]
    ns.translation = mods["pisa.core.translation"]
    ns.bin_indexing = mods["pisa.core.bin_indexing"]
    ns.barr_simple = mods["pisa.stages.flux.barr_simple"]
    ns.flux_weights = mods["pisa.utils.flux_weights"]
    ns.resources = os.path.join(REFERENCE_ROOT, "pisa_examples", "resources")
    _loaded = ns
    return ns
