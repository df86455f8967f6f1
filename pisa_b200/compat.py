"""Make ``import pisa...`` resolve to this package (opt-in).

The reference's user code and pipeline configs name modules below ``pisa`` (``from pisa.core.pipeline import
Pipeline``, ``pisa.stages.osc.prob3``; pipeline.py:284-296).  After

    import pisa_b200.compat
    pisa_b200.compat.install_as_pisa()

every ``pisa.<x>`` import is served by ``pisa_b200.<x>`` (the same module objects, no copies), so scripts written
against the reference run on the CUDA path without edits -- for the modules this package has; anything else raises
``ModuleNotFoundError`` naming the missing ``pisa_b200`` module instead of falling back to another implementation.
Nothing is aliased unless the function is called, and it refuses to shadow a real PISA that is already imported.
"""
import importlib
import importlib.abc
import importlib.util
import sys

__all__ = ["install_as_pisa", "uninstall"]

_PREFIX = "pisa"
_TARGET = "pisa_b200"


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, target):
        self.target = target

    def create_module(self, spec):
        return importlib.import_module(self.target)      # hand out the pisa_b200 module object itself

    def exec_module(self, module):                        # already executed under its own name
        pass


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != _PREFIX and not fullname.startswith(_PREFIX + "."):
            return None
        real = _TARGET + fullname[len(_PREFIX):]
        try:
            real_spec = importlib.util.find_spec(real)
        except ModuleNotFoundError:
            real_spec = None
        if real_spec is None:
            raise ModuleNotFoundError("No module named %r (pisa_b200 has no %r; only the oscillation-reweighting "
                                      "path of PISA is implemented)" % (fullname, real), name=fullname)
        spec = importlib.util.spec_from_loader(fullname, _AliasLoader(real),
                                               is_package=real_spec.submodule_search_locations is not None)
        return spec


_finder = None


def install_as_pisa():
    """Idempotent.  Raises RuntimeError if a different ``pisa`` package is already imported."""
    global _finder
    existing = sys.modules.get(_PREFIX)
    if existing is not None and getattr(existing, "__name__", None) != _TARGET:
        raise RuntimeError("a different `pisa` package is already imported from %s"
                           % getattr(existing, "__file__", "?"))
    if _finder is None:
        _finder = _AliasFinder()
        sys.meta_path.insert(0, _finder)


def uninstall():
    global _finder
    if _finder is not None:
        sys.meta_path.remove(_finder)
        _finder = None
    for name in [m for m in sys.modules if m == _PREFIX or m.startswith(_PREFIX + ".")]:
        if getattr(sys.modules[name], "__name__", "").startswith(_TARGET):
            del sys.modules[name]
