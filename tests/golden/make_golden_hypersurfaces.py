"""Golden vectors for discr_sys.hypersurfaces from the UNMODIFIED reference code.

TEST INFRASTRUCTURE ONLY (build container; /root/reference does not exist on the GPU box).
    python tests/golden/make_golden_hypersurfaces.py        ->  tests/golden/ref_hypersurfaces_f8.npz

What runs is the reference's own
  * ``pisa/utils/hypersurface/hypersurface.py``: ``_load_hypersurfaces_data_release`` (:2065-2172) on the shipped
    IceCube-3y CSV hyperplanes, ``Hypersurface.evaluate`` (:356-476), ``HypersurfaceParam.evaluate`` (:1447-1480),
    ``linear_hypersurface_func`` (:80-99);
  * ``pisa/stages/discr_sys/hypersurfaces.py``: ``compute_function`` (:160-216) and ``apply_function`` (:219-243),
    called unbound on a stand-in ``self`` whose containers are plain dicts of numpy arrays.
The full ``pisa`` package cannot be imported in this image (pint, iminuit, uncertainties, h5py ... are absent and
cannot be fetched), so the two files are loaded inside the stub package of ``baseline/ref_pkg.py`` extended by the
import-only stubs below: none of them carries arithmetic that the exercised functions use.
"""
import os
import sys
import tempfile
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from baseline import ref_pkg  # noqa: E402

FILES = ["pisa/utils/hypersurface/hypersurface.py", "pisa/stages/discr_sys/hypersurfaces.py"]
STUBS = {
    "iminuit/__init__.py": "class Minuit: pass\n",
    "uncertainties/__init__.py": "def ufloat(*a, **k): raise NotImplementedError\ndef correlated_values(*a, **k): raise NotImplementedError\nfrom . import unumpy\n",
    "uncertainties/unumpy.py": "",
    "pisa/utils/jsons.py": "def from_json(*a, **k): raise NotImplementedError\ndef to_json(*a, **k): raise NotImplementedError\n",
    "pisa/core/pipeline.py": "class Pipeline: pass\n",
    "pisa/core/map.py": "class Map: pass\n",
    "pisa/utils/hypersurface/__init__.py": "from .hypersurface import *\n",
    "pisa/utils/hypersurface/hypersurface_plotting.py": "def plot_bin_fits(*a, **k): pass\ndef plot_bin_fits_2d(*a, **k): pass\n",
    "pisa/stages/discr_sys/__init__.py": "",
    "pisa/utils/fileio.py": "import numpy as np\ndef from_file(fname, as_array=False, **kw):\n    return np.loadtxt(fname)\ndef mkdir(*a, **k): pass\n",
    # the names and sizes of the dimensions are all the exercised code asks of a binning
    "pisa/core/binning.py": '''
        class OneDimBinning:
            def __init__(self, name, num_bins): self.name, self.num_bins = name, num_bins
        class MultiDimBinning:
            def __init__(self, dims): self.dims = list(dims)
            names = property(lambda self: [d.name for d in self.dims])
            shape = property(lambda self: tuple(d.num_bins for d in self.dims))
            def __getitem__(self, name): return {d.name: d for d in self.dims}[name]
        def is_binning(x): return isinstance(x, (OneDimBinning, MultiDimBinning))
        ''',
}


def load_reference():
    root = tempfile.mkdtemp(prefix="pisa_ref_hs_")
    ref_pkg.materialize(root, FILES, mode="symlink")
    for rel, txt in STUBS.items():
        p = os.path.join(root, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(textwrap.dedent(txt))
    mods = ref_pkg.load_modules(root, ["pisa", "pisa.core.binning", "pisa.utils.hypersurface.hypersurface",
                                       "pisa.stages.discr_sys.hypersurfaces"])
    return mods


class _Container(dict):
    """What the stage's compute/apply functions ask of a container: item access, ``keys``, name, size, mark_changed."""

    def __init__(self, name, size, **arrays):
        super().__init__(**arrays)
        self.name, self.size = name, size

    keys = property(lambda self: list(dict.keys(self)))

    def mark_changed(self, key):
        pass


class _Data(list):
    representation = "binned"

    def link_containers(self, *a):
        pass

    def unlink_containers(self):
        pass


class _Quantity:
    def __init__(self, m):
        self.m = m


def main():
    mods = load_reference()
    binning_mod = mods["pisa.core.binning"]
    hs_mod = mods["pisa.utils.hypersurface.hypersurface"]
    stage_mod = mods["pisa.stages.discr_sys.hypersurfaces"]
    binning = binning_mod.MultiDimBinning([binning_mod.OneDimBinning("reco_energy", 8),
                                           binning_mod.OneDimBinning("reco_coszen", 8), binning_mod.OneDimBinning("pid", 2)])
    proto = os.path.join(ref_pkg.REFERENCE_ROOT, "pisa_examples", "resources", "events", "IceCube_3y_oscillations",
                         "hyperplanes_*.csv.bz2")
    surfaces = hs_mod._load_hypersurfaces_data_release(proto, binning)
    names = list(surfaces.values())[0].param_names
    rng = np.random.RandomState(20)
    cases = [dict(zip(names, vals)) for vals in ([0.0] * len(names), [1.0] * len(names),
                                                 *[list(rng.uniform(-2.0, 2.0, len(names))) for _ in range(4)],
                                                 [25.0, -30.0] + [5.0] * (len(names) - 2))]     # drives bins below 0
    out = dict(param_names=np.array(names), map_names=np.array(list(surfaces)),
               param_values=np.array([[c[n] for n in names] for c in cases]))
    size = int(np.prod(binning.shape))
    for ci, case in enumerate(cases):
        # Hypersurface.evaluate directly
        out["scales_%d" % ci] = np.stack([surfaces[m].evaluate(case) for m in surfaces])
        # the stage: compute_function + apply_function on binned containers
        data = _Data()
        for k, m in enumerate(surfaces):
            r = np.random.RandomState(100 * ci + k)
            data.append(_Container(m, size, weights=r.uniform(0.0, 50.0, size), errors=r.uniform(0.0, 3.0, size),
                                   bin_unc2=r.uniform(0.0, 50.0, size), hs_scales=np.empty(size)))
        for c in data:
            for key in ("weights", "errors", "bin_unc2"):
                out["in_%s_%d_%s" % (key, ci, c.name)] = c[key].copy()

        class _Stage:
            pass
        st = _Stage()
        st.links, st.data, st.hypersurfaces, st.hypersurface_param_names = None, data, surfaces, names
        st.params = {n: _Quantity(case[n]) for n in names}
        st.interpolated = st.fluctuate = st.propagate_uncertainty = False
        st.warning_issued = True
        st.error_method = "sumw2"
        stage_mod.hypersurfaces.compute_function(st)
        stage_mod.hypersurfaces.apply_function(st)
        for c in data:
            for key in ("weights", "errors", "bin_unc2", "hs_scales"):
                out["out_%s_%d_%s" % (key, ci, c.name)] = np.asarray(c[key]).copy()
    path = os.path.join(HERE, "ref_hypersurfaces_f8.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "cases", len(cases), "maps", list(surfaces), "params", names)
    print("min scale over cases:", min(float(out["scales_%d" % i].min()) for i in range(len(cases))))


if __name__ == "__main__":
    main()
