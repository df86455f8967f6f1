#!/usr/bin/env python
"""Generate tests/golden/ref_decay_f8.npz from the UNMODIFIED reference (neutrino decay branch).

Run in the build container only (needs /root/reference and numba):

    python tests/golden/make_golden_decay.py

Contents:
  * the reference's own golden pickles of the decay-only host functions
    (``get_dms_numerical_hostfunc``, ``get_H_decay_hostfunc``; the decay case of the other
    functions is already in ref_pickles_f8.npz), re-packed as "<func>/<case>/<arg>";
  * reference ``propagate_array`` with ``decay_flag = 1`` (numba_osc_kernels.py:445-451 ->
    ``get_dms_numerical`` -> ``numpy.linalg.eigvals``) on seeded synthetic events through
    PREM_12layer, for several parameter sets (nu / nubar, deltacp, a large alpha3, NSI + LRI,
    a general complex decay matrix); the generating inputs are stored alongside the outputs.
"""
import glob
import os
import pickle
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import ref_loader  # noqa: E402
from make_golden import PARAM_SETS, build_matrices  # noqa: E402

# (parameter set, alpha3 [eV^2]); decay_params.py:47-55: decay_matrix = diag(0, 0, -i alpha3)
DECAY_SETS = {
    "nufit20_nh_a1e-4": ("nufit20_nh", 1e-4),
    "nufit20_nh_dcp306_a5e-4": ("nufit20_nh_dcp306", 5e-4),
    "nufit20_ih_dcp254_a1e-4": ("nufit20_ih_dcp254", 1e-4),
    "nufit20_nh_dcp306_stdnsi_a2e-4": ("nufit20_nh_dcp306_stdnsi", 2e-4),
    "nufit20_nh_a0": ("nufit20_nh", 0.0),
}


def main(n_events=600):
    ns = ref_loader.load()
    FT, CT = ns.pisa.FTYPE, ns.pisa.CTYPE
    assert FT == np.float64
    out = {}
    d = os.path.join(ns.resources, "osc", "numba_osc_tests_data")
    for func in ("get_dms_numerical_hostfunc", "get_H_decay_hostfunc"):
        for f in sorted(glob.glob(os.path.join(d, "%s__*__f8.pkl" % func))):
            _, case, _ = os.path.basename(f)[:-4].split("__")
            with open(f, "rb") as fh:
                o = pickle.load(fh)
            for k, v in o.items():
                out["%s/%s/%s" % (func, case, k)] = np.asarray(v)

    rng = np.random.default_rng(7)
    energy = (10 ** rng.uniform(0, 3, n_events)).astype(FT)
    coszen = rng.uniform(-1, 1, n_events).astype(FT)
    coszen[:8] = np.array([-1.0, 1.0, 0.0, -0.5, -0.8376, -0.9815, 1e-3, -1e-3], dtype=FT)
    L = ns.layers.Layers(os.path.join(ns.resources, "osc", "PREM_12layer.dat"), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    L.calcLayers(coszen)
    rho = L.density.reshape(n_events, L.max_layers)
    dist = L.distance.reshape(n_events, L.max_layers)
    out["energy"], out["coszen"] = energy, coszen
    out["earth"] = np.array([2.0, 20.0, 0.4656, 0.4656, 0.4957])
    zero_f = np.zeros((3, 3), dtype=FT)
    lri = np.diag([1e-14, -1e-14, 0.0]).astype(FT)
    for name, (ps_name, alpha3) in DECAY_SETS.items():
        dm, mix, mat_pot, _ = build_matrices(ns, PARAM_SETS[ps_name])
        mat_decay = np.zeros((3, 3), dtype=CT)
        mat_decay[2, 2] = 0 - alpha3 * 1j
        lri_pot = lri if "stdnsi" in name else zero_f
        for nubar in (1, -1):
            prob = np.empty((n_events, 3, 3), dtype=FT)
            ns.hostfuncs.propagate_array(dm, mix, mat_pot, 1, mat_decay, lri_pot, nubar, energy, rho, dist, out=prob)
            key = "%s/%s" % (name, "nu" if nubar > 0 else "nubar")
            out[key + "/dm"], out[key + "/mix"], out[key + "/mat_pot"] = dm, mix, mat_pot
            out[key + "/mat_decay"], out[key + "/lri_pot"] = mat_decay, lri_pot
            out[key + "/nubar"] = np.int64(nubar)
            out[key + "/probability"] = prob
            print(key, "min row sum", float(prob.sum(axis=2).min()), "max", float(prob.sum(axis=2).max()))
    # a general complex decay matrix (the kernel argument is a full 3x3: numba_osc_kernels.py:571-603)
    dm, mix, mat_pot, _ = build_matrices(ns, PARAM_SETS["nufit20_nh_dcp306"])
    mat_decay = np.array([[0, 0, 0], [0, -2e-5j, 1e-5 - 1e-5j], [0, 1e-5 + 1e-5j, -1e-4j]], dtype=CT)
    for nubar in (1, -1):
        prob = np.empty((n_events, 3, 3), dtype=FT)
        ns.hostfuncs.propagate_array(dm, mix, mat_pot, 1, mat_decay, zero_f, nubar, energy, rho, dist, out=prob)
        key = "general_matrix/%s" % ("nu" if nubar > 0 else "nubar")
        out[key + "/dm"], out[key + "/mix"], out[key + "/mat_pot"] = dm, mix, mat_pot
        out[key + "/mat_decay"], out[key + "/lri_pot"] = mat_decay, zero_f
        out[key + "/nubar"] = np.int64(nubar)
        out[key + "/probability"] = prob
        print(key, "min row sum", float(prob.sum(axis=2).min()))
    np.savez_compressed(os.path.join(HERE, "ref_decay_f8.npz"), **out)
    print("wrote ref_decay_f8.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
