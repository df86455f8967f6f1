#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and numba):

    python tests/golden/make_golden.py            # FP64 fixtures, then re-execs itself for FP32

Outputs (all small, committed):
  ref_pickles_{f8,f4}.npz  the reference's own golden pickles for the hot-path
                           functions (pisa_examples/resources/osc/numba_osc_tests_data),
                           re-packed as arrays: "<func>/<case>/<arg>"
  ref_prob3_{f8,f4}.npz    reference propagate_array on seeded synthetic events
                           (energy, coszen -> probability[N,3,3]) for several
                           parameter sets incl. nubar, IO, deltacp, NSI, NLO
  ref_layers_{f8,f4}.npz   reference Layers/extCalcLayers for PREM 4/12/59
  ref_params_f8.npz        OscParams / StdNSIParams matrices
  ref_hist_f8.npz          find_index / lookup_regular_* / numpy.histogramdd
  ref_flux_{f8,f4}.npz     reference flux.barr_simple (apply_sys_vectorized) for 5 parameter sets
  ref_honda_f8.npz         reference flux_weights.calculate_2d_flux_weights on the Honda 2015 table
The generating inputs are stored alongside the outputs, so tests never need the
reference at run time.
"""
import glob
import os
import pickle
import subprocess
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import ref_loader  # noqa: E402

PICKLE_FUNCS = [
    "propagate_scalar",
    "get_transition_matrix_hostfunc",
    "get_transition_matrix_massbasis_hostfunc",
    "get_H_vac_hostfunc",
    "get_H_mat_hostfunc",
    "get_dms_hostfunc",
    "product_hostfunc",
]

# Parameter sets for the synthetic-event fixtures (angles in degrees).
PARAM_SETS = {
    # settings/osc/nufitv20.cfg NH, deltacp as in osc_example.cfg (0 deg)
    "nufit20_nh": dict(t12=33.48, t13=8.5, t23=42.3, dcp=0.0, dm21=7.5e-5, dm31=2.457e-3),
    "nufit20_nh_dcp306": dict(t12=33.48, t13=8.5, t23=42.3, dcp=306.0, dm21=7.5e-5, dm31=2.457e-3),
    "nufit20_ih_dcp254": dict(t12=33.48, t13=8.51, t23=49.5, dcp=254.0, dm21=7.5e-5, dm31=-2.374e-3),
    "nufit20_nh_nlo": dict(t12=33.48, t13=8.5, t23=42.3, dcp=0.0, dm21=7.5e-5, dm31=2.457e-3, nlo=True),
    # standard NSI with the values the reference test intended (numba_osc_tests.py:130-135)
    "nufit20_nh_dcp306_stdnsi": dict(t12=33.48, t13=8.5, t23=42.3, dcp=306.0, dm21=7.5e-5, dm31=2.457e-3,
                                     nsi=dict(eps_emu=(0.07, 340.0), eps_etau=(0.06, 35.0),
                                              eps_mutau=(0.003, 175.0), eps_ee=0.01, eps_mumu=0.0,
                                              eps_tautau=-0.02)),
}


def build_matrices(ns, ps):
    FT = ns.pisa.FTYPE
    op = ns.osc_params.OscParams()
    op.theta12 = np.deg2rad(ps["t12"])
    op.theta13 = np.deg2rad(ps["t13"])
    op.theta23 = np.deg2rad(ps["t23"])
    op.deltacp = np.deg2rad(ps["dcp"])
    op.dm21 = ps["dm21"]
    op.dm31 = ps["dm31"]
    mix = op.mix_matrix_complex
    dm = op.dm_matrix
    # prob3.py:540-559
    std = np.zeros((3, 3), dtype=FT) + 1.j * np.zeros((3, 3), dtype=FT)
    std[0, 0] += 1.020 if ps.get("nlo") else 1.0
    if "nsi" in ps:
        nsi = ns.nsi_params.StdNSIParams()
        n = ps["nsi"]
        nsi.eps_ee = n["eps_ee"]
        nsi.eps_emu = (n["eps_emu"][0], np.deg2rad(n["eps_emu"][1]))
        nsi.eps_etau = (n["eps_etau"][0], np.deg2rad(n["eps_etau"][1]))
        nsi.eps_mumu = n["eps_mumu"]
        nsi.eps_mutau = (n["eps_mutau"][0], np.deg2rad(n["eps_mutau"][1]))
        nsi.eps_tautau = n["eps_tautau"]
        eps = nsi.eps_matrix
        mat_pot = std + eps
    else:
        eps = np.zeros((3, 3), dtype=ns.pisa.CTYPE)
        mat_pot = std
    return dm, mix, mat_pot, eps


def gen_pickles(ns, tag):
    d = os.path.join(ns.resources, "osc", "numba_osc_tests_data")
    out = {}
    for f in sorted(glob.glob(os.path.join(d, "*__%s.pkl" % tag))):
        func, case, _ = os.path.basename(f)[:-4].split("__")
        if func not in PICKLE_FUNCS:
            continue
        with open(f, "rb") as fh:
            o = pickle.load(fh)
        for k, v in o.items():
            out["%s/%s/%s" % (func, case, k)] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "ref_pickles_%s.npz" % tag), **out)
    print("pickles", tag, len(out))


def gen_prob3(ns, tag, n_events=1200):
    FT, CT = ns.pisa.FTYPE, ns.pisa.CTYPE
    out = {}
    rng = np.random.default_rng(0)
    energy = (10 ** rng.uniform(0, 3, n_events)).astype(FT)
    coszen = rng.uniform(-1, 1, n_events).astype(FT)
    # a few hand-picked directions: vertical up/down, horizon, core tangents
    coszen[:8] = np.array([-1.0, 1.0, 0.0, -0.5, -0.8376, -0.9815, 1e-3, -1e-3], dtype=FT)
    L = ns.layers.Layers(os.path.join(ns.resources, "osc", "PREM_12layer.dat"), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    L.calcLayers(coszen)
    rho = L.density.reshape(n_events, L.max_layers)
    dist = L.distance.reshape(n_events, L.max_layers)
    out["energy"], out["coszen"] = energy, coszen
    out["earth"] = np.array([2.0, 20.0, 0.4656, 0.4656, 0.4957])
    out["densities"], out["distances"] = rho[:64].copy(), dist[:64].copy()  # spot-check rows
    zero_c = np.zeros((3, 3), dtype=CT)
    zero_f = np.zeros((3, 3), dtype=FT)
    lri = np.diag([1e-14, -1e-14, 0.0]).astype(FT)  # as numba_osc_tests.py:100-103
    for name, ps in PARAM_SETS.items():
        dm, mix, mat_pot, eps = build_matrices(ns, ps)
        for nubar in (1, -1):
            for lri_name, lri_pot in (("", zero_f), ("_lri", lri)):
                if lri_name and (nubar == -1 and name != "nufit20_nh"):
                    continue
                if lri_name and name not in ("nufit20_nh", "nufit20_nh_dcp306_stdnsi"):
                    continue
                prob = np.empty((n_events, 3, 3), dtype=FT)
                ns.hostfuncs.propagate_array(dm, mix, mat_pot, -1, zero_c, lri_pot, nubar, energy,
                                             rho, dist, out=prob)
                key = "%s/%s%s" % (name, "nu" if nubar > 0 else "nubar", lri_name)
                out[key + "/dm"], out[key + "/mix"], out[key + "/mat_pot"] = dm, mix, mat_pot
                out[key + "/lri_pot"] = lri_pot
                out[key + "/nubar"] = np.int64(nubar)
                out[key + "/probability"] = prob
                print("prob3", tag, key, float(np.abs(prob.sum(axis=2) - 1).max()))
    np.savez_compressed(os.path.join(HERE, "ref_prob3_%s.npz" % tag), **out)


def gen_layers(ns, tag):
    FT = ns.pisa.FTYPE
    out = {}
    rng = np.random.default_rng(1)
    cases = [
        ("PREM_4layer", 1.0, 20.0, (0.4656, 0.4656, 0.4957)),
        ("PREM_4layer", 10.0, 18.0, (0.5, 0.5, 0.5)),
        ("PREM_12layer", 2.0, 20.0, (0.4656, 0.4656, 0.4957)),
        ("PREM_12layer", 1.0, 2.0, (0.4656, 0.4656, 0.4957)),
        ("PREM_59layer", 2.0, 20.0, (0.4656, 0.4656, 0.4957)),
        ("PREM_10layer", 2.0, 20.0, (0.45, 0.47, 0.5)),
    ]
    for model, depth, height, ye in cases:
        L = ns.layers.Layers(os.path.join(ns.resources, "osc", model + ".dat"), depth, height)
        L.setElecFrac(*ye)
        cz = np.concatenate([
            np.linspace(-1, 1, 101),
            rng.uniform(-1, 1, 150),
            # just above / below every tangent direction
            np.nextafter(L.coszen_limit.astype(np.float64), 2.0),
            np.nextafter(L.coszen_limit.astype(np.float64), -2.0),
        ])
        cz = np.clip(cz, -1, 1).astype(FT)
        # exact tangents (cz == limit) produce a zero-length innermost segment pair; the
        # reference handles them, keep a few
        cz = np.concatenate([cz, L.coszen_limit[L.coszen_limit > -1].astype(FT)])
        try:
            L.calcLayers(cz)
        except Exception as e:  # geometry the reference itself cannot do
            print("layers", model, depth, height, "reference raised", type(e).__name__)
            continue
        key = "%s/d%g_h%g" % (model, depth, height)
        out[key + "/cz"] = cz
        out[key + "/params"] = np.array([depth, height, *ye])
        out[key + "/radii"] = L.radii
        out[key + "/rhos"] = L.rhos
        out[key + "/coszen_limit"] = L.coszen_limit
        out[key + "/r_detector"] = np.float64(L.r_detector)
        out[key + "/max_layers"] = np.int64(L.max_layers)
        out[key + "/n_layers"] = L.n_layers
        out[key + "/density"] = L.density.reshape(len(cz), L.max_layers)
        out[key + "/distance"] = L.distance.reshape(len(cz), L.max_layers)
        print("layers", tag, key, L.max_layers, int(L.n_layers.max()))
    np.savez_compressed(os.path.join(HERE, "ref_layers_%s.npz" % tag), **out)


def gen_params(ns):
    out = {}
    for name, ps in PARAM_SETS.items():
        dm, mix, mat_pot, eps = build_matrices(ns, ps)
        out[name + "/dm"], out[name + "/mix"], out[name + "/mat_pot"], out[name + "/eps"] = dm, mix, mat_pot, eps
        op = ns.osc_params.OscParams()
        op.theta12, op.theta13, op.theta23 = (np.deg2rad(ps[k]) for k in ("t12", "t13", "t23"))
        op.deltacp = np.deg2rad(ps["dcp"])
        op.dm21, op.dm31 = ps["dm21"], ps["dm31"]
        out[name + "/mix_reparam"] = op.mix_matrix_reparam_complex
    # degenerate splittings (osc_params.py:270-280)
    op = ns.osc_params.OscParams()
    op.dm21, op.dm31 = 0.0, 0.0
    out["degenerate/dm"] = op.dm_matrix
    # vacuum-like NSI parameterisation (nsi_params.py:184-384): seeded parameter sets + the reference's default
    rng = np.random.default_rng(7)
    sets = [dict(eps_scale=1.0, eps_prime=0.0, phi12=0.0, phi13=0.0, phi23=0.0, alpha1=0.0, alpha2=0.0, deltansi=0.0)]
    for _ in range(6):
        sets.append(dict(eps_scale=rng.uniform(0.5, 1.5), eps_prime=rng.uniform(-0.3, 0.3),
                         phi12=rng.uniform(-np.pi, np.pi), phi13=rng.uniform(-np.pi, np.pi),
                         phi23=rng.uniform(-np.pi, np.pi), alpha1=rng.uniform(0, 2 * np.pi),
                         alpha2=rng.uniform(0, 2 * np.pi), deltansi=rng.uniform(0, 2 * np.pi)))
    names = sorted(sets[0])
    out["vacuum_nsi/names"] = np.array(names)
    out["vacuum_nsi/values"] = np.array([[s[k] for k in names] for s in sets])
    mats = []
    for s in sets:
        v = ns.nsi_params.VacuumLikeNSIParams()
        for k in names:
            setattr(v, k, s[k])
        mats.append(v.eps_matrix)
    out["vacuum_nsi/eps"] = np.array(mats)
    np.savez_compressed(os.path.join(HERE, "ref_params_f8.npz"), **out)


def gen_hist(ns):
    tr = ns.translation
    out = {}
    rng = np.random.default_rng(0)
    # find_index edge cases in the spirit of translation.py:821-940
    for name, edges in (("lin", np.linspace(-1, 1, 9)), ("log", np.logspace(0, 2, 11)),
                        ("irr", np.array([5.62341325, 7.49894209, 10.0, 13.33521432, 17.7827941,
                                          23.71373706, 31.6227766, 42.16965034, 56.23413252])),
                        ("one", np.array([0.0, 1.0])), ("inf", np.array([-np.inf, 0.55, np.inf]))):
        edges = edges.astype(np.float64)
        fin = edges[np.isfinite(edges)]
        vals = np.concatenate([edges, np.nextafter(fin, np.inf), np.nextafter(fin, -np.inf),
                               [-np.inf, np.inf, np.nan, fin.min() - 1, fin.max() + 1],
                               rng.uniform(fin.min() - 0.5, fin.max() + 0.5, 200)])
        idx = np.array([tr.find_index(v, edges) for v in vals], dtype=np.int64)
        out["find_index/%s/edges" % name], out["find_index/%s/vals" % name] = edges, vals
        out["find_index/%s/idx" % name] = idx
    # regular lookups with the reference's own njit functions
    n = 5000
    x = rng.uniform(-0.2, 3.2, n)
    y = rng.uniform(-1.3, 1.3, n)
    z = rng.uniform(-0.5, 2.5, n)
    # exact edge values and +-1ulp
    x[:6] = [0.0, 3.0, np.nextafter(3.0, 0), np.nextafter(0.0, -1), 1.5, 0.75]
    y[:6] = [-1.0, 1.0, np.nextafter(1.0, 0), np.nextafter(-1.0, -2), 0.0, 0.25]
    nx, ny, nz = 12, 8, 2
    h1 = rng.normal(size=nx)
    h2 = rng.normal(size=nx * ny)
    h3 = rng.normal(size=nx * ny * nz)
    h2a = rng.normal(size=(nx * ny, 3))
    o1, o2, o3 = np.zeros(n), np.zeros(n), np.zeros(n)
    o2a = np.zeros((n, 3))
    tr.lookup_regular_1d(x, h1, 0.0, 3.0, nx, o1)
    tr.lookup_regular_2d(x, y, h2, 0.0, 3.0, nx, -1.0, 1.0, ny, o2)
    tr.lookup_regular_3d(x, y, z, h3, 0.0, 3.0, nx, -1.0, 1.0, ny, 0.0, 2.0, nz, o3)
    tr.lookup_regular_2d_array(x, y, h2a, 0.0, 3.0, nx, -1.0, 1.0, ny, o2a)
    for k, v in dict(x=x, y=y, z=z, h1=h1, h2=h2, h3=h3, h2a=h2a, o1=o1, o2=o2, o3=o3, o2a=o2a).items():
        out["lookup/" + k] = v
    out["lookup/binning"] = np.array([0.0, 3.0, nx, -1.0, 1.0, ny, 0.0, 2.0, nz])
    # numpy.histogramdd (the reference's histogram_np branch and the pin of test_histogram,
    # translation.py:779-818): 10000 uniform samples, seed 0, weighted
    rs = np.random.RandomState(0)
    s = rs.rand(10000, 3)
    w = rs.rand(10000)
    for d in (1, 2, 3):
        edges = [np.linspace(0, 1, nb + 1) for nb in (10, 7, 5)[:d]]
        hw, _ = np.histogramdd(s[:, :d], bins=edges, weights=w)
        hc, _ = np.histogramdd(s[:, :d], bins=edges)
        out["histdd/%dd/weighted" % d], out["histdd/%dd/counts" % d] = hw, hc
    out["histdd/sample"], out["histdd/weights"] = s, w
    np.savez_compressed(os.path.join(HERE, "ref_hist_f8.npz"), **out)


FLUX_CASES = {
    # name: (nue_numu_ratio, nu_nubar_ratio, delta_index, Barr_uphor_ratio, Barr_nu_nubar_ratio)
    "nominal": (1.0, 1.0, 0.0, 0.0, 0.0),
    "all_up": (1.03, 0.97, 0.05, 0.3, -0.2),
    "all_down": (0.95, 1.1, -0.1, -1.0, 1.0),
    "uphor_only": (1.0, 1.0, 0.0, 2.0, 0.0),
    "nubar_only": (1.0, 1.0, 0.0, 0.0, -3.0),
}


def gen_flux(ns, tag, n_events=1500):
    """Reference flux.barr_simple (apply_sys_vectorized, barr_simple.py:200-226) on seeded events."""
    ft = ns.pisa.FTYPE
    rng = np.random.default_rng(42)
    e = (10 ** rng.uniform(0, 3, n_events)).astype(ft)
    cz = rng.uniform(-1, 1, n_events).astype(ft)
    nu = rng.uniform(0.5, 1.5, (n_events, 2)).astype(ft)
    nb = rng.uniform(0.5, 1.5, (n_events, 2)).astype(ft)
    nu[:3] = 0.0   # the in1 == in2 == 0 branch of apply_ratio_scale
    nb[:2] = 0.0
    out = {"true_energy": e, "true_coszen": cz, "nu_flux_nominal": nu, "nubar_flux_nominal": nb}
    for name, pars in FLUX_CASES.items():
        out[name + "/params"] = np.array(pars, dtype=np.float64)
        for nubar in (1, -1):
            res = np.empty((n_events, 2), dtype=ft)
            ns.barr_simple.apply_sys_vectorized(e, cz, nu, nb, nubar, *[ft(p) for p in pars], out=res)
            out["%s/%s" % (name, "nu" if nubar > 0 else "nubar")] = res
    np.savez_compressed(os.path.join(HERE, "ref_flux_%s.npz" % tag), **out)


def gen_honda(ns, n_events=600):
    """Reference honda_ip arithmetic (flux_weights.load_2d_table + calculate_2d_flux_weights) on seeded events,
    incl. energies outside the table (extrapolation) and the coszen end points."""
    fw = ns.flux_weights
    table = "flux/honda-2015-spl-solmin-aa.d"
    splines = fw.load_2d_table(table)
    rng = np.random.default_rng(7)
    e = 10 ** rng.uniform(-1.2, 4.3, n_events)
    cz = rng.uniform(-1, 1, n_events)
    cz[:6] = [-1.0, 1.0, -0.8, 0.8, 0.0, -0.95]
    out = {"true_energy": e, "true_coszen": cz, "table": np.array(table)}
    for prim in ("nue", "numu", "nuebar", "numubar"):
        out[prim] = fw.calculate_2d_flux_weights(e, cz, splines[prim])
        # spline coefficients of two rows, to pin the product's own table construction bit for bit
        for key in ("-0.95", "0.45"):
            t, c, k = splines[prim][key]
            out["tck/%s/%s/t" % (prim, key)], out["tck/%s/%s/c" % (prim, key)] = np.asarray(t), np.asarray(c)
    np.savez_compressed(os.path.join(HERE, "ref_honda_f8.npz"), **out)


def main():
    ns = ref_loader.load()
    tag = "f4" if ns.pisa.FTYPE == np.float32 else "f8"
    gen_pickles(ns, tag)
    gen_layers(ns, tag)
    gen_prob3(ns, tag)
    gen_flux(ns, tag)
    if tag == "f8":
        gen_params(ns)
        gen_hist(ns)
        gen_honda(ns)
        env = dict(os.environ, PISA_FTYPE="fp32")
        subprocess.check_call([sys.executable, os.path.abspath(__file__)], env=env)


if __name__ == "__main__":
    main()
