#!/usr/bin/env python
"""Time the UNMODIFIED reference numba CPU path on this machine's host cores (CPU baseline of bench.py).

Runs in its own process (the reference fixes FTYPE / TARGET at import time: ``PISA_FTYPE``, ``PISA_TARGET``,
pisa/__init__.py:152-215) and prints ONE JSON line.  The stage sequence is the reference's own:

  setup (untimed, like prob3.setup_function, prob3.py:406-409):
      Layers.calcLayers(true_coszen) -> densities, distances [N, max_layers]
  step (timed; prob3.compute_function :581-605, apply_function :621-622, hist.apply_function hist.py:198-209):
      propagate_array(dm, mix, mat_pot, decay_flag, mat_decay, lri_pot, nubar, energy, densities, distances)
      fill_probs(probability, 0 | 1, flav)                       (numba_osc_hostfuncs.py:206-221)
      weights = w0 * (flux_e * prob_e + flux_mu * prob_mu)
      histogram of (reco_energy, reco_coszen, pid) with weights and weights**2 (sumw2) -- through
      ``numpy.histogramdd``, the reference's own ``histogram_np`` branch (translation.py:207-223), because
      ``fast_histogram`` is not installed in this image.

The modules come from ``baseline/_ref`` (byte-identical copies of the reference files behind a stub package root,
``baseline/ref_pkg.py``), or straight from ``/root/reference`` when that tree is present.
"""
import argparse
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ftype", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--target", default="parallel", choices=["cpu", "parallel"])
    ap.add_argument("--events", type=int, default=1200000)
    ap.add_argument("--seconds-per-step", type=float, default=0.0,
                    help="size the sample for about this much CPU time per step (pilot run after the JIT); overrides --events")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--nsi", action="store_true")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--dump", default=None, help="write the histograms of the last step to this .npy (parity check)")
    args = ap.parse_args()

    os.environ["PISA_FTYPE"] = args.ftype
    os.environ["PISA_TARGET"] = args.target
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "pisa_ref_numba_cache"))
    import numpy as np
    from baseline import ref_pkg
    if ref_pkg.ref_built():
        root, origin = ref_pkg.REF_DIR, "baseline/_ref"
    elif ref_pkg.reference_available():
        root = ref_pkg.materialize(tempfile.mkdtemp(prefix="pisa_ref_stub_"), ref_pkg.BENCH_FILES, mode="symlink")
        origin = ref_pkg.REFERENCE_ROOT
    else:
        print(json.dumps({"unavailable": "baseline/_ref not built and no reference tree"}))
        return 0
    t_imp = time.perf_counter()
    mods = ref_pkg.load_modules(root, ["pisa", "pisa.stages.osc.prob3numba.numba_osc_hostfuncs", "pisa.stages.osc.layers"])
    import numba
    pisa = mods["pisa"]
    host = mods["pisa.stages.osc.prob3numba.numba_osc_hostfuncs"]
    Layers = mods["pisa.stages.osc.layers"].Layers
    FT, CT, IT = pisa.FTYPE, pisa.CTYPE, pisa.ITYPE
    t_imp = time.perf_counter() - t_imp

    # the product's own host-side parameter code builds the (pinned, bit-identical) matrices and the events
    from pisa_b200.utils import synthetic as syn
    dm, mix, mat_pot = syn.osc_matrices(nsi=syn.STD_NSI if args.nsi else None)
    dm, mix, mat_pot = dm.astype(FT), mix.astype(CT), mat_pot.astype(CT)
    mat_decay = np.zeros((3, 3), dtype=CT)
    lri_pot = np.zeros((3, 3), dtype=FT)
    prem = os.path.join(ROOT, "pisa_b200", "resources", syn.EARTH["earth_model"])
    L = Layers(prem, syn.EARTH["detector_depth"], syn.EARTH["prop_height"])
    L.setElecFrac(syn.EARTH["YeI"], syn.EARTH["YeO"], syn.EARTH["YeM"])

    edges = [syn.DRAGON_E_EDGES, np.linspace(-1.0, 1.0, 9), np.linspace(-0.5, 1.5, 3)]
    blocks = []

    def make(n):
        n = n // 12 * 12
        ev = syn.make_events_numpy(n, args.seed, dtype=FT)
        per = n // 12
        del blocks[:]
        for c, (name, nubar, flav) in enumerate(syn.CONTAINERS):
            sl = slice(c * per, (c + 1) * per)
            b = {k: np.ascontiguousarray(v[sl]) for k, v in ev.items()}
            L.calcLayers(b["true_coszen"])                                 # setup_function
            b["densities"] = L.density.reshape((per, L.max_layers)).astype(FT)
            b["distances"] = L.distance.reshape((per, L.max_layers)).astype(FT)
            b["probability"] = np.empty((per, 3, 3), dtype=FT)
            b["prob_e"] = np.empty(per, dtype=FT)
            b["prob_mu"] = np.empty(per, dtype=FT)
            blocks.append((nubar, flav, b))
        return n

    def step():
        out = np.zeros((12, 2, 128))
        for c, (nubar, flav, b) in enumerate(blocks):
            host.propagate_array(dm, mix, mat_pot, IT(-1), mat_decay, lri_pot, IT(nubar), b["true_energy"],
                                 b["densities"], b["distances"], out=b["probability"])
            host.fill_probs(b["probability"], 0, flav, out=b["prob_e"])
            host.fill_probs(b["probability"], 1, flav, out=b["prob_mu"])
            w = b["weights"] * ((b["nu_flux"][:, 0] * b["prob_e"]) + (b["nu_flux"][:, 1] * b["prob_mu"]))
            sample = [b["reco_energy"], b["reco_coszen"], b["pid"]]
            out[c, 0] = np.histogramdd(sample, bins=edges, weights=w)[0].ravel()
            out[c, 1] = np.histogramdd(sample, bins=edges, weights=w * w)[0].ravel()
        return out

    n = make(48000 if args.seconds_per_step > 0 else args.events)
    t0 = time.perf_counter()
    step()                                                             # first call: includes JIT compilation
    t_first = time.perf_counter() - t0
    if args.seconds_per_step > 0:
        t0 = time.perf_counter()
        step()
        rate = n / (time.perf_counter() - t0)
        n = make(int(min(max(rate * args.seconds_per_step, 48000), 2e7)))
        step()
    for _ in range(max(0, args.warmup - 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step()
    dt = time.perf_counter() - t0
    if args.dump:
        np.save(args.dump, out)
    threads = numba.get_num_threads() if args.target == "parallel" else 1
    print(json.dumps({
        "value": n * args.steps / dt, "unit": "events/s", "events": n, "steps": args.steps, "seconds": dt,
        "first_call_seconds": t_first, "import_seconds": t_imp, "cores": threads, "host_cpus": os.cpu_count(),
        "target": args.target, "ftype": args.ftype, "nsi": bool(args.nsi), "origin": origin,
        "numba": numba.__version__, "threading_layer": numba.threading_layer() if args.target == "parallel" else None,
        "hist_total": float(out[:, 0].sum()),
        "what": "reference numba propagate_array + fill_probs + reweight + numpy.histogramdd (w, w^2); "
                "layers in setup like prob3.setup_function"}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
