"""Decay template kernel variants (prebuilt in scratch/variants/ by the build container, -D flags in the file names):
time of one template over 12 containers, FP64 storage, CUDA events, best of 5; checksum against the first variant."""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import _lib, build as B

def run(path, sizes=(12_000_000,)):
    _lib._lib = None
    B.LIB = path
    _lib._build.LIB = path
    from pisa_b200 import ops
    from pisa_b200.engine import ReweightEngine
    from pisa_b200.stages.osc.layers import Layers
    from pisa_b200.utils import synthetic as syn
    dev = torch.device("cuda:0")
    L = Layers(os.path.join(ROOT, "pisa_b200/resources/osc/PREM_12layer.dat"), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = L.earth_struct()
    dm, mix, mp = syn.osc_matrices()
    md = np.zeros((3, 3), complex); md[2, 2] = -1e-4j
    dec = ops.OscConsts.from_matrices(dm, mix, mp, 1, md)
    binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
    res = []
    for n_total in sizes:
        eng = ReweightEngine(earth, syn.DRAGON_NBINS, np.float64, dev)
        for c, (name, nubar, flav) in enumerate(syn.CONTAINERS):
            ev = syn.make_events_torch(n_total // 12, seed=c + 1, dtype=np.float64, device=dev)
            index = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
            eng.add_container(name, nubar, flav, true_energy=ev["true_energy"], true_coszen=ev["true_coszen"],
                              nu_flux=ev["nu_flux"], weights=ev["weights"], index=index)
        for _ in range(2):
            eng.evaluate(dec)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = eng.evaluate(dec); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        res.append((n_total, best, float(out[:, 0].sum())))
        del eng
    ev = syn.make_events_torch(4_000_000, seed=99, dtype=np.float64, device=dev)
    order = ops.layer_order(earth, ev["true_coszen"])
    for want, od in ((True, None), (True, order), (False, None), (False, order)):
        f = lambda: ops.propagate_earth(dec, earth, 1, ev["true_energy"], ev["true_coszen"], flav=1, want_probability=want, order=od)
        f(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); o = f(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        res.append((4_000_000, best, float((o[0] if want else o[1]).sum())))
    return res

if __name__ == "__main__":
    for path in sorted(glob.glob(os.path.join(ROOT, "scratch", "variants", "libpisa_*.so"))):
        try:
            r = run(path)
            print("%-14s " % os.path.basename(path)[8:-3] + " | ".join("%9d ev: %8.3f ms %.3e ev/s chk %.10e" % (n, ms, n / ms * 1e3, chk) for n, ms, chk in r), flush=True)
        except Exception as e:
            print(path, "FAILED", repr(e)[:300], flush=True)
