"""A/B of template-kernel builds (scratch/variants/libpisa_*.so, prebuilt in the build container, plus the in-tree
library as 'base'): one FP64 template over 12 containers (standard matter and standard NSI), CUDA events, best of 7,
with a checksum of the histograms (bit-identical results expected for schedule-only changes)."""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import _lib, build as B
BASE = B.LIB

def run(path, n_total=48_000_000, dtype=np.float64):
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
    binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
    eng = ReweightEngine(earth, syn.DRAGON_NBINS, dtype, dev)
    for c, (name, nubar, flav) in enumerate(syn.CONTAINERS):
        ev = syn.make_events_torch(n_total // 12, seed=c + 1, dtype=dtype, device=dev)
        index = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
        eng.add_container(name, nubar, flav, true_energy=ev["true_energy"], true_coszen=ev["true_coszen"],
                          nu_flux=ev["nu_flux"], weights=ev["weights"], index=index)
    res = []
    for nsi in (None, syn.STD_NSI):
        dm, mix, mp = syn.osc_matrices(nsi=nsi)
        c = ops.OscConsts.from_matrices(dm, mix, mp)
        for _ in range(3):
            eng.evaluate(c)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); out = eng.evaluate(c); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        res.append((best, float(out.double().sum()), float(out[:, 0].double().abs().sum())))
    del eng
    torch.cuda.empty_cache()
    return res

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48_000_000
    dtype = np.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else np.float64
    libs = [("base", BASE)] + [(os.path.basename(p)[8:-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "scratch", "variants", "libpisa_*.so")))]
    for rep in range(2):
        for tag, path in libs:
            try:
                (t0, c0, _), (t1, c1, _) = run(path, n, dtype)
                print("%-12s std %8.3f ms %.3e ev/s chk %.15e | nsi %8.3f ms %.3e ev/s chk %.15e" % (
                    tag, t0, n / t0 * 1e3, c0, t1, n / t1 * 1e3, c1), flush=True)
            except Exception as e:
                print(tag, "FAILED", repr(e)[:300], flush=True)
