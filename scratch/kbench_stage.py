"""Times the Stage-API kernels (propagate_earth FULL / row mode, hist_index, hist_accumulate, lookup)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import ops
from pisa_b200.stages.osc.layers import Layers
from pisa_b200.utils import synthetic as syn

def timeit(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

if __name__ == '__main__':
    dev = torch.device("cuda:0")
    L = Layers(os.path.join(ROOT, "pisa_b200/resources/osc/PREM_12layer.dat"), 2.0, 20.0); L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = L.earth_struct()
    n = 8_333_333
    ev = syn.make_events_torch(n, 3, np.float64, dev)
    binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
    idx = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
    order = ops.layer_order(earth, ev["true_coszen"])
    for nsi in (None, syn.STD_NSI):
        dm, mix, mp = syn.osc_matrices(nsi=nsi); consts = ops.OscConsts.from_matrices(dm, mix, mp)
        prob = torch.empty((n, 3, 3), dtype=torch.float64, device=dev); pe = torch.empty(n, dtype=torch.float64, device=dev); pm = torch.empty_like(pe)
        t = timeit(lambda: ops.propagate_earth(consts, earth, 1, ev["true_energy"], ev["true_coszen"], probability=prob, order=order))
        print("nsi=%s propagate_earth FULL  %.3f ms  %.3e ev/s" % (nsi is not None, t, n / t * 1e3))
        t = timeit(lambda: ops.propagate_earth(consts, earth, 1, ev["true_energy"], ev["true_coszen"], flav=1, prob_e=pe, prob_mu=pm, want_probability=False, order=order))
        print("nsi=%s propagate_earth row   %.3f ms  %.3e ev/s" % (nsi is not None, t, n / t * 1e3))
        t = timeit(lambda: ops.reweight_hist(consts, earth, 1, 1, ev["true_energy"], ev["true_coszen"], ev["nu_flux"], ev["weights"], idx, 128, order=order))
        print("nsi=%s reweight_hist fused   %.3f ms  %.3e ev/s" % (nsi is not None, t, n / t * 1e3))
    t = timeit(lambda: ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]]))
    print("hist_index        %.3f ms  %.1f GB/s (28 B/event)" % (t, 28 * n / t / 1e6))
    t = timeit(lambda: ops.hist_accumulate(idx, ev["weights"], 128))
    print("hist_accumulate   %.3f ms  %.1f GB/s (12 B/event)" % (t, 12 * n / t / 1e6))
    flat = torch.rand(128, dtype=torch.float64, device=dev)
    t = timeit(lambda: ops.lookup(idx, flat))
    print("lookup            %.3f ms  %.1f GB/s (12 B/event)" % (t, 12 * n / t / 1e6))
    cz = ev["true_coszen"]
    t = timeit(lambda: ops.layers_calc(earth, cz[:2_000_000].contiguous()))
    print("layers_calc 2e6   %.3f ms  %.1f GB/s (%d B/event)" % (t, (8 + 2 * 8 * earth.max_layers + 4) * 2e6 / t / 1e6, 8 + 16 * earth.max_layers + 4))
