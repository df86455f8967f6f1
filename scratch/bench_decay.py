"""Timing of the neutrino-decay branch (reweight_hist_decay_kernel, prob3_earth_decay_kernel) next to the standard
kernels on the bench's C3 workload shape (12 containers, PREM_12layer, dragon binning).  CUDA events, best of 5."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import ops
from pisa_b200.engine import ReweightEngine
from pisa_b200.stages.osc.layers import Layers
from pisa_b200.utils import synthetic as syn

dev = torch.device("cuda:0")
L = Layers(os.path.join(ROOT, "pisa_b200", "resources", syn.EARTH["earth_model"]), syn.EARTH["detector_depth"], syn.EARTH["prop_height"])
L.setElecFrac(syn.EARTH["YeI"], syn.EARTH["YeO"], syn.EARTH["YeM"])
earth = L.earth_struct()
dm, mix, mat_pot = syn.osc_matrices()
md = np.zeros((3, 3), complex); md[2, 2] = -1e-4j
std = ops.OscConsts.from_matrices(dm, mix, mat_pot)
dec = ops.OscConsts.from_matrices(dm, mix, mat_pot, 1, md)
binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

for dtype in (np.float64, np.float32):
    for n_total in (1_200_000, 12_000_000, 48_000_000):
        eng = ReweightEngine(earth, syn.DRAGON_NBINS, dtype, dev)
        per = n_total // 12
        for c, (name, nubar, flav) in enumerate(syn.CONTAINERS):
            ev = syn.make_events_torch(per, seed=c + 1, dtype=dtype, device=dev)
            index = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
            eng.add_container(name, nubar, flav, true_energy=ev["true_energy"], true_coszen=ev["true_coszen"],
                              nu_flux=ev["nu_flux"], weights=ev["weights"], index=index)
        t_std = timed(lambda: eng.evaluate(std))
        t_dec = timed(lambda: eng.evaluate(dec))
        h_std, h_dec = eng.evaluate(std).clone(), eng.evaluate(dec).clone()
        print("%s %9d events: standard %.3f ms (%.2e ev/s) | decay %.3f ms (%.2e ev/s, x%.2f) | sum_w decay/std %.4f" % (
            np.dtype(dtype).name, n_total, t_std, n_total / t_std * 1e3, t_dec, n_total / t_dec * 1e3, t_dec / t_std,
            float(h_dec[:, 0].sum() / h_std[:, 0].sum())))
        del eng
    ev = syn.make_events_torch(4_000_000, seed=99, dtype=dtype, device=dev)
    for name, c in (("standard", std), ("decay", dec)):
        t = timed(lambda: ops.propagate_earth(c, earth, 1, ev["true_energy"], ev["true_coszen"]))
        print("%s propagate_earth (full 3x3) 4e6 events, %s: %.3f ms (%.2e ev/s)" % (np.dtype(dtype).name, name, t, 4e6 / t * 1e3))
