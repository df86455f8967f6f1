"""Template over the 40 x 40 x 2 stress binning (3200 bins) at bench size: per-evaluate time; run under
`ncu --metrics gpu__time_duration.sum` for the launch breakdown.  usage: python scratch/large_bins_probe.py [n] [f64|f32]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import ops
from pisa_b200.engine import ReweightEngine
from pisa_b200.stages.osc.layers import Layers
from pisa_b200.utils import synthetic as syn
n_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
dtype = np.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else np.float64
dev = torch.device("cuda:0")
L = Layers(os.path.join(ROOT, "pisa_b200/resources/osc/PREM_12layer.dat"), 2.0, 20.0)
L.setElecFrac(0.4656, 0.4656, 0.4957)
dims = [dict(name="reco_energy", kind="log", n_bins=40, lo=5.62341325, hi=56.23413252),
        dict(name="reco_coszen", kind="lin", n_bins=40, lo=-1.0, hi=1.0), dict(name="pid", kind="lin", n_bins=2, lo=-0.5, hi=1.5)]
binning, keep = ops.make_binning(dims, dev)
dm, mix, mat_pot = syn.osc_matrices()
consts = ops.OscConsts.from_matrices(dm, mix, mat_pot)
for n_bins, bn in ((3200, binning),):
    eng = ReweightEngine(L.earth_struct(), n_bins, dtype, dev)
    per = n_total // 12
    for i, (name, nubar, flav) in enumerate(syn.CONTAINERS):
        t = syn.make_events_torch(per, 100 + i, dtype, dev)
        idx = ops.hist_index(bn, [t["reco_energy"], t["reco_coszen"], t["pid"]])
        eng.add_container(name, nubar, flav, t["true_energy"], t["true_coszen"], t["nu_flux"], t["weights"], idx)
        del t
    for _ in range(3): eng.evaluate(consts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = eng.evaluate(consts)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%s, %d bins, %d events: %.3f ms per template = %.2fe9 events/s; sum w %.6e" % (
        np.dtype(dtype).name, n_bins, eng.n_events, ms, eng.n_events / ms / 1e6, float(out[:, 0].sum())))
