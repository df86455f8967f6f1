"""BASELINE configs[4]: theta23 x dm31 scan, every point = one full template (osc + hist over all resident events)
+ device-side mod_chi2.  Reports templates/s and events/s through pisa_b200.scan.scan_chi2 (no host sync in the loop)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import ops, scan
from pisa_b200.engine import ReweightEngine
from pisa_b200.stages.osc.layers import Layers
from pisa_b200.utils import synthetic as syn

dev = torch.device("cuda:0")
L = Layers(os.path.join(ROOT, "pisa_b200/resources/osc/PREM_12layer.dat"), 2.0, 20.0); L.setElecFrac(0.4656, 0.4656, 0.4957)
binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
p = syn.NUFIT20_NH
fixed = dict(theta12=np.deg2rad(p["theta12"]), theta13=np.deg2rad(p["theta13"]), deltacp=np.deg2rad(p["deltacp"]), dm21=p["deltam21"])
for n_total in (1.2e5, 1.2e6, 1.2e7, 1.2e8):
    per = int(n_total) // 12
    eng = ReweightEngine(L.earth_struct(), 128, np.float64, dev)
    for c, (name, nubar, flav) in enumerate(syn.CONTAINERS):
        ev = syn.make_events_torch(per, seed=c + 1, dtype=np.float64, device=dev)
        idx = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
        eng.add_container(name, nubar, flav, ev["true_energy"], ev["true_coszen"], ev["nu_flux"], ev["weights"], idx)
    eng.set_scales([1.0 + 0.01 * c for c in range(12)])
    observed = scan.asimov(eng, scan.osc_consts(theta23=np.deg2rad(p["theta23"]), dm31=p["deltam31"], **fixed))
    pts = [(t, d) for t in np.deg2rad(np.linspace(35, 55, 10)) for d in np.linspace(2.2e-3, 2.7e-3, 10)]
    for batch in (1, 100):
        scan.scan_chi2(eng, observed, pts, fixed, batch=batch); torch.cuda.synchronize()
        t0 = time.perf_counter()
        chi2 = scan.scan_chi2(eng, observed, pts, fixed, batch=batch)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = pts[int(chi2.argmin())]
        print("events/template %11d  templates/launch %3d  100-point scan %8.2f ms  %8.1f templates/s  %.3e events/s  chi2_min %.3e"
              % (12 * per, batch, 1e3 * dt, len(pts) / dt, 12 * per * len(pts) / dt, float(chi2.min())), flush=True)
    del eng
