"""Where the time of ONE small template goes (sequential-minimiser case)."""
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
# sequential gradient fit (objective + central differences in one launch per iteration)
import time as _time
from pisa_b200 import scan as _scan
def fit_timing(eng, observed, fixed, start, bounds):
    _scan.fit_chi2(eng, observed, start, fixed, bounds=bounds)
    torch.cuda.synchronize(); t0 = _time.perf_counter()
    res = _scan.fit_chi2(eng, observed, start, fixed, bounds=bounds)
    dt = _time.perf_counter() - t0
    print("fit theta23, dm31: %d objective calls (%d templates) in %.2f ms = %.1f us per call; chi2 %.2e; x = %s" % (
        res.nfev, res.n_templates, dt * 1e3, dt / res.nfev * 1e6, res.fun, res.x))

for per in (10_000, 100_000):
    eng = ReweightEngine(L.earth_struct(), 128, np.float64, dev)
    for c, (name, nubar, flav) in enumerate(syn.CONTAINERS):
        ev = syn.make_events_torch(per, seed=c + 1, dtype=np.float64, device=dev)
        idx = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
        eng.add_container(name, nubar, flav, ev["true_energy"], ev["true_coszen"], ev["nu_flux"], ev["weights"], idx)
    consts = scan.osc_consts(theta23=0.74, dm31=2.5e-3, **fixed)
    obs = scan.asimov(eng, consts)
    N = 300
    def t(f):
        f(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(N): f()
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / N * 1e6
    out = torch.empty(1, dtype=torch.float64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.evaluate(consts); torch.cuda.synchronize(); e0.record(); eng.evaluate(consts); e1.record(); torch.cuda.synchronize()
    print("events/template %8d: ONE-CALL evaluate_chi2 (template + reduce/chi2 epilogue) %6.1f us | with osc_consts %6.1f us"
          % (12 * per, t(lambda: eng.evaluate_chi2(consts, obs, chi2_out=out)),
             t(lambda: eng.evaluate_chi2(scan.osc_consts(theta23=0.74, dm31=2.5e-3, **fixed), obs, chi2_out=out))), flush=True)
    print("events/template %8d: osc_consts %6.1f us | evaluate (host+GPU, pipelined) %6.1f us | + chi2 %6.1f us | full (consts+eval+chi2) %6.1f us | GPU time of one evaluate %6.1f us"
          % (12 * per, t(lambda: scan.osc_consts(theta23=0.74, dm31=2.5e-3, **fixed)), t(lambda: eng.evaluate(consts)),
             t(lambda: ops.template_chi2(eng.evaluate(consts), obs, out=out)),
             t(lambda: ops.template_chi2(eng.evaluate(scan.osc_consts(theta23=0.74, dm31=2.5e-3, **fixed)), obs, out=out)),
             e0.elapsed_time(e1) * 1e3), flush=True)
    fit_timing(eng, obs, fixed, dict(theta23=np.deg2rad(39.0), dm31=2.6e-3),
               dict(theta23=(np.deg2rad(30.0), np.deg2rad(45.0)), dm31=(2.0e-3, 3.0e-3)))
