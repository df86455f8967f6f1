"""Fit step with floating flux systematics at bench size: flux.barr_simple as its own pass (pisab_flux_barr_apply_batch,
writes nu_flux) + template, against the template kernel that evaluates it in registers (PISAB_CONTAINER_FLUX_SYS).
usage: python scratch/bench_flux_fold.py [n_events_total] [f64|f32]"""
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
binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
dm, mix, mat_pot = syn.osc_matrices()
consts = ops.OscConsts.from_matrices(dm, mix, mat_pot)
eng = ReweightEngine(L.earth_struct(), 128, dtype, dev)
per = n_total // 12
for i, (name, nubar, flav) in enumerate(syn.CONTAINERS):
    t = syn.make_events_torch(per, 100 + i, dtype, dev)
    idx = ops.hist_index(binning, [t["reco_energy"], t["reco_coszen"], t["pid"]])
    eng.add_container(name, nubar, flav, t["true_energy"], t["true_coszen"], t["nu_flux"].clone(), t["weights"], idx,
                      nu_flux_nominal=t["nu_flux"], nubar_flux_nominal=(t["nu_flux"] * 0.8).contiguous())
    del t
pars = dict(nue_numu_ratio=1.03, nu_nubar_ratio=0.95, delta_index=0.05, Barr_uphor_ratio=-0.5, Barr_nu_nubar_ratio=0.8)
def timed(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def staged():
    eng.set_flux_params(**pars, materialize=True); return eng.evaluate(consts)
def folded():
    eng.set_flux_params(**pars); return eng.evaluate(consts)
def fixed():
    return eng.evaluate(consts)
a = staged().clone(); b = folded().clone()
print("bit-identical:", bool(torch.equal(a, b)))
eng.set_flux_params(**pars, materialize=True)
t_fixed = timed(fixed); t_staged = timed(staged); t_fold = timed(folded)
n = eng.n_events
print("%s, %d events: template with fixed flux %.3f ms | flux pass + template %.3f ms | flux inside the template kernel %.3f ms "
      "(%.2fe9 events/s, %+.1f %% vs the two-pass form)" % (np.dtype(dtype).name, n, t_fixed, t_staged, t_fold, n / t_fold / 1e6,
      100 * (t_fold / t_staged - 1)))
