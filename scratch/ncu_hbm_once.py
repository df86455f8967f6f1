"""One launch of every HBM-bound / stage kernel at 1e8 events, for a single `ncu --set full` capture
(scratch/gpu_prof_hbm.sh).  Each op is called once warm (before cudaProfilerStart) and once inside the range."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import ops
from pisa_b200.stages.osc.layers import Layers
from pisa_b200.utils import synthetic as syn
from pisa_b200.utils.flux_weights import HondaTable2D
dev = torch.device("cuda:0")
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
L = Layers(os.path.join(ROOT, "pisa_b200/resources/osc/PREM_12layer.dat"), 2.0, 20.0); L.setElecFrac(0.4656, 0.4656, 0.4957)
earth = L.earth_struct()
ev = syn.make_events_torch(n, 3, np.float64, dev)
binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
coords = [ev["reco_energy"], ev["reco_coszen"], ev["pid"]]
idx = ops.hist_index(binning, coords)
w = ev["weights"]
flat = torch.rand(128, dtype=torch.float64, device=dev)
lo = torch.empty(n, dtype=torch.float64, device=dev)
pe = torch.rand(n, dtype=torch.float64, device=dev); pm = torch.rand(n, dtype=torch.float64, device=dev)
ww = w.clone()
nb_nom = (ev["nu_flux"] * 0.8).contiguous(); fo = torch.empty_like(ev["nu_flux"])
HT = HondaTable2D("flux/honda-2015-spl-solmin-aa.d")
nu_o = torch.empty_like(ev["nu_flux"]); nb_o = torch.empty_like(ev["nu_flux"])
terms = ops.flux_barr_terms(ev["true_energy"], ev["true_coszen"])
cz_small = ev["true_coszen"][:2_000_000].contiguous()
def all_ops():
    ops.hist_index(binning, coords, out=idx)
    ops.hist_accumulate(idx, w, 128)
    ops.lookup(idx, flat, out=lo)
    ops.apply_osc_weights(ev["nu_flux"], pe, pm, ww)
    ops.flux_barr_apply(terms, ev["nu_flux"], nb_nom, 1, 1.03, 0.97, 0.05, 0.3, -0.2, out=fo)
    ops.flux_honda_2d(HT, ev["true_energy"], ev["true_coszen"], nu_o, nb_o)
    ops.layers_calc(earth, cz_small)
all_ops(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
all_ops(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
