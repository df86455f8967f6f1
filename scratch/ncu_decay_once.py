"""One decay template over 4.8e7 events (12 containers) for an ncu capture of reweight_hist_decay_kernel."""
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
dm, mix, mat_pot = syn.osc_matrices()
md = np.zeros((3, 3), complex); md[2, 2] = -1e-4j
dec = ops.OscConsts.from_matrices(dm, mix, mat_pot, 1, md)
binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
eng = ReweightEngine(L.earth_struct(), syn.DRAGON_NBINS, np.float64, dev)
n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 48_000_000
for c, (name, nubar, flav) in enumerate(syn.CONTAINERS):
    ev = syn.make_events_torch(n_total // 12, seed=c + 1, dtype=np.float64, device=dev)
    index = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
    eng.add_container(name, nubar, flav, true_energy=ev["true_energy"], true_coszen=ev["true_coszen"],
                      nu_flux=ev["nu_flux"], weights=ev["weights"], index=index)
for _ in range(3):
    out = eng.evaluate(dec)
torch.cuda.synchronize()
print("sum_w", float(out[:, 0].sum()))
