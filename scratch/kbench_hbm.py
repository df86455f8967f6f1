"""HBM-bound kernels at a size far above L2 (1e8 events): achieved GB/s vs MEASURED_PEAKS.json."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import ops
from pisa_b200.stages.osc.layers import Layers
from pisa_b200.utils import synthetic as syn
from kbench_stage import timeit  # noqa
peak = 6455.6
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = torch.device("cuda:0")
L = Layers(os.path.join(ROOT, "pisa_b200/resources/osc/PREM_12layer.dat"), 2.0, 20.0); L.setElecFrac(0.4656, 0.4656, 0.4957)
earth = L.earth_struct()
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
ev = syn.make_events_torch(n, 3, np.float64, dev)
binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
coords = [ev["reco_energy"], ev["reco_coszen"], ev["pid"]]
idx = ops.hist_index(binning, coords)
w = ev["weights"]
def rep(name, t, bytes_per_event):
    gbs = bytes_per_event * n / t / 1e6
    print("%-28s %9.3f ms  %8.1f GB/s  %5.1f%% of measured %.0f GB/s  (%d B/event)" % (name, t, gbs, 100 * gbs / peak, peak, bytes_per_event), flush=True)
out_idx = torch.empty_like(idx)
rep("hist_index (edges,lin,lin)", timeit(lambda: ops.hist_index(binning, coords, out=out_idx)), 28)
rep("hist_accumulate w+w2", timeit(lambda: ops.hist_accumulate(idx, w, 128)), 12)
rep("hist_accumulate counts", timeit(lambda: ops.hist_accumulate(idx, None, 128, want_w2=False)), 4)
flat = torch.rand(128, dtype=torch.float64, device=dev)
lo = torch.empty(n, dtype=torch.float64, device=dev)
rep("lookup width 1", timeit(lambda: ops.lookup(idx, flat, out=lo)), 12)
pe = torch.rand(n, dtype=torch.float64, device=dev); pm = torch.rand(n, dtype=torch.float64, device=dev)
ww = w.clone()
rep("apply_osc_weights", timeit(lambda: ops.apply_osc_weights(ev["nu_flux"], pe, pm, ww)), 48)
a = torch.empty(n * 4, dtype=torch.float64, device=dev); b = torch.empty_like(a)
t = timeit(lambda: b.copy_(a)); print("torch copy 3.2 GB            %9.3f ms  %8.1f GB/s" % (t, 2 * a.numel() * 8 / t / 1e6))
nb_nom = (ev["nu_flux"] * 0.8).contiguous(); fo = torch.empty_like(ev["nu_flux"])
rep("flux_barr_simple", timeit(lambda: ops.flux_barr_simple(ev["true_energy"], ev["true_coszen"], ev["nu_flux"], nb_nom, 1, 1.03, 0.97, 0.05, 0.3, -0.2, out=fo)), 64)
from pisa_b200.utils.flux_weights import HondaTable2D
HT = HondaTable2D("flux/honda-2015-spl-solmin-aa.d")
nu_o = torch.empty_like(ev["nu_flux"]); nb_o = torch.empty_like(ev["nu_flux"])
rep("flux_honda_2d (4 primaries)", timeit(lambda: ops.flux_honda_2d(HT, ev["true_energy"], ev["true_coszen"], nu_o, nb_o)), 48)
terms = ops.flux_barr_terms(ev["true_energy"], ev["true_coszen"])
rep("flux_barr_terms (setup)", timeit(lambda: ops.flux_barr_terms(ev["true_energy"], ev["true_coszen"], out=terms)), 48)
rep("flux_barr_apply (per template)", timeit(lambda: ops.flux_barr_apply(terms, ev["nu_flux"], nb_nom, 1, 1.03, 0.97, 0.05, 0.3, -0.2, out=fo)), 80)
