"""One sequential template at analysis sizes: pipelined time per evaluate_chi2 call against the sample size (how many
'rounds' of resident threads the sample needs), FP64 and FP32 mode.  Run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel durations."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("PISAB_LIB"):   # A/B of prebuilt library variants (scratch/variants/)
    from pisa_b200 import _lib, build as _B
    _B.LIB = os.path.abspath(os.environ["PISAB_LIB"])
    _lib._build.LIB = _B.LIB
from pisa_b200 import ops, scan
from pisa_b200.engine import ReweightEngine
from pisa_b200.stages.osc.layers import Layers
from pisa_b200.utils import synthetic as syn
dev = torch.device("cuda:0")
L = Layers(os.path.join(ROOT, "pisa_b200/resources/osc/PREM_12layer.dat"), 2.0, 20.0); L.setElecFrac(0.4656, 0.4656, 0.4957)
binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
p = syn.NUFIT20_NH
fixed = dict(theta12=np.deg2rad(p["theta12"]), theta13=np.deg2rad(p["theta13"]), deltacp=np.deg2rad(p["deltacp"]), dm21=p["deltam21"])
consts = scan.osc_consts(theta23=0.74, dm31=2.5e-3, **fixed)
N = int(os.environ.get("SWEEP_REPS", "200"))
sizes = [int(x) for x in os.environ.get("SWEEP_SIZES", "3000,6000,10000,12500,25000,100000").split(",")]
for dtype in (np.float64, np.float32):
    for per in sizes:
        eng = ReweightEngine(L.earth_struct(), 128, dtype, dev)
        for c, (name, nubar, flav) in enumerate(syn.CONTAINERS):
            ev = syn.make_events_torch(per, seed=c + 1, dtype=dtype, device=dev)
            idx = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
            eng.add_container(name, nubar, flav, ev["true_energy"], ev["true_coszen"], ev["nu_flux"], ev["weights"], idx)
        obs = scan.asimov(eng, consts)
        out = torch.empty(1, dtype=torch.float64, device=dev)
        f = lambda: eng.evaluate_chi2(consts, obs, chi2_out=out)
        for _ in range(5): f()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(N): f()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / N * 1e6
        print("%s %8d events/template: %6.1f us per evaluate_chi2 (pipelined) = %.2fe9 events/s" % (
            np.dtype(dtype).name, 12 * per, dt, 12 * per / dt / 1e3), flush=True)
