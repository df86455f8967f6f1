"""FP32-mode pair kernel variants (GPU box): builds library variants with -D flags and times ReweightEngine.evaluate on
12 float32 containers of 2e6 events (pair-aligned), plus the scalar FP32-mode kernel and the FP64 kernel as yardsticks."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scratch"))
from pisa_b200 import _lib, build as B
from kbench import build_variant

def timeit(f, reps=5):
    for _ in range(3): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

def run(path, per=2_000_000):
    _lib._lib = None; B.LIB = path; _lib._build.LIB = path
    from pisa_b200 import ops
    from pisa_b200.engine import ReweightEngine
    from pisa_b200.stages.osc.layers import Layers
    from pisa_b200.utils import synthetic as syn
    ops._workspaces.clear()
    dev = torch.device("cuda:0")
    L = Layers(os.path.join(ROOT, "pisa_b200/resources/osc/PREM_12layer.dat"), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = L.earth_struct()
    binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
    out = {}
    for nsi in (False, True):
        dm, mix, mp = syn.osc_matrices(nsi=syn.STD_NSI if nsi else None)
        consts = ops.OscConsts.from_matrices(dm, mix, mp)
        e32 = ReweightEngine(earth, 128, np.float32, dev); e64 = ReweightEngine(earth, 128, np.float64, dev)
        for c, (name, nubar, flav) in enumerate(syn.CONTAINERS):
            ev = syn.make_events_torch(per, seed=c + 1, dtype=np.float32, device=dev)
            idx = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
            e32.add_container(name, nubar, flav, ev["true_energy"], ev["true_coszen"], ev["nu_flux"], ev["weights"], idx)
            e64.add_container(name, nubar, flav, ev["true_energy"].double(), ev["true_coszen"].double(), ev["nu_flux"].double(), ev["weights"].double(), idx)
        n = 12 * per
        ops.set_f32_math("mixed")
        t_pair = timeit(lambda: e32.evaluate(consts)); h_pair = e32.evaluate(consts).clone()
        for b in e32.blocks: b.flags = 0
        e32._batches = None
        t_one = timeit(lambda: e32.evaluate(consts)); h_one = e32.evaluate(consts).clone()
        t_64 = timeit(lambda: e64.evaluate(consts))
        rel = float(((h_pair - h_one).abs() / h_one.abs().clamp_min(1e-300)).max())
        out["nsi%d" % nsi] = "pair %.3f ms (%.3e ev/s, %.2fx fp64) | one-event %.3f ms (%.2fx) | fp64 %.3f ms | pair vs one-event rel %.1e" % (
            t_pair, n / t_pair * 1e3, t_64 / t_pair, t_one, t_64 / t_one, t_64, rel)
    return out

if __name__ == "__main__":
    variants = json.loads(sys.argv[1]) if len(sys.argv) > 1 else {"base": []}
    for tag, flags in variants.items():
        try:
            r = run(build_variant(tag, flags))
            for k, v in r.items(): print("%-18s %s  %s" % (tag, k, v), flush=True)
        except Exception as e:
            print(tag, "FAILED", repr(e)[:300], flush=True)
