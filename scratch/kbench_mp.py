"""FP32-mode kernel tuning harness (GPU box): builds library variants with -D flags, times the fused kernel on one
pre-sorted 8.33e6-event container in f64 / f32-mixed / f32-fp64math, and reports the accuracy of the mixed mode
against the FP64 kernel and (on 1e6 events) against the CPU oracle."""
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

def run(path, n=8_333_333, nsi=False, oracle_check=False):
    _lib._lib = None; B.LIB = path; _lib._build.LIB = path
    from pisa_b200 import ops
    from pisa_b200.stages.osc.layers import Layers
    from pisa_b200.utils import synthetic as syn
    ops._workspaces.clear()
    dev = torch.device("cuda:0")
    L = Layers(os.path.join(ROOT, "pisa_b200/resources/osc/PREM_12layer.dat"), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = L.earth_struct()
    dm, mix, mp = syn.osc_matrices(nsi=syn.STD_NSI if nsi else None)
    consts = ops.OscConsts.from_matrices(dm, mix, mp)
    ev32 = syn.make_events_torch(n, 3, np.float32, dev)
    binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
    idx = ops.hist_index(binning, [ev32["reco_energy"], ev32["reco_coszen"], ev32["pid"]])
    o = ops.layer_order(earth, ev32["true_coszen"]).long()
    ev32 = {k: v[o].contiguous() for k, v in ev32.items()}; idx = idx[o].contiguous()
    ev64 = {k: v.double() for k, v in ev32.items()}
    res = {}
    for name, ev, math in (("f64", ev64, None), ("f32mixed", ev32, "mixed"), ("f32fp64", ev32, "fp64")):
        if math: ops.set_f32_math(math)
        a = (consts, earth, 1, 1, ev["true_energy"], ev["true_coszen"], ev["nu_flux"], ev["weights"], idx, 128)
        t = timeit(lambda: ops.reweight_hist(*a))
        h, h2 = ops.reweight_hist(*a)
        res[name] = (t, h.clone())
    ops.set_f32_math("mixed")
    m = 2_000_000
    p64, _, _ = ops.propagate_earth(consts, earth, 1, ev64["true_energy"][:m].contiguous(), ev64["true_coszen"][:m].contiguous())
    p32, _, _ = ops.propagate_earth(consts, earth, 1, ev32["true_energy"][:m].contiguous(), ev32["true_coszen"][:m].contiguous())
    _, pe32, pm32 = ops.propagate_earth(consts, earth, 1, ev32["true_energy"][:m].contiguous(), ev32["true_coszen"][:m].contiguous(), flav=1, want_probability=False)
    d_full = float((p32.double() - p64).abs().max()); d_row = float(max((pe32.double() - p64[:, 0, 1]).abs().max(), (pm32.double() - p64[:, 1, 1]).abs().max()))
    t_full32 = timeit(lambda: ops.propagate_earth(consts, earth, 1, ev32["true_energy"], ev32["true_coszen"]))
    t_full64 = timeit(lambda: ops.propagate_earth(consts, earth, 1, ev64["true_energy"], ev64["true_coszen"]))
    relh = float(((res["f32mixed"][1] - res["f64"][1]).abs() / res["f64"][1].abs().clamp_min(1e-300)).max())
    out = dict(ms_f64=res["f64"][0], ms_f32mixed=res["f32mixed"][0], ms_f32fp64=res["f32fp64"][0],
               speedup=res["f64"][0] / res["f32mixed"][0], evs_f32=n / res["f32mixed"][0] * 1e3,
               dP_full_vs_gpu64=d_full, dP_row_vs_gpu64=d_row, hist_rel=relh, ms_full32=t_full32, ms_full64=t_full64)
    if oracle_check:
        import oracle
        k = 1_000_000
        # events in the order of the sort put the deepest first: take a strided sample over all directions
        sel = torch.arange(0, n, n // k, device=dev)[:k]
        e = ev32["true_energy"][sel].contiguous(); cz = ev32["true_coszen"][sel].contiguous()
        OL = oracle.OracleLayers(np.loadtxt(os.path.join(ROOT, "pisa_b200/resources/osc/PREM_12layer.dat")), 2.0, 20.0)
        OL.setElecFrac(0.4656, 0.4656, 0.4957)
        _, den, dis = OL.calcLayers(cz.double().cpu().numpy())
        zc = np.zeros((3, 3), complex)
        for nubar in (1, -1):
            ref = oracle.propagate_array(dm, mix, mp, -1, zc, np.zeros((3, 3)), nubar, e.double().cpu().numpy(), den, dis, n_threads=os.cpu_count())
            p, _, _ = ops.propagate_earth(consts, earth, nubar, e, cz)
            d = np.abs(p.double().cpu().numpy() - ref).max(axis=(1, 2))
            out["oracle_max_dP_nubar%+d" % nubar] = float(d.max()); out["oracle_q999_nubar%+d" % nubar] = float(np.quantile(d, 0.999))
    return out

if __name__ == "__main__":
    variants = json.loads(sys.argv[1]) if len(sys.argv) > 1 else {"base": []}
    first = True
    for tag, flags in variants.items():
        try:
            p = build_variant(tag, flags)
            for nsi in (False, True):
                r = run(p, nsi=nsi, oracle_check=first)
                print("%-22s nsi=%d " % (tag, nsi) + " ".join("%s=%.4g" % kv for kv in r.items()), flush=True)
            first = False
        except Exception as e:
            print(tag, "FAILED", repr(e)[:300], flush=True)
