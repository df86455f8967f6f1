"""Kernel tuning harness (run on the GPU box): builds variants of the library with -D flags and
times the fused kernel on one 8.3M-event container."""
import ctypes, os, subprocess, sys, glob, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pisa_b200 import _lib, build as B

def build_variant(tag, flags):
    out = "/tmp/libpisa_%s.so" % tag
    cmd = ["nvcc"] + B.NVCC_FLAGS + flags + ["-o", out] + B.sources()
    env = dict(os.environ); env.pop("CC", None)
    subprocess.check_call(cmd, cwd=B.CSRC, env=env)
    return out

def run(path, n=8_333_333, reps=5):
    _lib._lib = None
    B.LIB = path
    _lib._build.LIB = path
    from pisa_b200 import ops
    from pisa_b200.stages.osc.layers import Layers
    from pisa_b200.utils import synthetic as syn
    dev = torch.device("cuda:0")
    L = Layers(os.path.join(ROOT, "pisa_b200/resources/osc/PREM_12layer.dat"), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = L.earth_struct()
    dm, mix, mp = syn.osc_matrices()
    consts = ops.OscConsts.from_matrices(dm, mix, mp)
    ev = syn.make_events_torch(n, 3, np.float64, dev)
    binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
    idx = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
    order = ops.layer_order(earth, ev["true_coszen"])
    args = (consts, earth, 1, 1, ev["true_energy"], ev["true_coszen"], ev["nu_flux"], ev["weights"], idx, 128)
    o = order.long()
    sargs = (consts, earth, 1, 1, ev["true_energy"][o].contiguous(), ev["true_coszen"][o].contiguous(),
             ev["nu_flux"][o].contiguous(), ev["weights"][o].contiguous(), idx[o].contiguous(), 128)
    out = []
    for a, kw in ((args, dict(order=order)), (sargs, dict())):
        for _ in range(3):
            ops.reweight_hist(*a, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); h, h2 = ops.reweight_hist(*a, **kw); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out.append((min(ts), float(h.sum())))
    return out

if __name__ == "__main__":
    variants = json.loads(sys.argv[1]) if len(sys.argv) > 1 else {"base": []}
    for tag, flags in variants.items():
        try:
            p = build_variant(tag, flags)
            (ms, chk), (ms2, chk2) = run(p)
            print("%-28s order: %8.3f ms %.3e ev/s | presorted: %8.3f ms %.3e ev/s | checksums %.12e %.12e" % (
                tag, ms, 8_333_333 / ms * 1e3, ms2, 8_333_333 / ms2 * 1e3, chk, chk2), flush=True)
        except Exception as e:
            print(tag, "FAILED", repr(e)[:300], flush=True)
