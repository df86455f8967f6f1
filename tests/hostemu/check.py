"""Compare the host-emulated device math with the CPU oracle.

Test infrastructure (tests/test_device_math_emulation.py builds libemu.so with g++ and calls `run`); also a
stand-alone development tool:  g++ -O2 -fopenmp -shared -fPIC -std=c++17 -DPISAB_HOST_EMU -include cuda_shim.h
-I../../pisa_b200/csrc -I../../include -o libemu.so emu.cpp && python check.py"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle
from pisa_b200._lib import OscConsts, Earth
from pisa_b200.utils import synthetic as syn

emu = None


def load(path=None):
    global emu
    emu = ctypes.CDLL(path or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libemu.so"))
    return emu


def layers_obj(model="PREM_12layer.dat", depth=2.0, height=20.0):
    prem = np.loadtxt(os.path.join(ROOT, "pisa_b200", "resources", "osc", model))
    L = oracle.OracleLayers(prem, depth, height)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    return L

def earth_struct(L):
    return Earth.from_arrays(L.radii, L.rhos, L.coszen_limit, L.r_detector, L.max_layers)

def run(n=200000, nsi=False, nubar=1, lri=None, model="PREM_12layer.dat", seed=0, depth=2.0):
    rng = np.random.default_rng(seed)
    e = 10 ** rng.uniform(0, 3, n); cz = rng.uniform(-1, 1, n)
    L = layers_obj(model, depth)
    dm, mix, mp = syn.osc_matrices(nsi=syn.STD_NSI if nsi else None)
    zc = np.zeros((3, 3), complex); lr = np.zeros((3, 3)) if lri is None else lri
    c = OscConsts.from_matrices(dm, mix, mp, -1, zc, lr)
    E = earth_struct(L)
    prob = np.empty((n, 3, 3)); pe = np.empty(n); pm = np.empty(n)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    worst = 0
    for flav in (0, 1, 2):
        rc = emu.emu_propagate(ctypes.byref(c), ctypes.byref(E), nubar, flav, vp(e), vp(cz), ctypes.c_int64(n), vp(prob), vp(pe), vp(pm))
        assert rc == 0
        _, den, dis = L.calcLayers(cz)
        ref = oracle.propagate_array(dm, mix, mp, -1, zc, lr, nubar, e, den, dis)
        d_full = np.abs(prob - ref).max()
        d_row = max(np.abs(pe - ref[:, 0, flav]).max(), np.abs(pm - ref[:, 1, flav]).max())
        worst = max(worst, d_full, d_row)
    print("n=%d nsi=%s nubar=%+d lri=%s %s: max|dP| = %.3e" % (n, nsi, nubar, lri is not None, model, worst))
    return worst

def run_mp(n=200000, nsi=False, nubar=1, lri=None, model="PREM_12layer.dat", seed=0, depth=2.0, e_max=3.0, verbose=True):
    """FP32 mode (emu_propagate_mp) against the FP64 oracle on float32-rounded inputs: max |dP| (BASELINE: 1e-5)."""
    rng = np.random.default_rng(seed)
    e = (10 ** rng.uniform(0, e_max, n)).astype(np.float32).astype(np.float64)
    cz = rng.uniform(-1, 1, n).astype(np.float32).astype(np.float64)
    L = layers_obj(model, depth)
    dm, mix, mp = syn.osc_matrices(nsi=syn.STD_NSI if nsi else None)
    zc = np.zeros((3, 3), complex); lr = np.zeros((3, 3)) if lri is None else lri
    c = OscConsts.from_matrices(dm, mix, mp, -1, zc, lr)
    E = earth_struct(L)
    prob = np.empty((n, 3, 3)); pe = np.empty(n); pm = np.empty(n)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _, den, dis = L.calcLayers(cz)
    ref = oracle.propagate_array(dm, mix, mp, -1, zc, lr, nubar, e, den, dis)
    worst = 0
    for flav in (0, 1, 2):
        rc = emu.emu_propagate_mp(ctypes.byref(c), ctypes.byref(E), nubar, flav, vp(e), vp(cz), ctypes.c_int64(n),
                                  vp(prob) if flav == 0 else None, vp(pe), vp(pm))
        assert rc == 0
        if flav == 0:
            d = np.abs(prob - ref).max(axis=(1, 2))
            worst = max(worst, d.max())
            if verbose:
                i = int(d.argmax())
                print("  full: max %.3e at E=%.3f cz=%.4f; 99.9%% %.2e; mean %.2e" % (d.max(), e[i], cz[i], np.quantile(d, 0.999), d.mean()))
        d_row = max(np.abs(pe - ref[:, 0, flav]).max(), np.abs(pm - ref[:, 1, flav]).max())
        worst = max(worst, d_row)
    if verbose:
        print("MP n=%d nsi=%s nubar=%+d lri=%s %s: max|dP| = %.3e" % (n, nsi, nubar, lri is not None, model, worst))
    return worst

def run_decay(n=50000, nsi=False, nubar=1, lri=None, alpha3=1e-4, mat_decay=None, model="PREM_12layer.dat", seed=0, depth=2.0,
              e_max=3.0):
    """Decay branch (emu_propagate_decay: prob3_decay.cuh through the Earth walk) against the oracle's restatement of
    the reference's eigvals branch: max |dP| over the full matrix and the row outputs."""
    rng = np.random.default_rng(seed)
    e = 10 ** rng.uniform(0, e_max, n); cz = rng.uniform(-1, 1, n)
    L = layers_obj(model, depth)
    dm, mix, mp = syn.osc_matrices(nsi=syn.STD_NSI if nsi else None)
    md = np.zeros((3, 3), complex)
    md[2, 2] = -1j * alpha3
    if mat_decay is not None:
        md = np.asarray(mat_decay, complex)
    lr = np.zeros((3, 3)) if lri is None else lri
    c = OscConsts.from_matrices(dm, mix, mp, 1, md, lr)
    E = earth_struct(L)
    prob = np.empty((n, 3, 3)); pe = np.empty(n); pm = np.empty(n)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _, den, dis = L.calcLayers(cz)
    ref = oracle.propagate_array(dm, mix, mp, 1, md, lr, nubar, e, den, dis, n_threads=8)
    worst = 0
    for flav in (0, 1, 2):
        rc = emu.emu_propagate_decay(ctypes.byref(c), ctypes.byref(E), nubar, flav, vp(e), vp(cz), ctypes.c_int64(n),
                                     vp(prob) if flav == 0 else None, vp(pe), vp(pm))
        assert rc == 0
        if flav == 0:
            worst = max(worst, np.abs(prob - ref).max())
        worst = max(worst, np.abs(pe - ref[:, 0, flav]).max(), np.abs(pm - ref[:, 1, flav]).max())
    print("decay n=%d nsi=%s nubar=%+d lri=%s alpha3=%g %s: max|dP| = %.3e (min row sum %.3f)" % (
        n, nsi, nubar, lri is not None, alpha3, model, worst, ref.sum(axis=2).min()))
    return worst


def run_pairs(n=100000, nsi=False, nubar=1, seed=0):
    """Two-events-per-thread form of the FP32 mode (emu_propagate_mp_pairs, lane-packed float part) against the
    one-event form on events sorted by the number of crossed shells: returns (all equal-class pairs bit-identical,
    all unequal pairs flagged, no equal pair flagged)."""
    rng = np.random.default_rng(seed)
    e = (10 ** rng.uniform(0, 3, n)).astype(np.float32).astype(np.float64)
    cz = rng.uniform(-1, 1, n).astype(np.float32).astype(np.float64)
    L = layers_obj()
    k = (np.asarray(L.coszen_limit)[None, :] > cz[:, None]).sum(axis=1)
    order = np.argsort(-k, kind="stable")
    e, cz, k = e[order], cz[order], k[order]
    dm, mix, mp = syn.osc_matrices(nsi=syn.STD_NSI if nsi else None)
    c = OscConsts.from_matrices(dm, mix, mp)
    E = earth_struct(L)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    same = k[0::2] == k[1::2]
    ok = np.repeat(same, 2)
    identical, flagged_ok = True, True
    for flav in (0, 1, 2):
        pe, pm, pe2, pm2 = np.empty(n), np.empty(n), np.empty(n), np.empty(n)
        mm = np.zeros(n // 2, dtype=np.int32)
        assert emu.emu_propagate_mp(ctypes.byref(c), ctypes.byref(E), nubar, flav, vp(e), vp(cz), ctypes.c_int64(n), None, vp(pe), vp(pm)) == 0
        assert emu.emu_propagate_mp_pairs(ctypes.byref(c), ctypes.byref(E), nubar, flav, vp(e), vp(cz), ctypes.c_int64(n), vp(pe2), vp(pm2), vp(mm)) == 0
        identical = identical and np.array_equal(pe[ok], pe2[ok]) and np.array_equal(pm[ok], pm2[ok])
        flagged_ok = flagged_ok and bool(mm[~same].all()) and not bool(mm[same].any())
    return identical, flagged_ok, int((~same).sum())


if __name__ == "__main__":
    load()
    w = 0
    for nubar in (1, -1):
        for nsi in (False, True):
            w = max(w, run(nsi=nsi, nubar=nubar))
    w = max(w, run(lri=np.diag([1e-14, -1e-14, 0.0])))
    w = max(w, run(model="PREM_10layer.dat", n=50000))
    w = max(w, run(model="PREM_4layer.dat", n=50000, depth=10.0))
    print("worst", w)
