"""oracle/honda.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's integral-preserving Honda flux interpolation
(icecube/pisa, pisa/utils/flux_weights.py:50-130 ``load_2d_honda_table`` and :267-350
``calculate_2d_flux_weights``; call site pisa/stages/flux/honda_ip.py:86-104).

The arithmetic of this path lives in a third-party dependency that is not vendored under
/root/reference: SciPy's FITPACK wrappers ``scipy.interpolate.splrep`` / ``splev`` (the reference pins
only ``scipy>=1.6``, setup.py; this image has scipy 1.18.1).  The oracle therefore CALLS the same two
functions in the same sequence as the reference does, per event:
    spline_vals[j+1] = splev(log10 E, energy_spline[j], der=1)        j = 0..19 coszen rows
    int_vals         = cumsum(spline_vals) * 0.1
    flux             = splev(coszen, splrep(linspace(-1, 1, 21), int_vals, s=0), der=1) / E**enpow
Parity status: PINNED against outputs of the unmodified reference functions
(tests/golden/ref_honda_f8.npz, made by tests/golden/make_golden.py).
Pure-Python per-event loop: small cases only (~0.7 ms per event and primary).
"""
import numpy as np
from scipy import interpolate

N_CZ = 20


def honda_2d_flux(true_energy, true_coszen, energy_splines, enpow=1):
    """Flux of ONE primary at each (energy, coszen); ``energy_splines``: {"%.2f" % cz: tck} as built by
    ``load_2d_honda_table`` (the product's ``pisa_b200.utils.flux_weights.load_2d_table`` makes the same
    ``splrep`` calls; the golden test checks both against the reference)."""
    e = np.asarray(true_energy, dtype=np.float64)
    cz = np.asarray(true_coszen, dtype=np.float64)
    if e.shape != cz.shape:
        raise ValueError("length of energy and coszen arrays must match")
    if not ((cz >= -1.0).all() and (cz <= 1.0).all()):
        raise ValueError("Not all coszens found between -1 and 1")
    keys = ["%.2f" % x for x in np.linspace(-0.95, 0.95, N_CZ)]
    nodes = np.linspace(-1, 1, N_CZ + 1)
    out = np.empty_like(e)
    vals = np.zeros(N_CZ + 1)
    for i in range(e.size):
        loge = np.log10(e[i])
        for j, key in enumerate(keys):
            vals[j + 1] = interpolate.splev(loge, energy_splines[key], der=1)
        tck = interpolate.splrep(nodes, np.cumsum(vals) * 0.1, s=0)
        out[i] = interpolate.splev(cz[i], tck, der=1) / np.power(e[i], enpow)
    return out
