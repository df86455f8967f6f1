"""Honda atmospheric flux tables -> per-event nominal fluxes (``flux.honda_ip``).

Reference: pisa/utils/flux_weights.py -- ``load_2d_honda_table`` (:50-130) builds, for each primary and each of
the 20 coszen rows of an azimuth-averaged table, an interpolating cubic spline (``scipy.interpolate.splrep``,
``s=0``) of the flux INTEGRATED over log10(E) ("integral preserving"); ``calculate_2d_flux_weights`` (:267-350)
then does, PER EVENT: 20 spline derivatives at log10(E), a cumulative sum over coszen, a NEW interpolating
spline through those 21 points and its derivative at the event's coszen, divided by E.  That per-event spline
fit is the 23 s of setup time in the reference's published profile.

Here the table side stays on the host (it is 80 tiny splines, built once with the same ``splrep`` call so that
the B-spline coefficients are bit-identical to the reference's -- the high-energy derivatives are differences
of coefficients 8 orders of magnitude larger, so an independent solver would differ at 1e-7), and everything
per event runs in ONE kernel (``pisab_flux_honda_2d``):

* the energy derivative is FITPACK's ``splder`` restated (coefficient differences precomputed here with its
  exact formula, the 3 non-zero quadratic B-splines by ``fpbspl``'s recursion on the device);
* the per-event coszen spline fit is linear in its 21 input values, so it collapses to a fixed table of
  "cardinal spline" derivative polynomials ``D[piece][k][3]``: out = sum_k D_k(coszen) * int_vals[k].
"""
import os

import numpy as np

from pisa_b200.utils.resources import find_resource

__all__ = ["PRIMARIES", "load_2d_table", "HondaTable2D"]

PRIMARIES = ["numu", "numubar", "nue", "nuebar"]   # column order of the Honda files (flux_weights.py:46)
# output order of flux.honda_ip: nu_flux_nominal = (nue, numu), nubar_flux_nominal = (nuebar, numubar)
OUT_ORDER = ["nue", "numu", "nuebar", "numubar"]
N_CZ = 20


def load_2d_table(flux_file, enpow=1):
    """``load_2d_table`` / ``load_2d_honda_table`` (flux_weights.py:50-130,205-264) for Honda azimuth-averaged
    tables: {primary: {"%.2f" % coszen: (t, c, k)}} with the reference's own ``splrep`` call."""
    from scipy import interpolate   # host-side table construction only (see module docstring)
    if not isinstance(enpow, int):
        raise TypeError("Energy power must be an integer")
    if not isinstance(flux_file, str):
        raise TypeError("Flux file name must be a string")
    if "aa" not in flux_file:
        raise ValueError("Azimuth-averaged tables are expected")
    if "honda" not in flux_file:
        raise NotImplementedError("only Honda azimuth-averaged tables are supported by pisa_b200 (got %s)" % flux_file)
    cols = ["energy"] + PRIMARIES
    path = flux_file if os.path.exists(flux_file) else find_resource(flux_file)
    table = np.genfromtxt(path, usecols=list(range(len(cols))))
    mask = np.all(np.isnan(table) | np.equal(table, 0), axis=1)
    table = table[~mask].T
    flux_dict = dict(zip(cols, table))
    for key in flux_dict:
        flux_dict[key] = np.array(np.split(flux_dict[key], N_CZ))   # 20 coszen rows of 101 energies
    flux_dict["energy"] = flux_dict["energy"][0]
    logenergy = np.linspace(-1.025, 4.025, 102)
    spline_dict = {}
    for nutype in PRIMARIES:
        splines = {}
        cz_iter = 1
        for energyfluxlist in flux_dict[nutype]:
            int_flux = []
            tot_flux = 0.0
            int_flux.append(tot_flux)
            for energyfluxval, energyval in zip(energyfluxlist, flux_dict["energy"]):
                tot_flux += energyfluxval * np.power(energyval, enpow) * 0.05
                int_flux.append(tot_flux)
            splines["%.2f" % (1.05 - cz_iter * 0.1)] = interpolate.splrep(logenergy, int_flux, s=0)
            cz_iter += 1
        spline_dict[nutype] = splines
    spline_dict["name"] = "honda"
    return spline_dict


def _cardinal_derivative_table():
    """D[piece][k][3]: derivative (quadratic in x - breakpoint) of the interpolating cubic spline through
    the unit vector e_k on the 21 points linspace(-1, 1, 21), per polynomial piece of that spline."""
    from scipy import interpolate
    pts = np.linspace(-1, 1, N_CZ + 1)
    breaks = None
    out = []
    for k in range(N_CZ + 1):
        e = np.zeros(N_CZ + 1)
        e[k] = 1.0
        pp = interpolate.PPoly.from_spline(interpolate.splrep(pts, e, s=0)).derivative()
        # PPoly repeats the end knots: keep the pieces of non-zero width
        keep = np.diff(pp.x) > 0
        if breaks is None:
            breaks = pp.x[:-1][keep]
        out.append(pp.c[:, keep])           # [3, pieces], highest power first
    table = np.transpose(np.array(out), (2, 0, 1))   # [pieces, k, 3]
    return np.ascontiguousarray(breaks), np.ascontiguousarray(table)


class HondaTable2D:
    """Device-ready tables of one flux file (all four primaries), see the module docstring."""

    def __init__(self, flux_file, enpow=1):
        self.spline_dict = load_2d_table(flux_file, enpow=enpow)
        self.enpow = enpow
        czkeys = ["%.2f" % x for x in np.linspace(-0.95, 0.95, N_CZ)]    # flux_weights.py:335
        t = None
        wrk = np.empty((4, N_CZ, 0))
        rows = []
        for prim in OUT_ORDER:
            for key in czkeys:
                tt, c, k = self.spline_dict[prim][key]
                assert k == 3
                if t is None:
                    t = np.ascontiguousarray(tt, dtype=np.float64)
                elif not np.array_equal(t, tt):
                    raise ValueError("the energy splines of a table must share one knot vector")
                n = len(tt)
                nk1 = n - 4
                w = np.array(c[:nk1], dtype=np.float64)
                # FITPACK splder, nu = 1: wrk(i) = k (wrk(i+1) - wrk(i)) / (t(i+k+1) - t(i+1))
                d = np.empty(nk1 - 1)
                for i in range(nk1 - 1):
                    fac = t[i + 4] - t[i + 1]
                    d[i] = 3.0 * (w[i + 1] - w[i]) / fac if fac > 0 else w[i]
                rows.append(d)
        self.knots = t                                           # [n]
        # layout [coef index][primary * 20 + coszen row]: the 80 values needed for one basis function are contiguous
        self.dcoef = np.ascontiguousarray(np.array(rows).T)      # [n - 5, 80]
        self.cz_breaks, self.cz_table = _cardinal_derivative_table()
        self.cells = self._cell_polynomials()
        self._dev = {}

    def _cell_polynomials(self):
        """Both steps are piecewise quadratic -- in s = log10(E) - t_l on knot interval l and in u = coszen - break_p
        on coszen piece p -- so on every (l, p) cell the flux of a primary is ONE biquadratic
        sum_ab K[l, p, primary, a, b] s^a u^b.  K is expanded here once (in extended precision: the cardinal-spline
        weights alternate in sign) from the same ``dcoef`` / ``cz_table`` the step-by-step evaluation uses:
        the three non-zero quadratic B-splines of fpbspl as polynomials in s, the cumulative sum over the coszen
        rows, the 0.1 bin width and the cardinal derivative polynomials."""
        LD = np.longdouble
        t = self.knots.astype(LD)
        n = len(t)
        nk1 = n - 4
        n_int = nk1 - 3                                # FITPACK intervals l = 4 .. nk1 (1-based)
        dcoef = self.dcoef.astype(LD).reshape(-1, 4, N_CZ)
        alpha = np.zeros((n_int, 3, 3), dtype=LD)      # [interval][basis m][power of s]
        for L in range(n_int):
            l = L + 4
            tl, tl1, tlm, tl2 = t[l - 1], t[l], t[l - 2], t[l + 1]
            d, a, b = tl1 - tl, tl - tlm, tl2 - tl
            alpha[L, 0] = np.array([d * d, -2 * d, 1], dtype=LD) / (d * (d + a))
            alpha[L, 2] = np.array([0, 0, 1], dtype=LD) / (d * b)
            alpha[L, 1] = (np.array([d * a, d - a, -1], dtype=LD) / (d * (d + a))
                           + np.array([0, b, -1], dtype=LD) / (d * b))
        # v[L, prim, j, a] = sum_m dcoef[L + m, prim, j] alpha[L, m, a]; cumulative over the coszen rows j
        v = sum(dcoef[m:m + n_int, :, :, None] * alpha[:, m, None, None, :] for m in range(3))
        acc = np.cumsum(v, axis=2)
        dz = self.cz_table.astype(LD)[:, 1:, ::-1]     # [piece][k = 1..20][power of u]
        cells = LD(0.1) * np.einsum("lqja,pjb->lpqab", acc, dz)
        return np.ascontiguousarray(cells.astype(np.float64))   # [n_int, pieces, 4, 3, 3]

    def evaluate_cells(self, true_energy, true_coszen):
        """Host restatement of the device evaluation (used by the CPU tests): [n, 4] in OUT_ORDER."""
        x = np.log10(np.asarray(true_energy, dtype=np.float64))
        cz = np.asarray(true_coszen, dtype=np.float64)
        t = self.knots
        nk1 = len(t) - 4
        l = np.clip(np.searchsorted(t, x, side="right"), 4, nk1)       # t(l) <= x < t(l+1), clamped like splder
        s = x - t[l - 1]
        p = np.clip(np.searchsorted(self.cz_breaks, cz, side="right") - 1, 0, len(self.cz_breaks) - 1)
        u = cz - self.cz_breaks[p]
        K = self.cells[l - 4, p]                                        # [n, 4, 3, 3]
        sp = np.stack([np.ones_like(s), s, s * s], axis=1)
        up = np.stack([np.ones_like(u), u, u * u], axis=1)
        return np.einsum("nqab,na,nb->nq", K, sp, up) / np.power(np.asarray(true_energy, dtype=np.float64),
                                                                  self.enpow)[:, None]

    def device_tables(self, device):
        import torch
        key = str(device)
        if key not in self._dev:
            # device layout of the cells: [interval][piece][a][b][primary] (one 32-byte group per monomial)
            cells_dev = np.ascontiguousarray(np.transpose(self.cells, (0, 1, 3, 4, 2)))
            self._dev[key] = tuple(torch.tensor(a, dtype=torch.float64, device=device)
                                   for a in (self.knots, self.cz_breaks, cells_dev))
        return self._dev[key]
