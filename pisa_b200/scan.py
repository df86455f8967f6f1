"""Fit-loop / parameter-scan driver (BASELINE.json configs[4]).

What the reference does per hypothesis inside ``Analysis.scan`` / a minimiser iteration
(pisa/analysis/analysis.py -> ``DistributionMaker.get_outputs`` -> ``Pipeline.run`` of every stage ->
``MapSet`` sum -> ``mod_chi2``, pisa/utils/stats.py:651-695) costs >= 40 ms of Python/`Map` overhead even
when every stage cache hits.  Here one hypothesis is: build the 3 host matrices
(microseconds), ONE fused launch over all resident flavour containers (``ReweightEngine.evaluate``), the
histogram exchange when sharded over GPUs, and one tiny kernel that sums the containers and evaluates
``mod_chi2`` into its slot of a device array.  Nothing synchronises inside the loop; the chi2 values are
read back once at the end.
"""
import math

import numpy as np
import torch

from . import ops

__all__ = ["osc_consts", "osc_consts_array", "scan_chi2", "asimov", "fit_chi2", "OSC_PARAM_NAMES", "DECAY_PARAM_NAME"]

OSC_PARAM_NAMES = ("theta12", "theta13", "theta23", "deltacp", "dm21", "dm31")
# optional seventh parameter of the drivers below: invisible decay of the third mass state (prob3 neutrino_decay=True,
# prob3.py:76-88; decay_params.py:47-55: mat_decay = diag(0, 0, -i alpha3), decay_flag = 1)
DECAY_PARAM_NAME = "decay_alpha3"


def osc_consts(theta12, theta13, theta23, deltacp, dm21, dm31, mat_pot=None, decay_alpha3=None):
    """OscConsts from mixing angles (rad) and mass splittings (eV^2), like prob3.compute_function builds
    them (prob3.py:485-559); ``mat_pot`` defaults to the standard matter potential diag(1, 0, 0).
    Scalar arithmetic with the formulas of ``OscParams`` (sines via numpy like its setters, cosines as
    sqrt(1 - s^2); osc_params.py:175-211,266-292), bit-identical to it and ~10x cheaper than going through
    the matrix properties: this sits in the inner loop of a sequential fit."""
    if not 0.0 <= deltacp <= 2 * np.pi:
        raise AssertionError("deltacp must be within [0, 2pi]")
    s12, s13, s23 = float(np.sin(theta12)), float(np.sin(theta13)), float(np.sin(theta23))
    c12, c13, c23 = math.sqrt(1.0 - s12 ** 2), math.sqrt(1.0 - s13 ** 2), math.sqrt(1.0 - s23 ** 2)
    sd, cd = float(np.sin(deltacp)), float(np.cos(deltacp))
    c = ops.OscConsts()
    m0, m1, m2 = 0.0, float(dm21), float(dm31)
    if m1 == 0.0:
        m0 -= 5.0e-9
    if m2 == 0.0:
        m2 += 5.0e-9
    c.dm[:] = (0.0, m0 - m1, m0 - m2, m1 - m0, 0.0, m1 - m2, m2 - m0, m2 - m1, 0.0)
    c.mix[:] = (c12 * c13, 0.0, s12 * c13, 0.0, s13 * cd, -s13 * sd,
                -s12 * c23 - c12 * s23 * s13 * cd, -c12 * s23 * s13 * sd,
                c12 * c23 - s12 * s23 * s13 * cd, -s12 * s23 * s13 * sd, s23 * c13, 0.0,
                s12 * s23 - c12 * c23 * s13 * cd, -c12 * c23 * s13 * sd,
                -c12 * s23 - s12 * c23 * s13 * cd, -s12 * c23 * s13 * sd, c23 * c13, 0.0)
    if mat_pot is None:
        c.mat_pot[0] = 1.0
    else:
        mp = np.asarray(mat_pot, dtype=np.complex128).reshape(3, 3)
        c.mat_pot[:] = np.stack([mp.real, mp.imag], axis=-1).ravel()
    c.decay_flag = -1
    if decay_alpha3 is not None:
        c.decay_flag = 1
        c.mat_decay[17] = -float(decay_alpha3)   # Im of element [2][2]
    return c


def osc_consts_array(theta12, theta13, theta23, deltacp, dm21, dm31, mat_pot=None, decay_alpha3=None):
    """Vectorised ``osc_consts``: any of the six parameters may be an array of P values (the others
    broadcast); returns a ctypes array ``OscConsts[P]`` for ``ops.reweight_hist_scan``.  Same arithmetic as
    ``OscParams`` (sines stored, cosines as sqrt(1 - s^2), osc_params.py:175-211,266-292); building the
    hypotheses one by one costs ~50 us each in Python, which would dominate a scan over an analysis-size
    sample.  ``decay_alpha3`` (scalar or array, eV^2; None = no decay) selects the decay branch for every hypothesis."""
    cols = [theta12, theta13, theta23, deltacp, dm21, dm31] + ([decay_alpha3] if decay_alpha3 is not None else [])
    cols = np.broadcast_arrays(*[np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in cols])
    t12, t13, t23, dcp, m21, m31 = cols[:6]
    n = t12.shape[0]
    if not np.all((dcp >= 0.0) & (dcp <= 2 * np.pi)):
        raise AssertionError("deltacp must be within [0, 2pi]")
    s12, s13, s23 = np.sin(t12), np.sin(t13), np.sin(t23)
    c12, c13, c23 = np.sqrt(1.0 - s12 ** 2), np.sqrt(1.0 - s13 ** 2), np.sqrt(1.0 - s23 ** 2)
    sd, cd = np.sin(dcp), np.cos(dcp)
    rec = np.zeros((n, 73), dtype=np.float64)   # dm[9] mix[18] mat_pot[18] mat_decay[18] lri_pot[9] decay_flag
    # dm matrix (osc_params.py:266-292)
    m = np.stack([np.zeros(n), m21, m31], axis=1)
    m[:, 0] -= np.where(m[:, 1] == 0.0, 5.0e-9, 0.0)
    m[:, 2] += np.where(m[:, 2] == 0.0, 5.0e-9, 0.0)
    dm = m[:, :, None] - m[:, None, :]
    dm[:, [0, 1, 2], [0, 1, 2]] = 0.0
    rec[:, 0:9] = dm.reshape(n, 9)
    # PMNS, standard parameterisation (re, im interleaved, row-major)
    u = np.zeros((n, 3, 3, 2))
    u[:, 0, 0, 0] = c12 * c13
    u[:, 0, 1, 0] = s12 * c13
    u[:, 0, 2, 0], u[:, 0, 2, 1] = s13 * cd, -s13 * sd
    u[:, 1, 0, 0], u[:, 1, 0, 1] = -s12 * c23 - c12 * s23 * s13 * cd, -c12 * s23 * s13 * sd
    u[:, 1, 1, 0], u[:, 1, 1, 1] = c12 * c23 - s12 * s23 * s13 * cd, -s12 * s23 * s13 * sd
    u[:, 1, 2, 0] = s23 * c13
    u[:, 2, 0, 0], u[:, 2, 0, 1] = s12 * s23 - c12 * c23 * s13 * cd, -c12 * c23 * s13 * sd
    u[:, 2, 1, 0], u[:, 2, 1, 1] = -c12 * s23 - s12 * c23 * s13 * cd, -s12 * c23 * s13 * sd
    u[:, 2, 2, 0] = c23 * c13
    rec[:, 9:27] = u.reshape(n, 18)
    if mat_pot is None:
        rec[:, 27] = 1.0
    else:
        mp = np.asarray(mat_pot, dtype=np.complex128).reshape(3, 3)
        rec[:, 27:45] = np.stack([mp.real, mp.imag], axis=-1).ravel()
    rec[:, 72] = np.array([-1], dtype=np.int64).view(np.float64)[0]   # decay_flag = -1
    if decay_alpha3 is not None:
        rec[:, 45 + 17] = -cols[6]                                     # mat_decay[2][2] = -i alpha3
        rec[:, 72] = np.array([1], dtype=np.int64).view(np.float64)[0]
    arr = (ops.OscConsts * n).from_buffer_copy(np.ascontiguousarray(rec).tobytes())
    return arr


def asimov(engine, consts):
    """Pseudo-data without fluctuations: the summed map of the template at ``consts`` ([n_bins])."""
    return engine.evaluate(consts)[:, 0].sum(dim=0).contiguous()


def scan_chi2(engine, observed, points, fixed, mat_pot=None, batch=64):
    """``mod_chi2`` of the template against ``observed`` at every (theta23, dm31) point.

    ``batch`` hypotheses are evaluated per kernel launch (``ReweightEngine.evaluate_many``): for the event
    samples of a real analysis (1e5 .. 1e6 events) one template cannot fill the GPU and a launch per template
    is bound by launch + host overhead; ``batch=1`` takes the one-launch-per-template path.

    engine   : ReweightEngine with resident containers (scales set through ``set_scales``)
    observed : float64 device tensor [n_bins]
    points   : iterable of (theta23 [rad], dm31 [eV^2])
    fixed    : dict theta12, theta13, deltacp [rad], dm21 [eV^2]; optionally decay_alpha3 [eV^2] (neutrino decay)
    Returns a float64 device tensor [n_points] (no host synchronisation happens here).
    """
    points = list(points)
    out = torch.empty(len(points), dtype=torch.float64, device=engine.device)
    alpha3 = fixed.get(DECAY_PARAM_NAME)
    if batch <= 1:
        for i, (t23, dm31) in enumerate(points):
            consts = osc_consts(fixed["theta12"], fixed["theta13"], t23, fixed["deltacp"], fixed["dm21"], dm31, mat_pot,
                                decay_alpha3=alpha3)
            engine.evaluate_chi2(consts, observed, chi2_out=out[i:i + 1])   # one call, two launches per hypothesis
        return out
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 2)
    for lo in range(0, len(points), batch):
        chunk = pts[lo:lo + batch]
        consts = osc_consts_array(fixed["theta12"], fixed["theta13"], chunk[:, 0], fixed["deltacp"], fixed["dm21"],
                                  chunk[:, 1], mat_pot, decay_alpha3=alpha3)
        hist = engine.evaluate_many(consts)
        ops.template_chi2_batch(hist, observed, out=out[lo:lo + len(chunk)])
    return out


def fit_chi2(engine, observed, start, fixed, bounds=None, steps=None, mat_pot=None, method="L-BFGS-B", options=None):
    """Minimise ``mod_chi2`` of the template against ``observed`` over the oscillation parameters in ``start``
    (what a minimiser iteration of ``Analysis.fit_hypo`` does through ``DistributionMaker.get_outputs``,
    pisa/analysis/analysis.py; here only the oscillation parameters of the hot path can float).

    Every objective call evaluates the point AND its central-difference gradient -- 2k + 1 hypotheses for k free
    parameters -- in ONE launch (``ReweightEngine.evaluate_many`` + ``template_chi2_batch``) and reads 2k + 1
    doubles back: for an analysis-size sample a launch is latency-bound, so the gradient costs nothing extra.

    start  : dict name -> start value for the free parameters (names from ``OSC_PARAM_NAMES``; rad / eV^2; and
             optionally ``decay_alpha3`` [eV^2]: neutrino decay, every hypothesis then runs the decay kernels)
    fixed  : dict with the remaining parameters (``decay_alpha3`` only when decay is wanted)
    bounds : optional dict name -> (lo, hi)
    steps  : optional dict name -> finite-difference half step (default 1e-4 of the start value's magnitude)
    Returns ``scipy.optimize.OptimizeResult`` with ``x`` as a dict, plus ``n_templates`` evaluated.
    """
    from scipy import optimize
    names = list(start)
    for nm in names + list(fixed):
        if nm not in OSC_PARAM_NAMES and nm != DECAY_PARAM_NAME:
            raise ValueError("unknown oscillation parameter %r (known: %s, %s)"
                             % (nm, ", ".join(OSC_PARAM_NAMES), DECAY_PARAM_NAME))
    with_decay = DECAY_PARAM_NAME in start or DECAY_PARAM_NAME in fixed
    missing = [nm for nm in OSC_PARAM_NAMES if nm not in start and nm not in fixed]
    if missing or set(start) & set(fixed):
        raise ValueError("every oscillation parameter must be either free or fixed (missing: %s)" % missing)
    k = len(names)
    x0 = np.array([float(start[nm]) for nm in names])
    # the minimiser works in units of `scale` so that angles (~1) and mass splittings (~1e-3) are comparable
    scale = np.where(x0 != 0.0, np.abs(x0), 1.0)
    h = np.array([float((steps or {}).get(nm, 1e-4 * s)) for nm, s in zip(names, scale)])
    # physical domains the host-side parameter code enforces (osc_params.py: deltacp in [0, 2 pi]): a stencil point or a
    # minimiser step outside would abort the fit with an AssertionError, so they act as default bounds
    domain = {"deltacp": (0.0, 2 * np.pi), DECAY_PARAM_NAME: (0.0, np.inf)}
    def _bound(nm, k):
        user = (bounds or {}).get(nm, (-np.inf, np.inf))[k]
        dom = domain.get(nm, (-np.inf, np.inf))[k]
        return max(user, dom) if k == 0 else min(user, dom)
    lo = np.array([_bound(nm, 0) for nm in names], dtype=np.float64)
    hi = np.array([_bound(nm, 1) for nm in names], dtype=np.float64)
    out = torch.empty(2 * k + 1, dtype=torch.float64, device=engine.device)
    counter = [0]

    def objective(u):
        x = np.clip(u * scale, lo, hi)
        pts = np.repeat(x[None, :], 2 * k + 1, axis=0)
        for i in range(k):
            # one-sided at a bound: keep both stencil points inside it
            up, dn = min(x[i] + h[i], hi[i]), max(x[i] - h[i], lo[i])
            pts[1 + 2 * i, i], pts[2 + 2 * i, i] = up, dn
        args = {nm: pts[:, i] for i, nm in enumerate(names)}
        args.update(fixed)
        consts = osc_consts_array(args["theta12"], args["theta13"], args["theta23"], args["deltacp"], args["dm21"],
                                  args["dm31"], mat_pot, decay_alpha3=args[DECAY_PARAM_NAME] if with_decay else None)
        ops.template_chi2_batch(engine.evaluate_many(consts), observed, out=out)
        counter[0] += 2 * k + 1
        c = out.cpu().numpy()
        grad = np.array([(c[1 + 2 * i] - c[2 + 2 * i]) / (pts[1 + 2 * i, i] - pts[2 + 2 * i, i]) for i in range(k)])
        return float(c[0]), grad * scale

    res = optimize.minimize(objective, x0 / scale, jac=True, method=method,
                            bounds=list(zip(lo / scale, hi / scale)) if np.isfinite(np.concatenate([lo, hi])).any() else None,
                            options=options)
    res.x = {nm: float(v) for nm, v in zip(names, res.x * scale)}
    res.n_templates = counter[0]
    return res
