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
import numpy as np
import torch

from . import ops
from .stages.osc.osc_params import OscParams

__all__ = ["osc_consts", "scan_chi2", "asimov"]


def osc_consts(theta12, theta13, theta23, deltacp, dm21, dm31, mat_pot=None):
    """OscConsts from mixing angles (rad) and mass splittings (eV^2), like prob3.compute_function builds
    them (prob3.py:485-559); ``mat_pot`` defaults to the standard matter potential diag(1, 0, 0)."""
    op = OscParams()
    op.theta12, op.theta13, op.theta23, op.deltacp = theta12, theta13, theta23, deltacp
    op.dm21, op.dm31 = dm21, dm31
    if mat_pot is None:
        mat_pot = np.zeros((3, 3), dtype=np.complex128)
        mat_pot[0, 0] = 1.0
    return ops.OscConsts.from_matrices(op.dm_matrix, op.mix_matrix_complex, mat_pot)


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
    fixed    : dict theta12, theta13, deltacp [rad], dm21 [eV^2]
    Returns a float64 device tensor [n_points] (no host synchronisation happens here).
    """
    points = list(points)
    out = torch.empty(len(points), dtype=torch.float64, device=engine.device)
    if batch <= 1:
        for i, (t23, dm31) in enumerate(points):
            consts = osc_consts(fixed["theta12"], fixed["theta13"], t23, fixed["deltacp"], fixed["dm21"], dm31, mat_pot)
            hist = engine.evaluate(consts)
            ops.template_chi2(hist, observed, out=out[i:i + 1])
        return out
    for lo in range(0, len(points), batch):
        chunk = points[lo:lo + batch]
        consts = [osc_consts(fixed["theta12"], fixed["theta13"], t23, fixed["deltacp"], fixed["dm21"], dm31, mat_pot)
                  for t23, dm31 in chunk]
        hist = engine.evaluate_many(consts)
        ops.template_chi2_batch(hist, observed, out=out[lo:lo + len(chunk)])
    return out
