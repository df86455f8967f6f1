"""Earth model -> layer table for the propagation kernels (host side), and ``calcLayers`` on device.

Mirrors ``Layers`` (pisa/stages/osc/layers.py:172-481): PREM rows reversed (surface first) plus an
atmosphere shell at ``r_earth + prop_height`` (:226-241), ``max_layers = 2 * len(radii)`` (:244),
tangent directions ``coszen_limit`` (:308-335), electron-fraction weighting with the hard-coded
region radii 1221.5 / 3480 / 6371 km (:411-439) and density scaling (:291-306).  The per-coszen
geometry (``extCalcLayers`` :38-169) runs in the CUDA library -- inside the propagation kernels, or
through ``calcLayers`` when the [N, max_layers] arrays are wanted.
"""
import numpy as np

from pisa_b200 import FTYPE

__all__ = ["Layers"]


class Layers:
    R_INNER, R_OUTER, R_MANTLE = 1221.5, 3480.0, 6371.0

    def __init__(self, prem_file, detector_depth=1.0, prop_height=2.0):
        if prem_file is None:
            raise ValueError("pisa_b200 needs an Earth model file (vacuum-only Layers is out of scope)")
        self.using_earth_model = True
        self.prem = np.loadtxt(prem_file) if isinstance(prem_file, str) else np.asarray(prem_file, dtype=np.float64)
        r_earth = self.prem[-1][0]
        self._prem_rhos = self.prem[..., 1][::-1].astype(FTYPE)
        self.rhos_unweighted = np.concatenate((np.ones(1, dtype=FTYPE), self._prem_rhos))
        self.rhos = self.rhos_unweighted.copy()
        self.radii = np.concatenate((np.array([r_earth + prop_height]), self.prem[..., 0][::-1].astype(FTYPE)))
        self.max_layers = 2 * len(self.radii)
        if not detector_depth > 0:
            raise AssertionError("ERROR: detector depth must be a positive value")
        if not detector_depth <= r_earth:
            raise AssertionError("ERROR: detector depth is deeper than one Earth radius!")
        if not prop_height >= 0:
            raise AssertionError("ERROR: neutrino production height must be positive")
        self.r_detector = r_earth - detector_depth
        self.prop_height = prop_height
        self.detector_depth = detector_depth
        self.default_elec_frac = 0.5
        self.YeFrac = None
        self.computeMinLengthToLayers()
        self._n_layers = self._density = self._distance = None

    def computeMinLengthToLayers(self):
        """cos(zenith) at which a track is tangent to each shell (layers.py:308-335)."""
        lim = [1.0 if rad >= self.r_detector else -np.sqrt(1 - (rad ** 2 / self.r_detector ** 2))
               for rad in self.radii]
        self.coszen_limit = np.array(lim, dtype=FTYPE)

    def setElecFrac(self, YeI, YeO, YeM):
        """layers.py:270-289 + weight_density_to_YeFrac :411-439 (always from the unweighted densities)."""
        self.YeFrac = np.array([YeI, YeO, YeM], dtype=FTYPE)
        r, rho = self.radii, self.rhos_unweighted
        inner = rho * self.YeFrac[0] * (r <= self.R_INNER)
        outer = rho * self.YeFrac[1] * (r <= self.R_OUTER) * (r > self.R_INNER)
        mantle = rho * self.YeFrac[2] * (r <= self.R_MANTLE) * (r > self.R_OUTER)
        self.rhos = (inner + outer + mantle).astype(FTYPE)

    def scaling(self, scaling_array):
        """Multiply the Earth-model densities by per-layer factors (layers.py:291-306).
        Like the reference this rebuilds ``rhos`` from the file densities; ``setElecFrac`` (which
        works from ``rhos_unweighted``) must be re-applied by the caller (prob3.py:524-533)."""
        rhos = self._prem_rhos.copy()
        if scaling_array is not None:
            rhos = rhos * scaling_array
        self.rhos = np.concatenate((np.ones(1, dtype=FTYPE), rhos))

    # --- device side -----------------------------------------------------------------------
    def earth_struct(self):
        """The ``pisab_earth_t`` the kernels consume."""
        from pisa_b200 import ops
        return ops.Earth.from_arrays(self.radii, self.rhos, self.coszen_limit, self.r_detector, self.max_layers)

    def calcLayers(self, cz):
        """``cz``: CUDA tensor.  Sets ``n_layers``, ``density``, ``distance`` (device tensors shaped
        like the reference's flattened arrays reshaped to [N, max_layers], layers.py:339-363)."""
        from pisa_b200 import ops
        self._n_layers, self._density, self._distance = ops.layers_calc(self.earth_struct(), cz)

    n_layers = property(lambda self: self._n_layers)
    density = property(lambda self: self._density)
    distance = property(lambda self: self._distance)
