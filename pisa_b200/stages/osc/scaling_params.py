"""Earth-tomography density scalings (pisa/stages/osc/scaling_params.py).

``Mass_scaling``: one factor for every layer.  ``Core_scaling_w_constrain``: the core factor is free and the
factors of the inner and middle mantle follow from keeping the Earth's mass and moment of inertia fixed in a
hard-coded 5-layer model (:52-95).  ``Core_scaling_wo_constrain``: three independent factors (:98-139).
Arrays are ordered like ``Layers.rhos`` without the atmosphere: surface first.
"""
import numpy as np

__all__ = ["Mass_scaling", "Core_scaling_w_constrain", "Core_scaling_wo_constrain", "FIVE_LAYER_RADII",
           "FIVE_LAYER_RHOS", "TOMOGRAPHY_ERROR_MSG", "TOMOGRAPHY_TYPES"]

TOMOGRAPHY_TYPES = ["mass_of_earth", "mass_of_core_w_constrain", "mass_of_core_wo_constrain"]
FIVE_LAYER_RADII = np.array([0.0, 1221.50, 3480.00, 5701.00, 6151.0, 6371.00])   # km
FIVE_LAYER_RHOS = np.array([13.0, 13.0, 10.96, 5.03, 3.7, 2.5])                  # g/cm^3
TOMOGRAPHY_ERROR_MSG = ("You need to provide the appropriate 5-layer Earth model, which has the same layer radii "
                        "(%s km) and densities (%s g/cm^3) as the one hard-coded for the chosen type of tomography "
                        "internally." % (FIVE_LAYER_RADII.tolist(), FIVE_LAYER_RHOS.tolist()))


class Mass_scaling:  # pylint: disable=invalid-name
    def __init__(self):
        self._density_scale = 0.0

    @property
    def density_scale(self):
        return self._density_scale

    @density_scale.setter
    def density_scale(self, value):
        if not value >= 0.0:
            raise AssertionError("density_scale must not be negative")
        self._density_scale = value


def _shell_moments(power):
    """rho_i (r_i^p - r_{i-1}^p) per shell of the 5-layer model, p = 3 (mass) or 5 (moment of inertia),
    up to the common factors 4 pi / 3 and 8 pi / 15."""
    r, rho = FIVE_LAYER_RADII, FIVE_LAYER_RHOS
    return rho[1:] * (r[1:] ** power - r[:-1] ** power)


class Core_scaling_w_constrain:  # pylint: disable=invalid-name
    def __init__(self):
        self.core_density_scale = 0.0

    @property
    def scaling_array(self):
        (a1, b1, c1, d1, e1), (a2, b2, c2, d2, e2) = (4 * np.pi / 3) * _shell_moments(3), \
                                                     (8 * np.pi / 15) * _shell_moments(5)
        mass, inertia = a1 + b1 + c1 + d1 + e1, a2 + b2 + c2 + d2 + e2
        alpha = self.core_density_scale
        gamma = ((inertia * c1 - mass * c2) - alpha * (c1 * a2 - c2 * a1) - alpha * (c1 * b2 - b1 * c2)
                 - (c1 * e2 - e1 * c2)) / (c1 * d2 - d1 * c2)
        beta = (inertia - alpha * a2 - alpha * b2 - gamma * d2 - e2) / c2
        if not (np.array([alpha, beta, gamma]) >= 0).all():
            raise AssertionError("density scale factors must not be negative")
        out = np.ones(6, dtype=np.float64)
        out[1], out[2], out[3:] = gamma, beta, alpha
        return out


class Core_scaling_wo_constrain:  # pylint: disable=invalid-name
    def __init__(self):
        self.core_density_scale = 0.0
        self.innermantle_density_scale = 0.0
        self.middlemantle_density_scale = 0.0

    @property
    def scaling_factor_array(self):
        out = np.ones(6, dtype=np.float64)
        out[1], out[2], out[3:] = self.middlemantle_density_scale, self.innermantle_density_scale, \
            self.core_density_scale
        return out
