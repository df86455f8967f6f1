"""Long-range-interaction potential of an anomaly-free L_a - L_b symmetry (pisa/stages/osc/lri_params.py:24-116).

One parameter, the potential ``v_lri`` [eV]; the flavour-basis matrix is ``v_lri * (e_a e_a^T - e_b e_b^T)``.
The propagation kernel takes it as the ``lri_pot`` input of ``propagate_array`` (numba_osc_kernels.py:435-440).
"""
import numpy as np

__all__ = ["LRIParams", "LRI_TYPES"]

LRI_TYPES = ["emu-symmetry", "etau-symmetry", "mutau-symmetry"]
_CHARGES = {"emu-symmetry": (0, 1), "etau-symmetry": (0, 2), "mutau-symmetry": (1, 2)}


class LRIParams:
    def __init__(self):
        self._v_lri = 0.0

    @property
    def v_lri(self):
        return self._v_lri

    @v_lri.setter
    def v_lri(self, value):
        if not value < 1.0:
            raise AssertionError("v_lri must be below 1 eV")
        self._v_lri = value

    def potential_matrix(self, lri_type):
        if lri_type not in _CHARGES:
            raise ValueError("Implemented symmetries are %s" % LRI_TYPES)
        plus, minus = _CHARGES[lri_type]
        v = np.zeros((3, 3), dtype=np.float64)
        v[plus, plus] = self.v_lri
        v[minus, minus] = -self.v_lri
        return v

    potential_matrix_emu = property(lambda self: self.potential_matrix("emu-symmetry"))
    potential_matrix_etau = property(lambda self: self.potential_matrix("etau-symmetry"))
    potential_matrix_mutau = property(lambda self: self.potential_matrix("mutau-symmetry"))
