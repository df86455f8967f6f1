"""``DecayParams``: the neutrino-decay parameter holder of ``osc.prob3`` (reference: pisa/stages/osc/decay_params.py).

The model is the invisible decay of the third mass eigenstate, parameterised by ``alpha3 = m3 / tau3`` in eV^2; the
kernels receive it as the 3x3 complex ``mat_decay`` in the mass basis (decay_params.py:47-55)."""
import numpy as np

__all__ = ["DecayParams"]


class DecayParams:
    def __init__(self):
        self._decay_alpha3 = 0.0

    @property
    def decay_alpha3(self):
        """alpha3 [eV^2]"""
        return self._decay_alpha3

    @decay_alpha3.setter
    def decay_alpha3(self, value):
        self._decay_alpha3 = value

    @property
    def decay_matrix(self):
        """diag(0, 0, -i alpha3)"""
        m = np.zeros((3, 3), dtype=np.complex128)
        m[2, 2] = 0 - self.decay_alpha3 * 1j
        return m
