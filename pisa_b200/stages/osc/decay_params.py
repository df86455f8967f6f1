"""Neutrino-decay parameter holder of ``osc.prob3`` (same attribute names as the reference's ``DecayParams``,
pisa/stages/osc/decay_params.py, so ``prob3`` reads like the reference's stage).

Model: invisible decay of the third mass eigenstate, ``alpha3 = m3 / tau3`` in eV^2.  The kernels take it as the
3x3 complex ``mat_decay`` in the MASS basis (``pisab_osc_consts_t.mat_decay``); the rotation to the flavour basis
(``U mat_decay U^dagger``, numba_osc_kernels.py:571-603) happens in ``build_decay_table`` (csrc/tables.cu)."""
import numpy as np

__all__ = ["DecayParams"]


class DecayParams:
    __slots__ = ("decay_alpha3",)

    def __init__(self, decay_alpha3=0.0):
        self.decay_alpha3 = decay_alpha3   # eV^2

    @property
    def decay_matrix(self):
        """``diag(0, 0, -i alpha3)`` as complex128 (decay_params.py:47-55)."""
        out = np.zeros((3, 3), dtype=np.complex128)
        out[2, 2] = complex(0.0, -float(self.decay_alpha3))
        return out
