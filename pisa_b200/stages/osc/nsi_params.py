"""Standard-parameterisation NSI couplings -> Hermitian epsilon matrix.

Mirrors ``StdNSIParams`` (pisa/stages/osc/nsi_params.py:77-181): diagonal couplings are real,
off-diagonal ones are set from (magnitude, phase) tuples, and ``eps_matrix`` subtracts the mumu
entry from the diagonal (:167-181).  The vacuum-like parameterisation is out of scope.
"""
import numpy as np

__all__ = ["StdNSIParams"]


def _magnitude_phase(value):
    try:
        magnitude, phase = value
    except TypeError:
        raise TypeError("off-diagonal NSI couplings are set from a (magnitude, phase) pair")
    return float(magnitude), float(phase)


class StdNSIParams:
    def __init__(self):
        self._eps = np.zeros((3, 3), dtype=np.complex128)

    def _set_diag(self, i, value, name):
        if isinstance(value, complex) or not np.isscalar(value):
            raise TypeError("%s must be a real number!" % name)
        self._eps[i, i] = value + 1.0j * self._eps[i, i].imag

    def _set_offdiag(self, i, j, value):
        magnitude, phase = _magnitude_phase(value)
        self._eps[i, j] = magnitude * (np.cos(phase) + 1.0j * np.sin(phase))
        self._eps[j, i] = np.conjugate(self._eps[i, j])

    eps_ee = property(lambda s: s.eps_matrix[0, 0].real, lambda s, v: s._set_diag(0, v, "eps_ee"))
    eps_mumu = property(lambda s: s.eps_matrix[1, 1].real, lambda s, v: s._set_diag(1, v, "eps_mumu"))
    eps_tautau = property(lambda s: s.eps_matrix[2, 2].real, lambda s, v: s._set_diag(2, v, "eps_tautau"))
    eps_emu = property(lambda s: s.eps_matrix[0, 1], lambda s, v: s._set_offdiag(0, 1, v))
    eps_etau = property(lambda s: s.eps_matrix[0, 2], lambda s, v: s._set_offdiag(0, 2, v))
    eps_mutau = property(lambda s: s.eps_matrix[1, 2], lambda s, v: s._set_offdiag(1, 2, v))

    @property
    def eps_matrix(self):
        eps = self._eps - self._eps[1, 1] * np.eye(3)
        for i in range(3):
            eps[i, i] = eps[i, i].real
        if not np.allclose(eps, eps.conj().T, rtol=1e-12, atol=np.finfo(np.float64).eps):
            raise AssertionError("NSI coupling matrix is not Hermitian")
        return eps
