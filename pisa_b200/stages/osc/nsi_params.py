"""Standard-parameterisation NSI couplings -> Hermitian epsilon matrix.

Mirrors ``StdNSIParams`` (pisa/stages/osc/nsi_params.py:77-181): diagonal couplings are real,
off-diagonal ones are set from (magnitude, phase) tuples, and ``eps_matrix`` subtracts the mumu
entry from the diagonal (:167-181).  ``VacuumLikeNSIParams`` (:184-384) builds the same matrix from a
"matter mixing matrix": two eigenvalues, three rotation angles and three phases.
"""
import numpy as np

__all__ = ["StdNSIParams", "VacuumLikeNSIParams"]


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


def _bounded(name, lo, hi):
    """Property whose setter asserts lo <= value <= hi (the reference's setters :227-292)."""
    attr = "_" + name

    def setter(self, value):
        if not lo <= value <= hi:
            raise AssertionError("%s must lie within [%g, %g]" % (name, lo, hi))
        setattr(self, attr, value)

    return property(lambda self: getattr(self, attr), setter)


def _real_scalar(name):
    attr = "_" + name

    def setter(self, value):
        if isinstance(value, complex) or not np.isscalar(value):
            raise TypeError("%s must be a real number!" % name)
        setattr(self, attr, value)

    return property(lambda self: getattr(self, attr), setter)


class VacuumLikeNSIParams:
    """Generalised matter potential  Q U diag(eps_scale, eps_prime, 0) U^+ Q^+  with
    U = R12(phi12) R13(phi13) R23(phi23, deltansi) and Q = diag(e^{i a1}, e^{i a2}, e^{-i(a1+a2)});
    the NSI coupling matrix is that potential minus its mumu entry on the diagonal and minus the standard
    CC term in ee (nsi_params.py:326-384)."""

    eps_scale = _real_scalar("eps_scale")
    eps_prime = _real_scalar("eps_prime")
    phi12 = _bounded("phi12", -np.pi, np.pi)
    phi13 = _bounded("phi13", -np.pi, np.pi)
    phi23 = _bounded("phi23", -np.pi, np.pi)
    alpha1 = _bounded("alpha1", 0.0, 2 * np.pi)
    alpha2 = _bounded("alpha2", 0.0, 2 * np.pi)
    deltansi = _bounded("deltansi", 0.0, 2 * np.pi)

    def __init__(self):
        self._eps_scale, self._eps_prime = 1.0, 0.0
        self._phi12 = self._phi13 = self._phi23 = 0.0
        self._alpha1 = self._alpha2 = self._deltansi = 0.0

    @staticmethod
    def _phase(angle):
        return complex(np.cos(angle), np.sin(angle))

    @property
    def eps_matrix(self):
        c12, s12 = np.cos(self.phi12), np.sin(self.phi12)
        c13, s13 = np.cos(self.phi13), np.sin(self.phi13)
        c23, s23 = np.cos(self.phi23), np.sin(self.phi23)
        r12 = np.array([[c12, s12, 0], [-s12, c12, 0], [0, 0, 1]], dtype=np.float64)
        r13 = np.array([[c13, 0, s13], [0, 1, 0], [-s13, 0, c13]], dtype=np.float64)
        r23 = np.array([[1, 0, 0],
                        [0, c23, s23 * self._phase(-self.deltansi)],
                        [0, -s23 * self._phase(self.deltansi), c23]])
        q = np.diag([self._phase(self.alpha1), self._phase(self.alpha2), self._phase(-(self.alpha1 + self.alpha2))])
        d = np.diag(np.array([self.eps_scale, self.eps_prime, 0], dtype=np.float64))
        u = r12 @ (r13 @ r23)
        # innermost product first, like the reference (rounding-identical association)
        pot = q @ (u @ (d @ (u.conj().T @ q.conj().T)))
        pot = pot - pot[1, 1] * np.eye(3)
        pot[0, 0] = pot[0, 0] - 1.0
        for i in range(3):
            pot[i, i] = pot[i, i].real
        if not np.allclose(pot, pot.conj().T, rtol=1e-12, atol=np.finfo(np.float64).eps):
            raise AssertionError("NSI coupling matrix is not Hermitian")
        return pot

    eps_ee = property(lambda s: s.eps_matrix[0, 0].real)
    eps_mumu = property(lambda s: s.eps_matrix[1, 1].real)
    eps_tautau = property(lambda s: s.eps_matrix[2, 2].real)
    eps_emu = property(lambda s: s.eps_matrix[0, 1])
    eps_etau = property(lambda s: s.eps_matrix[0, 2])
    eps_mutau = property(lambda s: s.eps_matrix[1, 2])
