"""Host-side oscillation parameters -> the 3x3 matrices the propagation kernels take.

Same attribute API as the reference's ``OscParams`` (pisa/stages/osc/osc_params.py:72-292):
angles are stored as their sines (so theta > 90 deg cannot be represented, :104,182-184),
``mix_matrix_complex`` is the PDG-parameterised PMNS matrix (:175-211), ``mix_matrix_reparam_complex``
its rephased variant (:214-263) and ``dm_matrix`` the antisymmetric matrix of squared-mass
differences including the degeneracy nudges (:266-292).  Always float64: these are kernel
*parameters*, not event data.
"""
import numpy as np

__all__ = ["OscParams"]


class OscParams:
    def __init__(self):
        self._sin12 = 0.0
        self._sin13 = 0.0
        self._sin23 = 0.0
        self._deltacp = 0.0
        self.dm21 = 0.0
        self.dm31 = 0.0

    @staticmethod
    def _checked_sine(value):
        if not abs(value) <= 1:
            raise AssertionError("sine of a mixing angle must be in [-1, 1]")
        return value

    sin12 = property(lambda self: self._sin12, lambda self, v: setattr(self, "_sin12", self._checked_sine(v)))
    sin13 = property(lambda self: self._sin13, lambda self, v: setattr(self, "_sin13", self._checked_sine(v)))
    sin23 = property(lambda self: self._sin23, lambda self, v: setattr(self, "_sin23", self._checked_sine(v)))

    theta12 = property(lambda self: np.arcsin(self._sin12), lambda self, v: setattr(self, "sin12", np.sin(v)))
    theta13 = property(lambda self: np.arcsin(self._sin13), lambda self, v: setattr(self, "sin13", np.sin(v)))
    theta23 = property(lambda self: np.arcsin(self._sin23), lambda self, v: setattr(self, "sin23", np.sin(v)))

    @property
    def deltacp(self):
        return self._deltacp

    @deltacp.setter
    def deltacp(self, value):
        if not (value >= 0.0 and value <= 2 * np.pi):
            raise AssertionError("deltacp must be within [0, 2pi]")
        self._deltacp = value

    def _trig(self):
        s12, s13, s23 = self._sin12, self._sin13, self._sin23
        return (s12, s13, s23, np.sqrt(1.0 - s12 ** 2), np.sqrt(1.0 - s13 ** 2), np.sqrt(1.0 - s23 ** 2),
                np.sin(self._deltacp), np.cos(self._deltacp))

    @property
    def mix_matrix_complex(self):
        """PMNS matrix, standard parameterisation (osc_params.py:175-211)."""
        s12, s13, s23, c12, c13, c23, sd, cd = self._trig()
        u = np.zeros((3, 3), dtype=np.complex128)
        u[0, 0] = c12 * c13
        u[0, 1] = s12 * c13
        u[0, 2] = complex(s13 * cd, -s13 * sd)
        u[1, 0] = complex(-s12 * c23 - c12 * s23 * s13 * cd, -c12 * s23 * s13 * sd)
        u[1, 1] = complex(c12 * c23 - s12 * s23 * s13 * cd, -s12 * s23 * s13 * sd)
        u[1, 2] = s23 * c13
        u[2, 0] = complex(s12 * s23 - c12 * c23 * s13 * cd, -c12 * c23 * s13 * sd)
        u[2, 1] = complex(-c12 * s23 - s12 * c23 * s13 * cd, -s12 * c23 * s13 * sd)
        u[2, 2] = c23 * c13
        return u

    @property
    def mix_matrix_reparam_complex(self):
        """diag(e^{i delta},1,1) U diag(e^{-i delta},1,1) (osc_params.py:214-263)."""
        s12, s13, s23, c12, c13, c23, sd, cd = self._trig()
        u = np.zeros((3, 3), dtype=np.complex128)
        u[0, 0] = c12 * c13
        u[0, 1] = complex(s12 * c13 * cd, s12 * c13 * sd)
        u[0, 2] = s13
        u[1, 0] = complex(-s12 * c23 * cd - c12 * s23 * s13, s12 * c23 * sd)
        u[1, 1] = complex(c12 * c23 - s12 * s23 * s13 * cd, -s12 * s23 * s13 * sd)
        u[1, 2] = s23 * c13
        u[2, 0] = complex(s12 * s23 * cd - c12 * c23 * s13, -s12 * s23 * sd)
        u[2, 1] = complex(-c12 * s23 - s12 * c23 * s13 * cd, -s12 * c23 * s13 * sd)
        u[2, 2] = c23 * c13
        return u

    @property
    def dm_matrix(self):
        """dm[i, j] = m_i^2 - m_j^2 with m_1^2 := 0 (osc_params.py:266-292)."""
        m = np.array([0.0, self.dm21, self.dm31], dtype=np.float64)
        delta = 5.0e-9
        if m[1] == 0.0:
            m[0] -= delta
        if m[2] == 0.0:
            m[2] += delta
        dm = m[:, None] - m[None, :]
        dm[np.diag_indices(3)] = 0.0
        return dm
