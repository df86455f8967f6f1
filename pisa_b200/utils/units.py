"""A deliberately tiny stand-in for the ``pint`` unit registry the reference uses.

The reference reads every stage parameter through ``param.value.m_as('unit')``
(e.g. pisa/stages/osc/prob3.py:360-364,485-490) and writes quantities in .cfg files as
``42. * units.degree`` / ``2. units.km`` / ``7.5e-5 units.eV**2``
(pisa/utils/config_parser.py:303-354).  Only that surface is restated: a ``Quantity`` with
``.m`` / ``.magnitude`` / ``.units`` / ``.m_as()`` / ``.to()``, multiplication by a unit, and a
registry ``ureg`` that understands the units appearing in the hot path's configs.  Anything else
raises, it does not guess.
"""
import math
import re

import numpy as np

__all__ = ["Quantity", "Unit", "ureg", "parse_quantity"]

# unit -> (dimension, factor to the dimension's base unit)
_UNITS = {
    "dimensionless": ("", 1.0),
    "rad": ("angle", 1.0), "radian": ("angle", 1.0), "radians": ("angle", 1.0),
    "deg": ("angle", math.pi / 180.0), "degree": ("angle", math.pi / 180.0), "degrees": ("angle", math.pi / 180.0),
    "eV": ("energy", 1.0), "keV": ("energy", 1e3), "MeV": ("energy", 1e6), "GeV": ("energy", 1e9),
    "TeV": ("energy", 1e12), "electron_volt": ("energy", 1.0), "gigaelectron_volt": ("energy", 1e9),
    "eV**2": ("energy**2", 1.0), "electron_volt**2": ("energy**2", 1.0),
    "m": ("length", 1.0), "meter": ("length", 1.0), "km": ("length", 1e3), "kilometer": ("length", 1e3),
    "cm": ("length", 1e-2),
    "s": ("time", 1.0), "second": ("time", 1.0), "sec": ("time", 1.0),
    "common_year": ("time", 365.0 * 86400.0), "year": ("time", 365.25 * 86400.0),
    "g/cm**3": ("density", 1.0),
}


def _canon(unit):
    if unit is None:
        return "dimensionless"
    if isinstance(unit, Unit):
        return unit.name
    u = str(unit).strip().replace(" ", "")
    if u in ("", "1"):
        return "dimensionless"
    u = u.replace("^", "**")
    if u not in _UNITS:
        raise ValueError("unit %r is not known to pisa_b200.utils.units" % unit)
    return u


class Unit:
    __array_ufunc__ = None  # let `ndarray * unit` defer to Unit.__rmul__

    def __init__(self, name):
        self.name = _canon(name)

    @property
    def dimensionality(self):
        return _UNITS[self.name][0]

    def __pow__(self, p):
        return Unit("%s**%d" % (self.name, p)) if p != 1 else self

    def __rmul__(self, other):
        return Quantity(other, self)

    def __mul__(self, other):
        if isinstance(other, (int, float, np.ndarray, list, tuple)):
            return Quantity(other, self)
        return NotImplemented

    def __eq__(self, other):
        try:
            return _canon(other) == self.name
        except ValueError:
            return False

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return self.name


class Quantity:
    """magnitude * unit."""
    __array_priority__ = 1000
    __array_ufunc__ = None

    def __init__(self, magnitude, units=None):
        if isinstance(magnitude, Quantity):
            magnitude, units = magnitude.magnitude, (units or magnitude.units)
        if isinstance(magnitude, (list, tuple)):
            magnitude = np.asarray(magnitude, dtype=np.float64)
        self.magnitude = magnitude
        self.units = units if isinstance(units, Unit) else Unit(units)

    m = property(lambda self: self.magnitude)
    dimensionality = property(lambda self: self.units.dimensionality)

    def to(self, unit):
        tgt = Unit(unit)
        d0, f0 = _UNITS[self.units.name]
        d1, f1 = _UNITS[tgt.name]
        if d0 != d1:
            raise ValueError("cannot convert %s to %s" % (self.units, tgt))
        if f0 == f1:
            return Quantity(self.magnitude, tgt)
        return Quantity(self.magnitude * (f0 / f1), tgt)

    def m_as(self, unit):
        return self.to(unit).magnitude

    def __mul__(self, other):
        if isinstance(other, Unit):
            if self.units.name != "dimensionless":
                raise ValueError("compound units are not supported")
            return Quantity(self.magnitude, other)
        if isinstance(other, Quantity):
            if other.units.name == "dimensionless":
                return Quantity(self.magnitude * other.magnitude, self.units)
            if self.units.name == "dimensionless":
                return Quantity(self.magnitude * other.magnitude, other.units)
            raise ValueError("compound units are not supported")
        return Quantity(self.magnitude * other, self.units)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Quantity):
            if other.units.name == self.units.name:
                return Quantity(self.magnitude / other.magnitude, "dimensionless")
            if other.units.name != "dimensionless":
                raise ValueError("compound units are not supported")
            other = other.magnitude
        return Quantity(self.magnitude / other, self.units)

    def _binary(self, other, op):
        if isinstance(other, Quantity):
            other = other.to(self.units).magnitude
        elif self.units.name != "dimensionless":
            raise ValueError("cannot combine %s with a bare number" % self.units)
        return Quantity(op(self.magnitude, other), self.units)

    def __add__(self, other):
        return self._binary(other, lambda a, b: a + b)

    __radd__ = __add__

    def __sub__(self, other):
        return self._binary(other, lambda a, b: a - b)

    def __neg__(self):
        return Quantity(-self.magnitude, self.units)

    def __getitem__(self, i):
        return Quantity(self.magnitude[i], self.units)

    def __len__(self):
        return len(self.magnitude)

    def __float__(self):
        return float(self.m_as("dimensionless") if self.dimensionality == "" else self.magnitude)

    def __eq__(self, other):
        if isinstance(other, Quantity):
            try:
                return bool(np.all(self.magnitude == other.to(self.units).magnitude))
            except ValueError:
                return False
        return self.units.name == "dimensionless" and bool(np.all(self.magnitude == other))

    def __hash__(self):
        m = self.magnitude
        return hash((tuple(np.ravel(m).tolist()) if isinstance(m, np.ndarray) else m, self.units.name))

    def __repr__(self):
        return "%r %s" % (self.magnitude, self.units)


class _Registry:
    Quantity = Quantity
    Unit = Unit

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return Unit(name)

    def __call__(self, spec):
        """ureg('deg') -> Quantity(1, deg); ureg(None) -> dimensionless 1 (like pint)."""
        return Quantity(1.0, Unit(spec))


ureg = _Registry()

_PM_RE = re.compile(r"^(?P<n>[^+]+?)\+/-(?P<s>.+)$")


def parse_quantity(string):
    """'1.2 +/- 0.7 * units.meter' -> (Quantity nominal, std_dev or nan).

    Same grammar as the reference's parse_quantity (config_parser.py:303-354): spaces and the '*'
    are optional, the unit follows 'units.', the uncertainty follows '+/-'.  Raises ValueError when
    the value is not a number (the caller then treats it as a string literal)."""
    value = string.replace(" ", "")
    unit = None
    if "units." in value:
        value, unit = value.split("units.")
    value = value.rstrip("*")
    m = _PM_RE.match(value)
    if m:
        nominal, std = float(m.group("n")), float(m.group("s"))
    else:
        nominal, std = float(value), float("nan")
    return Quantity(nominal, Unit(unit)), std
