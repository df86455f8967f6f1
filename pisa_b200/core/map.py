"""``Map`` / ``MapSet``: histogram + per-bin error on a ``MultiDimBinning`` (host side).

The thin subset of pisa/core/map.py (:221,2108) that ``ContainerSet.get_mapset`` and a chi-square
scan need: ``Map(name, hist, error_hist, binning)`` (container.py:792-800), ``nominal_values`` /
``std_devs`` (the reference packs both into ``uncertainties`` arrays), sums and element-wise
arithmetic of same-binning maps, ``MapSet`` lookup by name and summation, and ``to_json`` / ``from_json``
in the reference's file layout (map.py:1272-1362,2206-2262) so that templates can be handed to real PISA
tooling.  No plotting, slicing, rebinning or fluctuation.
"""
from collections import OrderedDict

import numpy as np

from pisa_b200.core.binning import MultiDimBinning

__all__ = ["Map", "MapSet"]


class Map:
    def __init__(self, name, hist, binning, error_hist=None, hash=None, tex=None, full_comparison=False):
        if isinstance(binning, dict):
            binning = MultiDimBinning(**binning)
        elif not isinstance(binning, MultiDimBinning):
            binning = MultiDimBinning(binning)
        hist = np.asarray(hist)
        if hist.shape != binning.shape:
            raise ValueError("hist shape %s does not match binning shape %s" % (hist.shape, binning.shape))
        self.name, self.binning, self.tex = name, binning, tex
        self.hash, self.full_comparison = hash, full_comparison
        self._hist = hist
        self._err = None
        if error_hist is not None:
            error_hist = np.asarray(error_hist)
            if error_hist.shape != binning.shape:
                raise ValueError("error_hist shape mismatch")
            self._err = error_hist

    @property
    def hist(self):
        return self._hist

    nominal_values = hist

    @property
    def std_devs(self):
        return np.zeros_like(self._hist) if self._err is None else self._err

    @property
    def shape(self):
        return self.binning.shape

    def set_poisson_errors(self):
        self._err = np.sqrt(self._hist)

    def set_errors(self, error_hist):
        self._err = None if error_hist is None else np.asarray(error_hist)

    def sum(self):
        return float(self._hist.sum())

    def _combine(self, other, op, err_op):
        if isinstance(other, Map):
            if other.binning != self.binning:
                raise ValueError("maps have different binnings")
            hist = op(self._hist, other._hist)
            err = err_op(self, other)
        else:
            hist = op(self._hist, other)
            err = None if self._err is None else np.abs(op(self._err, other) if op in (np.multiply, np.divide) else self._err)
        return Map(self.name, hist, self.binning, error_hist=err)

    def __add__(self, other):
        quad = lambda a, b: None if (a._err is None and b._err is None) else np.sqrt(a.std_devs ** 2 + b.std_devs ** 2)  # noqa: E731
        return self._combine(other, np.add, quad)

    __radd__ = __add__

    def __sub__(self, other):
        quad = lambda a, b: None if (a._err is None and b._err is None) else np.sqrt(a.std_devs ** 2 + b.std_devs ** 2)  # noqa: E731
        return self._combine(other, np.subtract, quad)

    def __mul__(self, other):
        if isinstance(other, Map):
            rel = lambda a, b: None if (a._err is None and b._err is None) else np.sqrt(  # noqa: E731
                (a.std_devs * b._hist) ** 2 + (b.std_devs * a._hist) ** 2)
            return self._combine(other, np.multiply, rel)
        return self._combine(other, np.multiply, None)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Map):
            with np.errstate(divide="ignore", invalid="ignore"):
                rel = lambda a, b: None if (a._err is None and b._err is None) else np.sqrt(  # noqa: E731
                    (a.std_devs / b._hist) ** 2 + (b.std_devs * a._hist / b._hist ** 2) ** 2)
                return self._combine(other, np.divide, rel)
        return self._combine(other, np.divide, None)

    def allclose(self, other, rtol=1e-12, atol=0.0):
        return self.binning == other.binning and np.allclose(self._hist, other._hist, rtol=rtol, atol=atol)

    def mod_chi2(self, expected):
        """sum over bins of (N_obs - N_exp)^2 / (sigma_exp^2 + N_exp) (pisa/utils/stats.py:651-695)."""
        e = np.clip(expected.hist, 1e-10, np.inf)
        return float(((self._hist - e) ** 2 / (expected.std_devs ** 2 + e)).sum())

    # --- serialisation (map.py:1272-1362) ----------------------------------------------------
    @property
    def serializable_state(self):
        stddevs = self.std_devs
        return OrderedDict([("name", self.name), ("hist", self._hist),
                            ("binning", self.binning.serializable_state),
                            ("error_hist", None if np.all(stddevs == 0) else stddevs), ("hash", self.hash),
                            ("tex", self.tex), ("full_comparison", self.full_comparison)])

    def to_json(self, filename, **kwargs):
        from pisa_b200.utils import jsons
        jsons.to_json(self.serializable_state, filename=filename, **kwargs)

    @classmethod
    def from_json(cls, resource):
        from pisa_b200.utils import jsons
        return cls(**jsons.from_json(resource))

    def __repr__(self):
        return "Map(%r, sum=%g, shape=%s)" % (self.name, self.sum(), self.shape)


class MapSet:
    def __init__(self, maps, name=None, tex=None, hash=None, collate_by_name=True):
        self.maps = [m if isinstance(m, Map) else Map(**m) for m in maps]
        self.name, self.tex, self.hash, self.collate_by_name = name, tex, hash, collate_by_name

    names = property(lambda self: [m.name for m in self.maps])

    def __iter__(self):
        return iter(self.maps)

    def __len__(self):
        return len(self.maps)

    def __getitem__(self, key):
        if isinstance(key, int):
            return self.maps[key]
        for m in self.maps:
            if m.name == key:
                return m
        raise KeyError(key)

    def __contains__(self, name):
        return name in self.names

    def sum(self):
        """Sum of all maps as one Map (``sum(mapset)`` / DistributionMaker return_sum)."""
        total = self.maps[0]
        for m in self.maps[1:]:
            total = total + m
        total.name = self.name or "total"
        return total

    def __add__(self, other):
        if isinstance(other, MapSet):
            return MapSet([a + other[a.name] for a in self.maps], name=self.name)
        return MapSet([a + other for a in self.maps], name=self.name)

    # --- serialisation (map.py:2206-2262) ----------------------------------------------------
    @property
    def serializable_state(self):
        return OrderedDict([("maps", [m.serializable_state for m in self.maps]), ("name", self.name),
                            ("tex", self.tex), ("collate_by_name", self.collate_by_name)])

    def to_json(self, filename, **kwargs):
        from pisa_b200.utils import jsons
        jsons.to_json(self.serializable_state, filename=filename, **kwargs)

    @classmethod
    def from_json(cls, resource):
        from pisa_b200.utils import jsons
        return cls(**jsons.from_json(resource))

    def __repr__(self):
        return "MapSet(%r: %s)" % (self.name, self.names)
