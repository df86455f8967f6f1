"""``OneDimBinning`` / ``MultiDimBinning``: the subset of pisa/core/binning.py the hot path uses.

Kept: edge construction (``np.logspace`` / ``np.linspace`` in FTYPE, binning.py:410-428), ``is_log`` /
``is_lin`` and the FTYPE-dependent ``is_irregular`` classification
(``np.allclose(ratios|diffs, first, **ALLCLOSE_KW)``, binning.py:879-890,1047-1130), ``weighted_centers``
(:900-911), ``domain``, row-major ``meshgrid`` (:2669-2712), ``+`` of two binnings, hashing by content so
a binning can key a container representation (container.py:540-553).
Dropped: slicing, rebinning, tex, serialisation, masks, VarBinning.
"""
from collections.abc import Sequence

from collections import OrderedDict

import numpy as np

from pisa_b200 import FTYPE, HASH_SIGFIGS
from pisa_b200.utils.units import Quantity, Unit, ureg

__all__ = ["OneDimBinning", "MultiDimBinning", "ALLCLOSE_KW"]

# pisa/utils/comparisons.py:81-93
FTYPE_PREC = np.finfo(FTYPE).eps
FTYPE_SIGFIGS = int(np.abs(np.ceil(np.log10(FTYPE_PREC))))
EQUALITY_PREC = 10 ** -min(HASH_SIGFIGS, FTYPE_SIGFIGS)
ALLCLOSE_KW = dict(rtol=EQUALITY_PREC, atol=FTYPE_PREC, equal_nan=True)


class OneDimBinning:
    def __init__(self, name, tex=None, bin_edges=None, units=None, domain=None, num_bins=None, is_lin=None,
                 is_log=None, bin_names=None):
        if not isinstance(name, str):
            raise TypeError("`name` must be a string")
        if is_lin and is_log:
            raise ValueError("`is_lin` and `is_log` are mutually exclusive")
        self.name, self.tex, self.bin_names = name, tex, bin_names
        if units is not None and not isinstance(units, Unit):
            units = units.units if isinstance(units, Quantity) else Unit(units)
        edges = None
        if bin_edges is not None:
            if isinstance(bin_edges, Quantity):
                if units is None:
                    units = bin_edges.units
                bin_edges = bin_edges.to(units).magnitude
            edges = np.asarray(bin_edges, dtype=FTYPE)
        dom = None
        if domain is not None:
            if isinstance(domain, Quantity):
                if units is None:
                    units = domain.units
                domain = domain.to(units).magnitude
            dom = (float(domain[0]), float(domain[1]))
        if units is None:
            units = ureg.dimensionless
        if edges is None:
            if num_bins is None or dom is None:
                raise ValueError("If not specifying bin edges explicitly, `domain` and `num_bins` must be "
                                 "specified (and optionally set `is_log=True`).")
            if is_log:
                edges = np.logspace(np.log10(dom[0]), np.log10(dom[1]), num_bins + 1, dtype=FTYPE)
            else:
                edges = np.linspace(dom[0], dom[1], num_bins + 1, dtype=FTYPE)
        elif dom is not None:
            assert dom[0] == edges[0] and dom[1] == edges[-1]
        if num_bins is not None and num_bins != len(edges) - 1:
            raise AssertionError("%s, %s" % (num_bins, edges))
        if len(edges) < 2 or np.any(np.diff(edges) <= 0):
            raise ValueError("bin edges must be monotonically increasing and define at least one bin")
        self._edges = edges
        self._units = units
        self._is_log = bool(is_log)
        self._is_irregular = None

    # --- basic properties -------------------------------------------------------------------
    units = property(lambda self: self._units)
    num_bins = property(lambda self: len(self._edges) - 1)
    edge_magnitudes = property(lambda self: self._edges)
    is_log = property(lambda self: self._is_log)
    is_lin = property(lambda self: not self._is_log)
    size = property(lambda self: len(self._edges) - 1)
    shape = property(lambda self: (len(self._edges) - 1,))

    @property
    def bin_edges(self):
        return Quantity(self._edges, self._units)

    @property
    def domain(self):
        return Quantity(np.array([np.min(self._edges), np.max(self._edges)]), self._units)

    @property
    def midpoints(self):
        return Quantity((self._edges[:-1] + self._edges[1:]) / 2.0, self._units)

    @property
    def weighted_centers(self):
        if self.is_log:
            return Quantity(np.sqrt(self._edges[:-1] * self._edges[1:]), self._units)
        return self.midpoints

    @staticmethod
    def is_bin_spacing_log_uniform(bin_edges):
        e = np.asarray(getattr(bin_edges, "magnitude", bin_edges))
        if len(e) < 3:
            raise ValueError("%d bin edge(s) passed; require at least 3 to determine nature of bin spacing." % len(e))
        with np.errstate(divide="raise", over="raise", under="raise", invalid="raise"):
            try:
                ratio = e[1:] / e[:-1]
            except (AssertionError, FloatingPointError, ZeroDivisionError):
                return False
        return bool(np.allclose(ratio, ratio[0], **ALLCLOSE_KW))

    @staticmethod
    def is_bin_spacing_lin_uniform(bin_edges):
        e = np.array(getattr(bin_edges, "magnitude", bin_edges))
        if len(e) == 1:
            raise ValueError("Single bin edge passed; require at least 2 to determine nature of bin spacing.")
        if not np.all(np.isfinite(e)):
            return False
        if len(e) == 2:
            return True
        d = np.diff(e)
        return bool(np.allclose(d, d[0], **ALLCLOSE_KW))

    @property
    def is_irregular(self):
        """binning.py:879-890: NOT uniform in the space (lin or log) the dimension is declared in."""
        if self._is_irregular is None:
            if self.num_bins == 1:
                self._is_irregular = False
            elif self.is_log:
                self._is_irregular = not self.is_bin_spacing_log_uniform(self._edges)
            else:
                self._is_irregular = not self.is_bin_spacing_lin_uniform(self._edges)
        return self._is_irregular

    # --- serialisation (binning.py:551-600,676-694) -----------------------------------------
    @property
    def serializable_state(self):
        return OrderedDict([("name", self.name), ("bin_edges", self._edges), ("units", str(self._units)),
                            ("is_log", self.is_log), ("is_lin", self.is_lin), ("bin_names", self.bin_names),
                            ("tex", self.tex)])

    def to_json(self, filename, **kwargs):
        from pisa_b200.utils import jsons
        jsons.to_json(self.serializable_state, filename=filename, **kwargs)

    @classmethod
    def from_json(cls, resource):
        from pisa_b200.utils import jsons
        return cls(**jsons.from_json(resource))

    # --- identity -----------------------------------------------------------------------------
    def _state(self):
        # binnings are immutable: the rounded-edge state (HASH_SIGFIGS significant figures, like the
        # reference's normQuant-based hashes) is computed once -- representations are hashed on every
        # container access, and re-formatting 200 edges each time cost 200 ms per template
        st = getattr(self, "_state_cache", None)
        if st is None:
            with np.errstate(invalid="ignore", over="ignore"):
                rounded = tuple(float(np.format_float_scientific(x, precision=HASH_SIGFIGS)) if np.isfinite(x) else x
                                for x in self._edges)
            st = (self.name, self._units.name, self._is_log, rounded)
            self._state_cache = st
            self._hash_cache = hash(st)
        return st

    def __hash__(self):
        self._state()
        return self._hash_cache

    def __eq__(self, other):
        return isinstance(other, OneDimBinning) and self._state() == other._state()

    def __len__(self):
        return self.num_bins

    def __repr__(self):
        return "OneDimBinning(%r, %d %s bins, [%g, %g] %s)" % (
            self.name, self.num_bins, "log" if self.is_log else "lin", self._edges[0], self._edges[-1], self._units)


class MultiDimBinning:
    def __init__(self, dimensions, name=None, mask=None):
        if isinstance(dimensions, OneDimBinning):
            dimensions = [dimensions]
        if isinstance(dimensions, MultiDimBinning):
            dimensions = list(dimensions)
        if mask is not None:
            raise NotImplementedError("bin masks are out of scope of pisa_b200")
        dims = []
        for d in dimensions:
            if isinstance(d, OneDimBinning):
                dims.append(d)
            elif isinstance(d, dict):
                dims.append(OneDimBinning(**d))
            else:
                raise TypeError("dimensions must be OneDimBinning objects or kwargs dicts")
        names = [d.name for d in dims]
        if len(set(names)) != len(names):
            raise ValueError("dimension names must be unique")
        self._dims = tuple(dims)
        self.name = name
        # binnings are immutable and these are read on every container access
        self._names = [d.name for d in dims]
        self._shape = tuple(d.num_bins for d in dims)
        self._size = int(np.prod(self._shape))

    dimensions = property(lambda self: self._dims)
    names = property(lambda self: list(self._names))
    num_dims = property(lambda self: len(self._dims))
    shape = property(lambda self: self._shape)
    num_bins = property(lambda self: [d.num_bins for d in self._dims])
    size = property(lambda self: self._size)
    tot_num_bins = size
    bin_edges = property(lambda self: [d.bin_edges for d in self._dims])
    domains = property(lambda self: [d.domain for d in self._dims])
    weighted_centers = property(lambda self: [d.weighted_centers for d in self._dims])
    midpoints = property(lambda self: [d.midpoints for d in self._dims])
    is_irregular = property(lambda self: any(d.is_irregular for d in self._dims))
    is_lin = property(lambda self: all(d.is_lin for d in self._dims))
    is_log = property(lambda self: all(d.is_log for d in self._dims))

    def iterdims(self):
        return iter(self._dims)

    def __iter__(self):
        return iter(self._dims)

    def __len__(self):
        return len(self._dims)

    def index(self, dim):
        if isinstance(dim, OneDimBinning):
            dim = dim.name
        if isinstance(dim, str):
            if dim not in self.names:
                raise ValueError("Dimension %r not present; have %s" % (dim, self.names))
            return self.names.index(dim)
        if isinstance(dim, int) and 0 <= dim < len(self):
            return dim
        raise ValueError("cannot locate dimension %r" % (dim,))

    def __getitem__(self, key):
        if isinstance(key, (str, int)):
            return self._dims[self.index(key)]
        raise TypeError("only dimension lookup by name / index is supported")

    def __getattr__(self, attr):
        if attr.startswith("_"):
            raise AttributeError(attr)
        for d in self._dims:
            if d.name == attr:
                return d
        raise AttributeError(attr)

    def __add__(self, other):
        other = MultiDimBinning(other) if not isinstance(other, MultiDimBinning) else other
        return MultiDimBinning(list(self._dims) + list(other._dims))

    def meshgrid(self, entity, attach_units=True):
        """numpy.meshgrid(indexing='ij') over an entity of every dimension (binning.py:2669-2712)."""
        arrays = [np.asarray(getattr(d, entity.lower().strip()).magnitude) for d in self._dims]
        grids = [g.astype(FTYPE) for g in np.meshgrid(*arrays, indexing="ij", copy=False)]
        if attach_units:
            return [Quantity(g, d.units) for g, d in zip(grids, self._dims)]
        return grids

    # --- serialisation (binning.py:1680-1740,1842-1859) ---------------------------------------
    @property
    def serializable_state(self):
        return OrderedDict([("dimensions", [d.serializable_state for d in self._dims]), ("name", self.name),
                            ("mask", None)])

    def to_json(self, filename, **kwargs):
        from pisa_b200.utils import jsons
        jsons.to_json(self.serializable_state, filename=filename, **kwargs)

    @classmethod
    def from_json(cls, resource):
        from pisa_b200.utils import jsons
        return cls(**jsons.from_json(resource))

    def __hash__(self):
        h = getattr(self, "_hash_cache", None)
        if h is None:
            h = self._hash_cache = hash(tuple(hash(d) for d in self._dims))
        return h

    def __eq__(self, other):
        if self is other:
            return True
        return isinstance(other, MultiDimBinning) and hash(self) == hash(other) and self._dims == other._dims

    def __repr__(self):
        return "MultiDimBinning(%s)" % ", ".join(repr(d) for d in self._dims)
