"""Parameters of a stage: ``Param``, ``ParamSet`` and ``ParamSelector``.

Thin restatement of the part of pisa/core/param.py (reference :77,769,1604) that stages and the
pipeline touch: ``params.<name>.value.m_as(unit)`` (prob3.py:485), name checking against
``expected_params`` (stage.py:270-298), the ``values_hash`` that drives the compute cache
(stage.py:538-542), fixed/free bookkeeping and ``nh``/``ih``-style selections
(config_parser.py:845-905).  Priors, ranges as rescaled values, serialisation are out of scope.
"""
from collections import OrderedDict

import numpy as np

from pisa_b200.utils.units import Quantity

__all__ = ["Param", "ParamSet", "ParamSelector"]


class Param:
    def __init__(self, name, value, prior=None, range=None, is_fixed=True, unique_id=None, tex=None,
                 scales_as_log=False, help=None, nominal_value=None):
        self.name = name
        self.unique_id = unique_id if unique_id is not None else name
        self.tex = tex
        self.help = help
        self.prior = prior
        self.range = range
        self.is_fixed = bool(is_fixed)
        self.scales_as_log = scales_as_log
        self._value = None
        self.value = value
        self.nominal_value = self._value if nominal_value is None else nominal_value

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, val):
        if isinstance(val, (int, float, np.integer, np.floating)) and not isinstance(val, bool):
            val = Quantity(float(val), "dimensionless")
        if self._value is not None and isinstance(self._value, Quantity) and isinstance(val, Quantity):
            # keep the declared units, like pint-backed Param does (param.py value setter)
            if val.dimensionality != self._value.dimensionality:
                raise ValueError("Param %s: value %r has wrong dimensionality" % (self.name, val))
        self._value = val

    def m_as(self, unit):
        """Shortcut used by some services (aeff.py:68-72): ``param.m_as('sec')``."""
        return self._value.m_as(unit)

    @property
    def m(self):
        return self._value.magnitude if isinstance(self._value, Quantity) else self._value

    @property
    def units(self):
        return self._value.units if isinstance(self._value, Quantity) else None

    @property
    def state(self):
        v = self._value
        if isinstance(v, Quantity):
            m = v.magnitude
            v = (tuple(np.ravel(m).tolist()) if isinstance(m, np.ndarray) else float(m), v.units.name)
        return (self.name, v, self.is_fixed)

    def reset(self):
        self.value = self.nominal_value

    def __repr__(self):
        return "Param(%s=%r, fixed=%s)" % (self.name, self._value, self.is_fixed)


class ParamSet:
    """Ordered, name-addressable collection of Params."""

    def __init__(self, *args):
        params = []
        for a in args:
            if a is None:
                continue
            if isinstance(a, Param):
                params.append(a)
            elif isinstance(a, ParamSet):
                params.extend(a._params)
            else:
                params.extend(list(a))
        object.__setattr__(self, "_params", [])
        for p in params:
            self.update(p)

    has_derived = False

    @property
    def names(self):
        return tuple(p.name for p in self._params)

    def index(self, name):
        for i, p in enumerate(self._params):
            if p.name == name:
                return i
        raise ValueError("No parameter named %r" % name)

    def update(self, obj, existing_must_match=False, extend=True):
        objs = [obj] if isinstance(obj, Param) else list(obj)
        for p in objs:
            if p.name in self.names:
                self._params[self.index(p.name)] = p
            elif extend:
                self._params.append(p)
            # (extend=False: params this set does not hold are ignored -- Pipeline.update_params passes every
            # param to every stage, pipeline.py:579-596 of the reference)

    def extend(self, obj):
        self.update(obj, extend=True)

    def __getattr__(self, name):
        try:
            return self._params[self.index(name)]
        except ValueError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        # params.theta23 = 45 * ureg.deg  (pipeline.params.<name> = value idiom)
        if name in self.names:
            self._params[self.index(name)].value = value
        else:
            object.__setattr__(self, name, value)

    def __getitem__(self, i):
        if isinstance(i, str):
            return self._params[self.index(i)]
        return self._params[i]

    def __iter__(self):
        return iter(self._params)

    def __len__(self):
        return len(self._params)

    def __contains__(self, name):
        return (name.name if isinstance(name, Param) else name) in self.names

    @property
    def free(self):
        return ParamSet([p for p in self._params if not p.is_fixed])

    @property
    def fixed(self):
        return ParamSet([p for p in self._params if p.is_fixed])

    def fix(self, names):
        for n in ([names] if isinstance(names, str) else names):
            self[n].is_fixed = True

    def unfix(self, names):
        for n in ([names] if isinstance(names, str) else names):
            self[n].is_fixed = False

    def reset_all(self):
        for p in self._params:
            p.reset()

    @property
    def values_hash(self):
        """Changes iff any parameter value changes: the stage compute-cache key (stage.py:538-542)."""
        return hash(tuple(p.state[:2] for p in self._params))

    @property
    def hash(self):
        return hash(tuple(p.state for p in self._params))

    def __repr__(self):
        return "ParamSet(%s)" % ", ".join(self.names)


class ParamSelector:
    """Regular params + alternative sets chosen by name (e.g. 'nh' / 'ih')."""

    def __init__(self, regular_params=None, selector_param_sets=None, selections=None):
        self._regular = ParamSet(regular_params)
        self._selector_sets = OrderedDict()
        if selector_param_sets:
            for sel, ps in selector_param_sets.items():
                self._selector_sets[sel.strip().lower()] = ParamSet(ps)
        if isinstance(selections, str):
            selections = [selections]
        self._selections = [s.strip().lower() for s in (selections or [])]
        self._current = None
        self._rebuild()

    def _rebuild(self):
        cur = ParamSet(self._regular)
        for sel in self._selections:
            if sel in self._selector_sets:
                cur.update(self._selector_sets[sel])
        self._current = cur

    @property
    def params(self):
        return self._current

    @property
    def param_selections(self):
        return list(self._selections)

    def select_params(self, selections=None, error_on_missing=False):
        """Apply the named selector sets to the current params IN ORDER (param.py:1649-1684 of the reference):
        every selection that exists is applied before a missing one raises (``error_on_missing``) or is skipped."""
        if selections is None:
            selections = list(self._selections)
        if isinstance(selections, str):
            selections = selections.split(",")
        distilled = []
        for sel in selections:
            if sel is None:
                continue
            if not isinstance(sel, str):
                raise ValueError("Selection should be a str. Got %s instead." % type(sel))
            sel = sel.strip().lower()
            if sel in self._selector_sets:
                # in place, so that stages holding a reference to `.params` see the change
                self._current.update(self._selector_sets[sel])
            elif error_on_missing:
                raise KeyError('No selection "%s" available; valid selections are %s (case-insensitive).'
                               % (sel, list(self._selector_sets)))
            distilled.append(sel)
        self._selections = sorted(distilled)
        return self._current

    def update(self, p, selector=None, existing_must_match=False, extend=True):
        """Reference semantics (param.py:1708-1730): without ``selector`` the regular set, the current set AND
        every active selector set take the new params, so that a later ``select_params`` does not revert them."""
        p = p if isinstance(p, ParamSet) else ParamSet(p)
        if selector is None:
            self._regular.update(p, existing_must_match=existing_must_match, extend=extend)
            self._current.update(p, existing_must_match=existing_must_match, extend=extend)
            for sel in self._selections:
                if sel in self._selector_sets:
                    self._selector_sets[sel].update(p, existing_must_match=existing_must_match, extend=extend)
        else:
            sel = selector.strip().lower()
            self._selector_sets.setdefault(sel, ParamSet()).update(p, existing_must_match=existing_must_match,
                                                                   extend=extend)
            self.select_params(error_on_missing=False)   # re-select in case the update touched an active set

    def get(self, name, selector=None):
        if selector is None:
            if name in self._regular:
                return self._regular[name]
            raise KeyError(name)
        sel = selector.strip().lower()
        if sel in self._selector_sets and name in self._selector_sets[sel]:
            return self._selector_sets[sel][name]
        raise KeyError((name, selector))

    def __iter__(self):
        seen = []
        for p in self._regular:
            seen.append(p)
        for ps in self._selector_sets.values():
            for p in ps:
                seen.append(p)
        return iter(seen)
