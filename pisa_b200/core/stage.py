"""``Stage``: base class of every service, same contract as pisa/core/stage.py (reference :26-586).

A service subclasses ``Stage`` and overrides ``setup_function`` / ``compute_function`` /
``apply_function`` (no arguments, no return; they read ``self.params`` and read/write ``self.data``).
The framework calls ``setup()`` once (representation = ``calc_mode``), then per template ``run()`` =
``compute()`` -- skipped when ``params.values_hash`` is unchanged, the only cache (:538-542) -- and
``apply()`` (representation = ``apply_mode``).  ``params`` must contain exactly ``expected_params``
(:270-298).  Stage / service names come from the module path ``...<stage>.<service>`` (:103-109).
"""
import time
from collections.abc import Mapping, Sequence

from pisa_b200.core.binning import MultiDimBinning
from pisa_b200.core.container import Container, ContainerSet
from pisa_b200.core.param import ParamSelector

__all__ = ["Stage"]

_MODES = ("calc_mode", "apply_mode")
_SELECTOR_KWARGS = frozenset(("regular_params", "selector_param_sets", "selections"))
_HOOKS = ("setup_function", "compute_function", "apply_function")


def _names(value, what):
    """None, one name or a sequence of names -> list of names (or None)."""
    if value is None:
        return None
    if isinstance(value, str):
        return [value]
    if isinstance(value, Sequence) and all(isinstance(v, str) for v in value):
        return list(value)
    raise TypeError("`%s` must be a string or a sequence of strings" % what)


def _selector_from(params):
    """The three ways a service may be handed its parameters (:111-124)."""
    if isinstance(params, ParamSelector):
        return params
    if isinstance(params, Mapping) and _SELECTOR_KWARGS == set(params):
        return ParamSelector(**params)
    return ParamSelector(regular_params=params)


class Stage:
    def __init__(self, data=None, params=None, expected_params=None, expected_container_keys=None,
                 debug_mode=None, error_method=None, supported_reps=None, calc_mode=None, apply_mode=None,
                 profile=False, in_standalone_mode=False):
        # `....<stage>.<service>` -> names
        *parents, service = self.__module__.split(".")
        self.stage_name, self.service_name = (parents[-1] if parents else service), service

        self.expected_params = _names(expected_params, "expected_params")
        self.expected_container_keys = _names(expected_container_keys, "expected_container_keys")
        self._param_selector = _selector_from(params)
        current = self._param_selector.params
        self._check_params(current, ignore_excess=getattr(current, "has_derived", False))
        self.validate_params(current)
        self._params = current

        self._debug_mode = debug_mode or None
        self._error_method = error_method
        # which hooks the service overrides decides which modes it can be given at all
        self.has_setup, self.has_compute, self.has_apply = (
            getattr(type(self), hook) is not getattr(Stage, hook) for hook in _HOOKS)
        self.supported_reps = self._normalised_reps(supported_reps)
        for mode, value in zip(_MODES, (calc_mode, apply_mode)):
            self._check_representation(value, mode, always_allow_none=True)
        self._calc_mode, self._apply_mode = calc_mode, apply_mode

        self.param_hash = None                       # None: never set up; -1: set up, nothing computed yet
        self.profile = profile
        self.setup_times, self.calc_times, self.apply_times = [], [], []
        self.in_standalone_mode = in_standalone_mode
        self._data = None
        self.data = data

    def __repr__(self):
        return 'Stage "%s"' % type(self).__name__

    def _normalised_reps(self, supported_reps):
        """{"calc_mode": [...], "apply_mode": [...]}: defaults are every array representation plus binnings for a
        mode whose hooks the service implements, and only ``None`` otherwise (:150-171)."""
        reps = dict(supported_reps or {})
        unknown = set(reps) - set(_MODES)
        assert not unknown, "unknown keys in supported_reps: %s" % sorted(unknown)
        used = {"calc_mode": self.has_setup or self.has_compute, "apply_mode": self.has_apply}
        for mode in _MODES:
            if mode not in reps:
                reps[mode] = [*Container.array_representations, MultiDimBinning] if used[mode] else [None]
            elif isinstance(reps[mode], str) or not isinstance(reps[mode], Sequence):
                reps[mode] = [reps[mode]]
        return reps

    # ------------------------------------------------------------------------ params -----
    def _check_params(self, params, ignore_excess=False):
        """``params`` must hold exactly ``expected_params`` (:270-298)."""
        assert self.expected_params is not None
        wanted, given = set(self.expected_params), set(params.names)
        missing, excess = sorted(wanted - given), sorted(given - wanted)
        problems = []
        if missing:
            problems.append("Missing params: %s" % ", ".join(missing))
        if excess and not ignore_excess:
            problems.append("Excess params provided: %s" % ", ".join(excess))
        if problems:
            raise ValueError("Expected parameters: %s;\n%s" % (", ".join(sorted(wanted)), ";\n".join(problems)))

    def validate_params(self, params):
        """Hook for services with constraints between parameters."""

    params = property(lambda self: self._params)
    param_selections = property(lambda self: sorted(self._param_selector.param_selections))
    debug_mode = property(lambda self: self._debug_mode)
    error_method = property(lambda self: self._error_method)

    def select_params(self, selections, error_on_missing=False):
        # selections this stage has are applied in order; a missing one raises only when asked to (stage.py:300-306
        # and param.py:1649-1684 of the reference)
        self._param_selector.select_params(selections, error_on_missing=error_on_missing)

    # ------------------------------------------------------------------------- modes -----
    def _set_mode(self, mode, value):
        if value == getattr(self, "_" + mode):
            return False
        self._check_representation(value, mode)
        setattr(self, "_" + mode, value)
        return True

    def _set_calc_mode(self, value):
        # a stand-alone stage that was already set up re-runs its setup in the new representation (:334-339)
        if self._set_mode("calc_mode", value) and self.in_standalone_mode and self.param_hash is not None:
            self.setup()

    calc_mode = property(lambda self: self._calc_mode, _set_calc_mode)
    apply_mode = property(lambda self: self._apply_mode, lambda self, value: self._set_mode("apply_mode", value))

    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, container_set):
        self._data = container_set

    def _check_representation(self, rep, mode, always_allow_none=False):
        allowed = self.supported_reps[mode]
        if rep is None:
            ok, shown = always_allow_none or None in allowed, "%s='%s'" % (mode, rep)
        elif isinstance(rep, str):
            ok, shown = rep in allowed, "%s='%s'" % (mode, rep)
        else:
            ok, shown = type(rep) in allowed, "%s of type %s" % (mode, type(rep))
        if not ok:
            raise ValueError("%s is not supported by %s.%s" % (shown, self.stage_name, self.service_name))

    is_map = property(lambda self: self.data.is_map)

    # --------------------------------------------------------------------- execution -----
    def _call(self, hook, mode, log):
        """Run one hook with the containers in the representation of `mode`, timed when profiling."""
        representation = getattr(self, mode)
        self._check_representation(representation, mode)
        if representation is not None:
            self.data.representation = representation
        if not self.profile:
            hook()
            return
        start = time.time()
        hook()
        log.append(time.time() - start)

    def setup(self):
        if self.data is not None and not isinstance(self.data, ContainerSet):
            raise TypeError("`data` must be a `pisa_b200.core.container.ContainerSet`")
        self._call(self.setup_function, "calc_mode", self.setup_times)
        self.param_hash = -1

    def compute(self):
        values_hash = self.params.values_hash
        if values_hash == self.param_hash:       # the one cache of the framework (:538-542)
            return
        self._call(self.compute_function, "calc_mode", self.calc_times)
        self.param_hash = values_hash

    def apply(self):
        self._call(self.apply_function, "apply_mode", self.apply_times)

    def run(self):
        self.compute()
        self.apply()

    def setup_function(self):
        """Overridden by services: one-time work (representation = calc_mode)."""

    def compute_function(self):
        """Overridden by services: work that depends on the parameters only (representation = calc_mode)."""

    def apply_function(self):
        """Overridden by services: per-template work (representation = apply_mode)."""

    def report_profile(self, detailed=False):
        print(self.stage_name, self.service_name)
        for label, log in (("- setup:   ", self.setup_times), ("- compute: ", self.calc_times),
                           ("- apply:   ", self.apply_times)):
            if log:
                print(label, "total %.5f s, n calls: %d, mean %.5f s" % (sum(log), len(log), sum(log) / len(log)))
            else:
                print(label, "0 calls")
