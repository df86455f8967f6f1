"""``Stage``: base class of every service, same contract as pisa/core/stage.py (reference :26-586).

A service subclasses ``Stage`` and overrides ``setup_function`` / ``compute_function`` /
``apply_function`` (no arguments, no return; they read ``self.params`` and read/write ``self.data``).
The framework calls ``setup()`` once (representation = ``calc_mode``), then per template ``run()`` =
``compute()`` -- skipped when ``params.values_hash`` is unchanged, the only cache (:538-542) -- and
``apply()`` (representation = ``apply_mode``).  ``params`` must contain exactly ``expected_params``
(:270-298).  Stage / service names come from the module path ``...<stage>.<service>`` (:103-109).
"""
from collections.abc import Mapping, Sequence
from time import time

from pisa_b200.core.binning import MultiDimBinning
from pisa_b200.core.container import Container, ContainerSet
from pisa_b200.core.param import ParamSelector, ParamSet

__all__ = ["Stage"]


def _str_seq(inputs, name):
    if inputs is None:
        return None
    if isinstance(inputs, str):
        return [inputs]
    if not isinstance(inputs, Sequence) or not all(isinstance(i, str) for i in inputs):
        raise TypeError("`%s` must be a string or a sequence of strings" % name)
    return list(inputs)


class Stage:
    def __init__(self, data=None, params=None, expected_params=None, expected_container_keys=None,
                 debug_mode=None, error_method=None, supported_reps=None, calc_mode=None, apply_mode=None,
                 profile=False, in_standalone_mode=False):
        expected_params = _str_seq(expected_params, "expected_params")
        expected_container_keys = _str_seq(expected_container_keys, "expected_container_keys")
        module_path = self.__module__.split(".")
        self.stage_name = module_path[-2] if len(module_path) > 1 else module_path[-1]
        self.service_name = module_path[-1]
        self.expected_params = expected_params
        self.expected_container_keys = expected_container_keys

        selector_keys = {"regular_params", "selector_param_sets", "selections"}
        if isinstance(params, Mapping) and set(params.keys()) == selector_keys:
            self._param_selector = ParamSelector(**params)
        elif isinstance(params, ParamSelector):
            self._param_selector = params
        else:
            self._param_selector = ParamSelector(regular_params=params)
        p = self._param_selector.params
        self._check_params(p, getattr(p, "has_derived", False))
        self.validate_params(p)
        self._params = p

        self._debug_mode = debug_mode if bool(debug_mode) else None
        self.has_setup = type(self).setup_function is not Stage.setup_function
        self.has_compute = type(self).compute_function is not Stage.compute_function
        self.has_apply = type(self).apply_function is not Stage.apply_function

        supported_reps = dict(supported_reps or {})
        assert set(supported_reps.keys()).issubset(("calc_mode", "apply_mode"))
        for mode_str in ("calc_mode", "apply_mode"):
            allowed = (self.has_setup or self.has_compute) if mode_str == "calc_mode" else self.has_apply
            if mode_str not in supported_reps:
                supported_reps[mode_str] = (list(Container.array_representations) + [MultiDimBinning]
                                            if allowed else [None])
            elif isinstance(supported_reps[mode_str], str) or not isinstance(supported_reps[mode_str], Sequence):
                supported_reps[mode_str] = [supported_reps[mode_str]]
        self.supported_reps = supported_reps

        self._check_representation(calc_mode, "calc_mode", always_allow_none=True)
        self._calc_mode = calc_mode
        self._check_representation(apply_mode, "apply_mode", always_allow_none=True)
        self._apply_mode = apply_mode
        self._error_method = error_method
        self.param_hash = None
        self.profile = profile
        self.setup_times, self.calc_times, self.apply_times = [], [], []
        self.in_standalone_mode = in_standalone_mode
        self._data = None
        self.data = data

    def __repr__(self):
        return 'Stage "%s"' % self.__class__.__name__

    # ------------------------------------------------------------------------ params -----
    def _check_params(self, params, ignore_excess=False):
        assert self.expected_params is not None
        exp_p, got_p = set(self.expected_params), set(params.names)
        if exp_p == got_p:
            return
        excess, missing = got_p - exp_p, exp_p - got_p
        errs = []
        if missing:
            errs.append("Missing params: %s" % ", ".join(sorted(missing)))
        if excess:
            if ignore_excess:
                if not errs:
                    return
            else:
                errs.append("Excess params provided: %s" % ", ".join(sorted(excess)))
        raise ValueError("Expected parameters: %s;\n" % ", ".join(sorted(exp_p)) + ";\n".join(errs))

    def validate_params(self, params):
        return

    params = property(lambda self: self._params)
    param_selections = property(lambda self: sorted(self._param_selector.param_selections))
    debug_mode = property(lambda self: self._debug_mode)
    error_method = property(lambda self: self._error_method)

    def select_params(self, selections, error_on_missing=False):
        try:
            self._param_selector.select_params(selections, error_on_missing=True)
        except KeyError:
            if error_on_missing:
                raise

    # ------------------------------------------------------------------------- modes -----
    @property
    def calc_mode(self):
        return self._calc_mode

    @calc_mode.setter
    def calc_mode(self, value):
        if value != self._calc_mode:
            self._check_representation(value, "calc_mode")
            self._calc_mode = value
            if self.in_standalone_mode and self.param_hash is not None:
                self.setup()

    @property
    def apply_mode(self):
        return self._apply_mode

    @apply_mode.setter
    def apply_mode(self, value):
        if value != self._apply_mode:
            self._check_representation(value, "apply_mode")
            self._apply_mode = value

    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, value):
        self._data = value

    def _check_representation(self, rep, mode, always_allow_none=False):
        where = "%s.%s" % (self.stage_name, self.service_name)
        if rep is None:
            if None not in self.supported_reps[mode] and not always_allow_none:
                raise ValueError("%s='%s' is not supported by %s" % (mode, rep, where))
        elif isinstance(rep, str):
            if rep not in self.supported_reps[mode]:
                raise ValueError("%s='%s' is not supported by %s" % (mode, rep, where))
        elif type(rep) not in self.supported_reps[mode]:
            raise ValueError("%s of type %s is not supported by %s" % (mode, type(rep), where))

    @property
    def is_map(self):
        return self.data.is_map

    # --------------------------------------------------------------------- execution -----
    def _timed(self, fn, times):
        if self.profile:
            t0 = time()
            fn()
            times.append(time() - t0)
        else:
            fn()

    def setup(self):
        if self.data is not None and not isinstance(self.data, ContainerSet):
            raise TypeError("`data` must be a `pisa_b200.core.container.ContainerSet`")
        self._check_representation(self.calc_mode, "calc_mode")
        if self.calc_mode is not None:
            self.data.representation = self.calc_mode
        self._timed(self.setup_function, self.setup_times)
        self.param_hash = -1

    def compute(self):
        new_hash = self.params.values_hash
        if new_hash == self.param_hash:
            return
        self._check_representation(self.calc_mode, "calc_mode")
        if self.calc_mode is not None:
            self.data.representation = self.calc_mode
        self._timed(self.compute_function, self.calc_times)
        self.param_hash = new_hash

    def apply(self):
        self._check_representation(self.apply_mode, "apply_mode")
        if self.apply_mode is not None:
            self.data.representation = self.apply_mode
        self._timed(self.apply_function, self.apply_times)

    def run(self):
        self.compute()
        self.apply()

    def setup_function(self):
        """Implement in services (subclasses of Stage)"""

    def compute_function(self):
        """Implement in services (subclasses of Stage)"""

    def apply_function(self):
        """Implement in services (subclasses of Stage)"""

    def report_profile(self, detailed=False):
        print(self.stage_name, self.service_name)
        for label, times in (("- setup:   ", self.setup_times), ("- compute: ", self.calc_times),
                             ("- apply:   ", self.apply_times)):
            if times:
                print(label, "total %.5f s, n calls: %d, mean %.5f s" % (sum(times), len(times), sum(times) / len(times)))
            else:
                print(label, "0 calls")
