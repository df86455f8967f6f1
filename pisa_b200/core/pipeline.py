"""``Pipeline``: ordered stages sharing one ``ContainerSet``; ``run()`` / ``get_outputs()`` -> ``MapSet``.

Same behaviour as pisa/core/pipeline.py (reference :41-1225) for the calls the hot path needs:
the stage factory imports ``<package>.stages.<stage>.<service>`` and, failing that, the external
module ``<stage>.<service>`` (:249-358; here the package is ``pisa_b200``, so the reference's cfg
files select the B200 services), instantiates ``service_cls(**settings, profile=...)`` and insists on
a ``Stage``; ``setup()`` builds a fresh ``ContainerSet`` and runs every stage's ``setup`` (:570-577);
``run()`` = ``stage.run()`` in order (:554-558); ``get_outputs()`` switches the data to the output
binning and returns ``data.get_mapset(output_key[, error])`` (:372-387,451-483).
"""
from collections import OrderedDict
from importlib import import_module
from time import time

from pisa_b200.core.binning import MultiDimBinning
from pisa_b200.core.container import ContainerSet
from pisa_b200.core.param import ParamSet
from pisa_b200.core.stage import Stage
from pisa_b200.utils.config_parser import PISAConfigParser, parse_pipeline_config

__all__ = ["Pipeline"]

STAGE_PACKAGES = ("pisa_b200.stages",)


class Pipeline:
    def __init__(self, config, profile=False):
        if isinstance(config, (str, PISAConfigParser)):
            config = parse_pipeline_config(config=config)
        elif not isinstance(config, OrderedDict):
            raise TypeError("`config` passed is of type %s but must be string, PISAConfigParser, or OrderedDict"
                            % type(config).__name__)
        self.name = config["pipeline"]["name"]
        self.detector_name = config["pipeline"]["detector_name"]
        self.data = ContainerSet(self.name)
        self.data["output_binning"] = config["pipeline"]["output_binning"]
        self.output_key = config["pipeline"]["output_key"]
        self._profile = profile
        self._setup_times, self._run_times, self._get_outputs_times = [], [], []
        self._stages = []
        self._config = config
        self._init_stages()

    config = property(lambda self: self._config)
    stages = property(lambda self: list(self._stages))
    stage_names = property(lambda self: [s.stage_name for s in self._stages])
    service_names = property(lambda self: [s.service_name for s in self._stages])
    profile = property(lambda self: self._profile)

    def __iter__(self):
        return iter(self._stages)

    def __getitem__(self, idx):
        if isinstance(idx, str):
            for s in self._stages:
                if idx in (s.stage_name, s.service_name, "%s.%s" % (s.stage_name, s.service_name)):
                    return s
            raise KeyError(idx)
        return self._stages[idx]

    def _init_stages(self):
        stages = []
        for name, settings in self._config.items():
            if name == "pipeline":
                continue
            stage_name, service_name = name
            service_name = service_name.replace("pi_", "")
            module = None
            for pkg in STAGE_PACKAGES:
                try:
                    module = import_module("%s.%s.%s" % (pkg, stage_name, service_name))
                    break
                except ImportError:
                    continue
            if module is None:   # external definition, like the reference (:284-293)
                module = import_module("%s.%s" % (stage_name, service_name))
            service_cls = getattr(module, service_name)
            service = service_cls(**settings, profile=self._profile)
            if not isinstance(service, Stage):
                raise TypeError('Trying to create service "%s" (%s), but object %s instantiated from class %s is not '
                                "a Stage type but instead is of type %s."
                                % (service_name, stage_name, service, service_cls, type(service)))
            stages.append(service)
        self._stages = stages
        selections = sorted({sel for s in stages for sel in s.param_selections})
        for s in stages:
            s.select_params(selections, error_on_missing=False)
        self.setup()

    # ----------------------------------------------------------------------------- params -----
    @property
    def params(self):
        ps = ParamSet()
        for s in self._stages:
            ps.extend(s.params)
        return ps

    @property
    def param_selections(self):
        return sorted({sel for s in self._stages for sel in s.param_selections})

    def select_params(self, selections, error_on_missing=False):
        """pipeline.py:598-625 of the reference: every stage applies the selections it has; KeyError only if asked
        for and NO stage has all of them."""
        successes = 0
        for s in self._stages:
            try:
                s.select_params(selections, error_on_missing=True)
            except KeyError:
                pass
            else:
                successes += 1
        if error_on_missing and successes == 0:
            raise KeyError("None of the stages in this pipeline has all of the selections %s available."
                           % (selections,))

    def update_params(self, params, existing_must_match=False, extend=False):
        """Through every stage's ParamSelector (pipeline.py:579-596 of the reference): the regular set, the current set
        and the active selector sets are updated, so a later ``select_params`` does not revert the new values."""
        plist = [params] if not hasattr(params, "__iter__") else list(params)
        for s in self._stages:
            s._param_selector.update(plist, existing_must_match=existing_must_match, extend=extend)

    # -------------------------------------------------------------------------- execution -----
    @property
    def output_binning(self):
        return self.data["output_binning"]

    @output_binning.setter
    def output_binning(self, binning):
        self.data["output_binning"] = binning
        self.setup()

    def _timed(self, fn, times):
        if self._profile:
            t0 = time()
            out = fn()
            times.append(time() - t0)
            return out
        return fn()

    def setup(self):
        def _setup():
            output_binning = self.data["output_binning"]
            self.data = ContainerSet(self.name)
            self.data["output_binning"] = output_binning
            for stage in self._stages:
                stage.data = self.data
                stage.setup()
        self._timed(_setup, self._setup_times)

    def run(self):
        def _run():
            for stage in self._stages:
                stage.run()
        self._timed(_run, self._run_times)

    def get_outputs(self, output_binning=None, output_key=None):
        def _get():
            original = None
            binning = output_binning
            if binning is None:
                self.run()
                binning = self.output_binning
            elif isinstance(binning, MultiDimBinning):
                original = self.output_binning
                self.output_binning = binning
                self.run()
            key = self.output_key if output_key is None else output_key
            assert isinstance(binning, MultiDimBinning)
            self.data.representation = binning
            if isinstance(key, tuple):
                assert len(key) == 2
                outputs = self.data.get_mapset(key[0], error=key[1])
            else:
                outputs = self.data.get_mapset(key)
            if original is not None:
                self.output_binning = original
            return outputs
        return self._timed(_get, self._get_outputs_times)

    def report_profile(self, detailed=False):
        for label, times in (("- setup:      ", self._setup_times), ("- run:        ", self._run_times),
                             ("- get_outputs:", self._get_outputs_times)):
            if times:
                print(self.name, label, "total %.5f s, n calls: %d" % (sum(times), len(times)))
        for s in self._stages:
            s.report_profile(detailed=detailed)
