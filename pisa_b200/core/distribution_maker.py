"""``DistributionMaker``: the caller of the hot path inside a fit.

Minimal counterpart of pisa/core/distribution_maker.py (reference :40-470) for what a minimiser iteration drives:
one or several pipelines, ``get_outputs(return_sum=True)`` (:251-294: every pipeline's MapSet summed to ONE map inside
a MapSet), ``update_params`` / ``select_params`` across pipelines (:296-325), the union of the pipelines' params, and
the rescaled free-parameter interface a minimiser talks to (:399-441).  Pipelines whose stage order allows it
(``osc.prob3 [-> aeff.aeff] -> utils.hist``, events mode) are evaluated through ``pisa_b200.fused.FusedPipeline`` --
one fused launch per template instead of one kernel per stage and container -- unless ``fused=False``; the returned
``MapSet`` is the same either way (tests: 1e-10).  Everything of the reference's class that concerns detectors,
covariances, profiling tables or hashing of source code is out of scope.
"""
from collections import OrderedDict

from pisa_b200.core.map import MapSet
from pisa_b200.core.param import ParamSet
from pisa_b200.core.pipeline import Pipeline

__all__ = ["DistributionMaker"]


class DistributionMaker:
    def __init__(self, pipelines, label=None, set_livetime_from_data=True, profile=False, fused=True):
        self.label = label
        self.metadata = OrderedDict()
        self._profile = profile
        if isinstance(pipelines, (str, OrderedDict, dict, Pipeline)):
            pipelines = [pipelines]
        self._pipelines = []
        for p in pipelines:
            if not isinstance(p, Pipeline):
                p = Pipeline(p, profile=profile)
            self._pipelines.append(p)
        # one evaluator per pipeline: fused where the stage order allows it
        self._evaluators = []
        for p in self._pipelines:
            ev = p
            if fused:
                try:
                    from pisa_b200.fused import FusedPipeline
                    ev = FusedPipeline(p)
                except NotImplementedError:
                    ev = p
            self._evaluators.append(ev)
        if set_livetime_from_data:
            livetime = None
            for ip, p in enumerate(self._pipelines):
                for istage, stage in enumerate(p.stages):
                    meta = getattr(stage, "metadata", None)
                    if not (isinstance(meta, dict) and "livetime" in meta):
                        continue
                    if livetime is None:
                        livetime = meta["livetime"]
                    if meta["livetime"] != livetime:
                        raise ValueError("Pipeline index %d, stage index %d has data livetime = %s, in disagreement with "
                                         "previously-found livetime = %s" % (ip, istage, meta["livetime"], livetime))

    def __iter__(self):
        return iter(self._pipelines)

    pipelines = property(lambda self: self._pipelines)
    evaluators = property(lambda self: self._evaluators)

    def run(self):
        for ev in self._evaluators:
            ev.run()

    def setup(self):
        for p in self._pipelines:
            p.setup()

    def get_outputs(self, return_sum=False, sum_map_name="total", sum_map_tex_name="Total", **kwargs):
        """List of every pipeline's ``MapSet``; with ``return_sum`` ONE ``MapSet`` holding the sum of all maps of all
        pipelines (distribution_maker.py:251-294)."""
        outputs = [ev.get_outputs(**kwargs) for ev in self._evaluators]
        if not return_sum:
            return outputs
        total = sum(sum(ms) for ms in outputs)        # Map.__radd__ takes the integer start value
        total.name, total.tex = sum_map_name, sum_map_tex_name
        return MapSet([total], name=sum_map_name)

    # ---------------------------------------------------------------------------- params ------
    @property
    def params(self):
        ps = ParamSet()
        for p in self._pipelines:
            ps.extend(p.params)
        return ps

    @property
    def param_selections(self):
        return sorted({s for p in self._pipelines for s in p.param_selections})

    def update_params(self, params):
        for p in self._pipelines:
            p.update_params(params)

    def select_params(self, selections, error_on_missing=True):
        """A pipeline that lacks one of the selections keeps its params (distribution_maker.py:300-325)."""
        if selections is None:
            return
        successes = 0
        for p in self._pipelines:
            try:
                p.select_params(selections, error_on_missing=True)
            except KeyError:
                pass
            else:
                successes += 1
        if error_on_missing and successes == 0:
            raise KeyError("None of the stages from any pipeline in this distribution maker has all of the selections "
                           "%s available." % (selections,))

    def set_free_params(self, values):
        """Values (Quantities or magnitudes in the params' own units) for the free params, in ``params.free`` order."""
        free = self.params.free
        if len(values) != len(free):
            raise ValueError("expected %d values for the free params, got %d" % (len(free), len(values)))
        for prm, v in zip(free, values):
            prm.value = v if hasattr(v, "magnitude") else v * prm.value.units
        self.update_params(free)
