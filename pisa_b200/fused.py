"""A ``Pipeline`` evaluated through the fused template kernel.

``Pipeline.run()`` executes ``osc.prob3 -> aeff.aeff -> utils.hist`` as separate stages: per container a propagation
kernel, two gathers, two multiplies and a histogram pass, with every intermediate (``probability``, ``prob_e``,
``prob_mu``, reweighted ``weights``) written to HBM, and ~70 Python-level container operations per template
(2.4 ms per template however small the sample, profiles/r01_pipeline_timings.txt).  ``FusedPipeline`` keeps the
pipeline object -- its cfg, its ``ParamSet``s, its stages before the oscillation stage (loaders, flux) and after the
histogram stage (``discr_sys.hypersurfaces``) -- and replaces exactly that segment by ONE launch of
``pisab_reweight_hist_batch`` over all containers (``ReweightEngine``), writing ``weights`` / ``errors`` /
``bin_unc2`` into the containers' output-binning representation as ``utils.hist`` would.  ``get_outputs()`` returns the
same ``MapSet`` as ``Pipeline.get_outputs()`` (tests: 1e-10 relative).

Supported shape (checked at construction, ``NotImplementedError`` otherwise): ``osc.prob3`` in events mode, then
optionally ``aeff.aeff``, then ``utils.hist`` with ``calc_mode = events`` and ``error_method = sumw2`` or None; any
stages before and after.  ``utils.hist``'s options (hist.py:141-145,198-209) map onto the fused kernel like this:
  * ``apply_unc_weights``: hist = sum(unc w), sumw2 = sum((unc w)^2) come from one launch over the weights ``unc * w0``;
    ``bin_unc2 = sum(unc^2 w)`` is the "sum w" plane of a second launch over ``unc^2 * w0`` (only with sumw2 errors);
  * ``unweighted``: the weights are ones, nothing has to be propagated: one pass of the histogram kernel per container;
  * ``astro_weights`` (an additive per-event term, hist.py:141-145): handed to the kernel, which adds it to the
    reweighted event weight before histogramming (times ``unc_weights`` / ``unc_weights^2`` like the weights).
A ``flux.barr_simple`` stage directly in front of ``osc.prob3`` is fused as well: the engine keeps the nominal fluxes and
the cached event terms, and a change of the five flux systematics costs no launch of its own -- the template kernel
evaluates them per event in registers (``PISAB_CONTAINER_FLUX_SYS``), ``nu_flux`` is not rewritten.
"""
import numpy as np
import torch

from pisa_b200.core.binning import MultiDimBinning
from pisa_b200.distributed import event_sharding
from pisa_b200.engine import ReweightEngine

__all__ = ["FusedPipeline"]


def _is(stage, stage_name, service_name):
    return stage.stage_name == stage_name and stage.service_name == service_name


class FusedPipeline:
    def __init__(self, pipeline):
        self.pipeline = pipeline
        stages = pipeline.stages
        osc = [i for i, s in enumerate(stages) if _is(s, "osc", "prob3")]
        if len(osc) != 1:
            raise NotImplementedError("FusedPipeline needs exactly one osc.prob3 stage")
        k = osc[0]
        self.pre, self.osc = stages[:k], stages[k]
        self.barr = self.pre.pop() if self.pre and _is(self.pre[-1], "flux", "barr_simple") else None
        rest = stages[k + 1:]
        self.aeff = rest.pop(0) if rest and _is(rest[0], "aeff", "aeff") else None
        if not rest or not _is(rest[0], "utils", "hist"):
            raise NotImplementedError("FusedPipeline needs utils.hist right after osc.prob3 (and an optional aeff.aeff)")
        self.hist, self.post = rest[0], rest[1:]
        if self.osc.calc_mode != "events" or self.osc.apply_mode != "events":
            raise NotImplementedError("FusedPipeline: osc.prob3 must calculate and apply in events mode")
        if self.hist.calc_mode != "events":
            raise NotImplementedError("FusedPipeline: utils.hist must histogram events (calc_mode = events)")
        if self.hist.error_method not in (None, "sumw2"):
            raise NotImplementedError("FusedPipeline: error_method must be None or sumw2")
        pipeline.run()                       # set-up of every stage, bin indices, first template the staged way
        self.binning = self.hist.apply_mode
        assert isinstance(self.binning, MultiDimBinning)
        self._engine = None
        self._engine_unc2 = None             # second engine (weights unc^2 * w0) for bin_unc2 with unc_weights
        self._pre_hash = None
        self._barr_hash = None
        self._scales = None

    # ------------------------------------------------------------------------------------------------
    def _inputs_hash(self):
        return tuple(s.params.values_hash for s in self.pre)

    def _build_engine(self, earth):
        data = self.pipeline.data
        for stage in self.pre:               # loaders reset `weights`, flux stages write `nu_flux`
            stage.run()
        if self.barr is not None:
            self.barr.run()                  # the containers' own nu_flux stays valid for this hypothesis
        engine = engine2 = None
        want_unc2 = self.hist.apply_unc_weights and self.hist.error_method == "sumw2"
        self._containers = list(data.containers)
        for c in self._containers:
            c.representation = "events"
            astro = c["astro_weights"] if "astro_weights" in c.keys else None
            w = c["weights"]
            if self.aeff is not None:
                w = w * c["weighted_aeff"]   # the per-event part of aeff.aeff; its scalar part goes in as `scale`
            if engine is None:
                dt = np.float64 if w.dtype == torch.float64 else np.float32
                engine = ReweightEngine(earth, self.binning.size, dt, w.device)
                engine2 = ReweightEngine(earth, self.binning.size, dt, w.device) if want_unc2 else None
            idx = c.bin_index(self.binning, "hist")
            unc = c["unc_weights"] if self.hist.apply_unc_weights else None
            args = (c.name, int(c["nubar"]), int(c["flav"]), c["true_energy"], c["true_coszen"], c["nu_flux"])
            kw = {}
            if self.barr is not None:
                kw.update(nu_flux_nominal=c["nu_flux_nominal"], nubar_flux_nominal=c["nubar_flux_nominal"])

            def term(x, power):              # x * unc^power (x: the weights or the additive astro term)
                if x is None:
                    return None
                return (x if unc is None else x * unc ** power).contiguous()
            engine.add_container(*args, term(w, 1), idx, astro_weights=term(astro, 1), **kw)
            if want_unc2:
                engine2.add_container(*args, term(w, 2), idx, astro_weights=term(astro, 2), **kw)
        self._engine, self._engine_unc2 = engine, engine2
        self._pre_hash = self._inputs_hash()
        self._barr_hash = None

    def _evaluate_unweighted(self):
        """``unweighted`` (hist.py:141-145): weights of one -- times ``unc_weights`` if asked for -- so the oscillation
        stage does not enter; [containers, 3, bins] = (sum, sumw2, bin_unc2) from the histogram kernel alone."""
        from pisa_b200 import ops
        from pisa_b200.distributed import combine_histograms
        for stage in self.pre:
            stage.run()
        data = self.pipeline.data
        self._containers = list(data.containers)
        rows = []
        for c in self._containers:
            c.representation = "events"
            idx = c.bin_index(self.binning, "hist")
            unc = c["unc_weights"].contiguous() if self.hist.apply_unc_weights else None
            h, h2 = ops.hist_accumulate(idx, unc, self.binning.size)
            rows.append(torch.stack([h, h2, h2]))            # w = unc: sum (unc w)^2 == sum unc^2 w == sum unc^2
        out = torch.stack(rows)
        if event_sharding():
            combine_histograms(out)
        return out

    def _evaluate(self):
        if self.hist.unweighted:
            return self._evaluate_unweighted()
        consts, earth = self.osc.update_hypothesis()
        if self._engine is None or self._inputs_hash() != self._pre_hash:
            self._build_engine(earth)
            self._scales = None
        self._engine.earth = earth
        if self.barr is not None and self.barr.params.values_hash != self._barr_hash:
            p = self.barr.params
            sys = {k: p[k].value.m_as("dimensionless") for k in ("nue_numu_ratio", "nu_nubar_ratio", "delta_index",
                                                                 "Barr_uphor_ratio", "Barr_nu_nubar_ratio")}
            for eng in (self._engine, self._engine_unc2):
                if eng is not None:
                    eng.set_flux_params(**sys)       # no launch: evaluated inside the next template kernel
            self._barr_hash = self.barr.params.values_hash
        if self.aeff is not None:
            scales = [self.aeff.container_scale(c.name) for c in self._containers]
            if scales != self._scales:
                self._engine.set_scales(scales)
                self._scales = scales
                if self._engine_unc2 is not None:
                    self._engine_unc2.set_scales(scales)
        # [containers, 2, bins], one launch; one exchange when the loaders sharded the events over GPUs
        out = self._engine.evaluate(consts, allreduce=event_sharding())
        if self._engine_unc2 is None:
            return out
        self._engine_unc2.earth = earth
        unc2 = self._engine_unc2.evaluate(consts, allreduce=event_sharding())
        return torch.cat([out, unc2[:, :1]], dim=1)          # [containers, 3, bins]: sum, sumw2, bin_unc2

    def run(self):
        """Like ``Pipeline.run()``: afterwards the containers hold the binned ``weights`` (and ``errors``,
        ``bin_unc2`` with sumw2 errors) of this hypothesis, stages after the histogram stage included."""
        out = self._evaluate().clone()                       # the engine reuses its result buffer
        want_w2 = self.hist.error_method == "sumw2"
        errors = torch.sqrt(out[:, 1]) if want_w2 else None
        for i, c in enumerate(self._containers):
            c.representation = self.binning
            c["weights"] = out[i, 0]
            if want_w2:
                c["errors"] = errors[i]
                # (without unc_weights: sum(unc^2 w) == sum(w))
                c["bin_unc2"] = (out[i, 2] if out.shape[1] > 2 else out[i, 0]).clone()
        for stage in self.post:
            stage.run()

    def get_outputs(self, output_binning=None, output_key=None):
        if output_binning is not None or output_key is not None:
            raise NotImplementedError("FusedPipeline.get_outputs: use the pipeline's own output_binning / output_key")
        key = self.pipeline.output_key
        error = key[1] if isinstance(key, tuple) else None
        name = key[0] if isinstance(key, tuple) else key
        if not self.post and name == "weights" and error in (None, "errors"):
            # nothing downstream needs the containers: one device->host copy, maps built directly
            from pisa_b200.core.map import Map, MapSet
            from pisa_b200 import FTYPE
            host = self._evaluate().cpu().numpy()
            shape = self.binning.shape
            # (containers store FTYPE arrays: same rounding as the staged path in FP32 mode)
            maps = [Map(name=c.name, hist=host[i, 0].astype(FTYPE).reshape(shape), binning=self.binning,
                        error_hist=np.sqrt(host[i, 1]).astype(FTYPE).reshape(shape) if error else None)
                    for i, c in enumerate(self._containers)]
            return MapSet(name=self.pipeline.data.name, maps=maps)
        self.run()
        data = self.pipeline.data
        data.representation = self.binning
        return data.get_mapset(name, error=error) if error else data.get_mapset(name)
