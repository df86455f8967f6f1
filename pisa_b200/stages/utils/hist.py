"""``utils.hist`` service: events -> weighted histograms (+ sumw2 errors) on the output binning.

Drop-in for pisa/stages/utils/hist.py (reference :27-223): constructor kwargs
``apply_unc_weights, unweighted``; ``expected_params = ()``; ``calc_mode`` "events", or a binning
disjoint from the output binning: then a per-container ``hist_transform`` [calc bins, output bins]
(event counts on the joint binning, :69-84) is built at setup and applied as a matrix product (:131-160);
``apply_mode`` defaults to ``data["output_binning"]`` (:64-67); with ``error_method == 'sumw2'`` the
stage also writes ``errors = sqrt(sum w^2)`` and ``bin_unc2`` (:205-218).

``setup_function`` classifies every output dimension like :86-127 (irregular -> searchsorted on the
real edges, log -> linear bins in log x, else linear) and caches one flat int32 bin index per
container on the device; ``apply_function`` is then ONE pass of the deterministic histogram kernel
per container producing sum w and sum w^2 together (the reference makes three fast_histogram
passes), reading 12 B/event.
"""
import torch

from pisa_b200 import ops
from pisa_b200.core.binning import MultiDimBinning
from pisa_b200.core.stage import Stage
from pisa_b200.distributed import combine_histograms, event_sharding

__all__ = ["hist", "init_test"]


class hist(Stage):  # pylint: disable=invalid-name
    def __init__(self, apply_unc_weights=False, unweighted=False, **std_kwargs):
        expected_container_keys = ["weights"]
        if apply_unc_weights:
            expected_container_keys.append("unc_weights")
        supported_reps = {"calc_mode": [MultiDimBinning, "events"], "apply_mode": [None, MultiDimBinning]}
        super().__init__(expected_params=(), expected_container_keys=expected_container_keys,
                         supported_reps=supported_reps, **std_kwargs)
        self.apply_unc_weights = apply_unc_weights
        self.unweighted = unweighted

    def setup_function(self):
        if self.apply_mode is None:
            self.apply_mode = self.data["output_binning"]
        else:
            assert self.apply_mode == self.data["output_binning"]
        if isinstance(self.calc_mode, MultiDimBinning):
            # the two binnings must be exclusive (hist.py:71-72)
            assert len(set(self.calc_mode.names) & set(self.apply_mode.names)) == 0
            # event counts on the joint binning calc_mode + apply_mode (hist.py:73-84): the joint flat index comes from
            # the two cached sub-indices, so the usual 2 + 3 dimensional layout is not limited by PISAB_MAX_DIMS.
            # Both sub-indices follow the edge rule of the reference's histogram on the joint binning
            # (translation.histogram: fast_histogram for linear dimensions, np.histogramdd otherwise).
            n_joint = self.calc_mode.size * self.apply_mode.size
            if n_joint > 2 ** 31 - 1:
                raise ValueError("hist_transform of %d x %d bins exceeds the 32-bit bin index"
                                 % (self.calc_mode.size, self.apply_mode.size))
            joint_dims = list(self.calc_mode) + list(self.apply_mode)
            policy = "edges" if any(d.is_irregular or d.is_log for d in joint_dims) else "generic"
            for container in self.data:
                container.representation = "events"
                joint = ops.joint_index(container.bin_index(self.calc_mode, policy),
                                        container.bin_index(self.apply_mode, policy), self.apply_mode.size)
                counts, _ = ops.hist_accumulate(joint, None, n_joint, want_w2=False)
                container.representation = self.calc_mode
                container["hist_transform"] = counts.reshape(self.calc_mode.size, self.apply_mode.size)
            return
        # regularised binning + static per-event bin index (hist.py:86-127; cached on the container)
        self.data["regularized_output_binning"] = self.apply_mode
        for container in self.data:
            container.bin_index(self.apply_mode, "hist")
            container.bin_plan(self.apply_mode, "hist")      # bin-sorted tiles: the per-template pass reads 10 B/event

    def _apply_transform(self):
        """calc_mode binned: hist = (unc * w) @ hist_transform (hist.py:131-160)."""
        if self.unweighted:
            raise NotImplementedError("Unweighted hist only implemented in event-wise calculation")
        local = []
        for container in self.data:
            container.representation = self.calc_mode
            weights = container["weights"]
            if "astro_weights" in container.keys:
                weights = weights + container["astro_weights"]
            unc = container["unc_weights"] if self.apply_unc_weights else None
            h, sumw2, bin_unc2 = ops.hist_transform(weights.contiguous(), unc, container["hist_transform"],
                                                    want_errors=self.error_method == "sumw2")
            local.append((container, h, sumw2, bin_unc2))
        self._write(self._exchange(local), revalidate_events=False)

    def _exchange(self, local):
        """Events sharded over GPUs (pisa_b200.distributed.enable_event_sharding): the single exchange of the path --
        all containers' (sum w, sum w^2, bin_unc2) in ONE buffer, rank-ordered sum, identical on every rank."""
        if not event_sharding() or not local:
            return local
        zeros = torch.zeros_like(local[0][1])
        buf = torch.stack([torch.stack([h, zeros if s is None else s, zeros if b is None else b])
                           for _, h, s, b in local]).to(torch.float64)
        combine_histograms(buf)
        return [(c, buf[i, 0], None if s is None else buf[i, 1], None if b is None else buf[i, 2])
                for i, (c, _, s, b) in enumerate(local)]

    def _write(self, results, revalidate_events):
        for container, h, sumw2, bin_unc2 in results:
            container.representation = self.apply_mode
            container["weights"] = h
            if revalidate_events:
                # histogramming does not invalidate the "events" representation (hist.py:213)
                container.validity["weights"][hash("events")] = True
            if self.error_method == "sumw2":
                container["errors"] = torch.sqrt(sumw2)
                container["bin_unc2"] = bin_unc2

    def apply_function(self):
        if isinstance(self.calc_mode, MultiDimBinning):
            return self._apply_transform()
        n_bins = self.apply_mode.size
        local = []
        for container in self.data:
            container.representation = "events"
            idx = container.bin_index(self.apply_mode, "hist")
            plan = container.bin_plan(self.apply_mode, "hist")
            weights = container["weights"]
            if "astro_weights" in container.keys:
                weights = weights + container["astro_weights"]
            if self.unweighted:
                weights = torch.ones_like(weights)
            bin_unc2 = None
            if self.apply_unc_weights:
                unc = container["unc_weights"]
                if self.error_method == "sumw2":
                    bin_unc2, _ = ops.hist_accumulate(idx, (unc * unc * weights).contiguous(), n_bins, want_w2=False,
                                                      plan=plan)
                weights = (unc * weights).contiguous()
            want_w2 = self.error_method == "sumw2"
            h, h2 = ops.hist_accumulate(idx, weights, n_bins, want_w2=want_w2, plan=plan)
            if want_w2 and bin_unc2 is None:
                bin_unc2 = h.clone()   # unc_weights == 1: sum(unc^2 * w) == sum(w); its own array, like the reference's
            local.append((container, h, h2, bin_unc2))
        self._write(self._exchange(local), revalidate_events=True)


def init_test(**param_kwargs):
    return hist(calc_mode="events")
