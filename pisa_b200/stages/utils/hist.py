"""``utils.hist`` service: events -> weighted histograms (+ sumw2 errors) on the output binning.

Drop-in for pisa/stages/utils/hist.py (reference :27-223): constructor kwargs
``apply_unc_weights, unweighted``; ``expected_params = ()``; ``calc_mode`` "events" (the
binned->binned ``hist_transform`` mode :73-84,131-164 is outside the hot path and raises);
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
            raise NotImplementedError("utils.hist with a binned calc_mode (hist_transform) is outside the "
                                      "pisa_b200 hot path; use calc_mode = events")
        # regularised binning + static per-event bin index (hist.py:86-127; cached on the container)
        self.data["regularized_output_binning"] = self.apply_mode
        for container in self.data:
            container.bin_index(self.apply_mode)

    def apply_function(self):
        n_bins = self.apply_mode.size
        for container in self.data:
            container.representation = "events"
            idx = container.bin_index(self.apply_mode)
            weights = container["weights"]
            if "astro_weights" in container.keys:
                weights = weights + container["astro_weights"]
            if self.unweighted:
                weights = torch.ones_like(weights)
            bin_unc2 = None
            if self.apply_unc_weights:
                unc = container["unc_weights"]
                if self.error_method == "sumw2":
                    bin_unc2, _ = ops.hist_accumulate(idx, (unc * unc * weights).contiguous(), n_bins, want_w2=False)
                weights = (unc * weights).contiguous()
            want_w2 = self.error_method == "sumw2"
            h, h2 = ops.hist_accumulate(idx, weights, n_bins, want_w2=want_w2)
            if want_w2 and bin_unc2 is None:
                bin_unc2 = h   # unc_weights == 1: sum(unc^2 * w) == sum(w)

            container.representation = self.apply_mode
            container["weights"] = h
            # histogramming does not invalidate the "events" representation (hist.py:213)
            container.validity["weights"][hash("events")] = True
            if want_w2:
                container["errors"] = torch.sqrt(h2)
                container["bin_unc2"] = bin_unc2


def init_test(**param_kwargs):
    return hist(calc_mode="events")
