"""``discr_sys.hypersurfaces`` service: per-bin detector-systematics scale factors.

Drop-in for pisa/stages/discr_sys/hypersurfaces.py (reference :49-243) restricted to what the shipped IceCube-3y
pipeline uses: the data-release CSV hyperplanes (``fit_results_file = .../hyperplanes_*.csv.bz2``, loaded like
``_load_hypersurfaces_data_release``, pisa/utils/hypersurface/hypersurface.py:2065-2172): four files
(``nue_cc``, ``numu_cc``, ``nutau_cc``, ``all_nc``) -> maps ``nue_cc+nuebar_cc`` ... ``nu_nc+nubar_nc``; every
column that is neither a binning dimension nor ``offset`` is a systematic parameter with a *linear* functional
form, evaluated without nominal shift (``using_legacy_data``, :427):

    scale[bin] = offset[bin] + sum_p gradient_p[bin] * value_p              (:421-428, linear: f(p) = m p :84-93)

``expected_params`` are those column names (:104,120).  ``links`` joins the containers that share a hyperplane
(:128-133,146-148).  ``compute_function`` writes ``hs_scales`` (non-finite -> 1, :196-203); ``apply_function``
scales ``errors`` and ``bin_unc2`` and then ``weights`` (clipped at 0) (:219-243).  Interpolated hypersurfaces,
fluctuations and uncertainty propagation of the fitted JSON format are not built (they need the fit machinery of
pisa/utils/hypersurface, outside the hot path): requesting them raises.

The arrays involved have one entry per analysis bin (128), so this stage is a handful of element-wise device
operations on tiny tensors (torch, plumbing) -- there is nothing to accelerate.
"""
import ast
from collections import OrderedDict
from collections.abc import Mapping

import numpy as np
import torch

from pisa_b200.core.binning import MultiDimBinning
from pisa_b200.core.stage import Stage
from pisa_b200.utils.resources import find_resource

__all__ = ["hypersurfaces", "load_hypersurfaces_data_release", "evaluate_hyperplane"]

_FILES = OrderedDict([("nue_cc+nuebar_cc", "nue_cc"), ("numu_cc+numubar_cc", "numu_cc"),
                      ("nutau_cc+nutaubar_cc", "nutau_cc"), ("nu_nc+nubar_nc", "all_nc")])


def load_hypersurfaces_data_release(input_file_prototype, binning):
    """{map name: {"offset": ndarray[binning.shape], "gradients": {param: ndarray[binning.shape]}}}."""
    import pandas as pd
    assert binning is not None, "Must provide binning when loading data release hypersurfaces"
    out = OrderedDict()
    param_names = None
    for map_name, tag in _FILES.items():
        table = pd.read_csv(find_resource(input_file_prototype.replace("*", tag)))
        for n in binning.names:
            midpoints_found = np.unique(table.pop(n).values)
            assert midpoints_found.size == binning[n].num_bins, "Mismatch between expected and actual binning dimensions"
        offset = table.pop("offset")
        if param_names is None:
            param_names = table.columns.tolist()
        else:
            assert param_names == table.columns.tolist(), "Mismatch between hypersurface params in different files"
        out[map_name] = dict(offset=offset.values.reshape(binning.shape).astype(np.float64),
                             gradients=OrderedDict((p, table[p].values.reshape(binning.shape).astype(np.float64))
                                                   for p in param_names))
    return out, param_names


def evaluate_hyperplane(surface, param_values):
    """offset + sum_p gradient_p * value_p over all bins (hypersurface.py:421-428 with linear terms)."""
    scales = surface["offset"].copy()
    for name, grad in surface["gradients"].items():
        scales += grad * param_values[name]
    return scales


class hypersurfaces(Stage):  # pylint: disable=invalid-name
    def __init__(self, fit_results_file, propagate_uncertainty=False, interpolated=False, links=None,
                 fluctuate=False, fluctuate_seed=None, **std_kwargs):
        if propagate_uncertainty or interpolated or fluctuate:
            raise NotImplementedError("only data-release hyperplanes without uncertainty propagation, interpolation "
                                      "or fluctuation are supported by pisa_b200")
        if ".csv" not in str(fit_results_file):
            raise NotImplementedError("only the data-release CSV hyperplanes are supported by pisa_b200")
        self.fit_results_file = fit_results_file
        self.propagate_uncertainty = False
        calc_mode = std_kwargs.get("calc_mode")
        if not isinstance(calc_mode, MultiDimBinning):
            raise ValueError("discr_sys.hypersurfaces needs a binned calc_mode")
        self.hypersurfaces, self.hypersurface_param_names = load_hypersurfaces_data_release(fit_results_file, calc_mode)
        expected_container_keys = ["weights"]
        if std_kwargs.get("error_method"):
            expected_container_keys.append("errors")
        super().__init__(expected_params=tuple(self.hypersurface_param_names),
                         expected_container_keys=expected_container_keys,
                         supported_reps={"calc_mode": MultiDimBinning}, **std_kwargs)
        if links is None:
            self.links = {}
        elif not isinstance(links, Mapping):
            self.links = ast.literal_eval(links)
        else:
            self.links = links

    def _link(self):
        for key, val in self.links.items():
            self.data.link_containers(key, val)

    def setup_function(self):
        self._link()
        for container in self.data:
            assert container.name in self.hypersurfaces, "No match for map %s found in the hypersurfaces" % container.name
            container["hs_scales"] = np.ones(container.size)
        self.data.unlink_containers()

    def compute_function(self):
        self._link()
        param_values = {name: float(self.params[name].m) for name in self.hypersurface_param_names}
        for container in self.data:
            scales = evaluate_hyperplane(self.hypersurfaces[container.name], param_values).reshape(container.size)
            scales[~np.isfinite(scales)] = 1.0
            container["hs_scales"] = scales
            container.mark_changed("hs_scales")
        self.data.unlink_containers()

    def apply_function(self):
        for container in self.data:
            scales = container["hs_scales"]
            if self.error_method == "sumw2":
                container["errors"] = container["errors"] * scales
                container.mark_changed("errors")
                if "bin_unc2" in container.keys:
                    container["bin_unc2"] = torch.clamp(container["bin_unc2"] * scales, min=0.0)
                    container.mark_changed("bin_unc2")
            container["weights"] = torch.clamp(container["weights"] * scales, min=0.0)
