"""``data.simple_data_loader`` service: PISA-style event files -> device-resident containers.

Drop-in for pisa/stages/data/simple_data_loader.py (reference :20-299) with the loading logic of
pisa/core/events_pi.py::EventsPi.load_events_file / apply_cut (reference :175-505, :510-600) that it relies on:
constructor kwargs ``events_file, mc_cuts, data_dict, neutrinos=True, required_metadata=None,
fraction_events_to_keep=None, events_subsample_index=0, seed=123456, output_names=None``; ``calc_mode`` is not
accepted, ``apply_mode`` is "events" (:108-113); ``data_dict`` maps container keys to file variables, a list of
variables is stacked into a 2-d array (events_pi.py:420-448); ``mc_cuts`` is a numpy boolean expression over the
container keys (:510-600); ``fraction_events_to_keep`` / ``events_subsample_index`` / ``seed`` select one of the
statistically independent sub-samples with the reference's algorithm (:455-500) and scale ``initial_weights`` by the
inverse fraction (:218-224); a ``weights`` field in the file is an error (:203-209); ``apply_function`` resets
``weights`` to ``initial_weights`` (:245-252).

On-disk formats.  The reference reads HDF5 through ``h5py``: one group per event category (``nue_cc`` ...
``nutaubar_nc``), one dataset per variable, metadata in the file attributes.  ``h5py`` is not installed in this image,
so the same layout is also accepted as a NumPy ``.npz`` archive whose keys are ``"<category>/<variable>"`` (plus
optional ``"__metadata__/<key>"`` scalars); ``.hdf5`` / ``.h5`` files are read when ``h5py`` is importable and raise an
ImportError that says so otherwise.  A mapping ``{category: {variable: array}}`` is accepted like in the reference.
Files that are not yet split by flavour and interaction are not supported (the reference's legacy formats).

Every column is uploaded once and stays in HBM as an SoA array of FTYPE; with event sharding on
(``pisa_b200.distributed``) each rank keeps its contiguous slice of every category.
"""
import ast
from collections import OrderedDict
from collections.abc import Mapping

import numpy as np

from pisa_b200 import FTYPE
from pisa_b200.core.container import Container
from pisa_b200.core.stage import Stage
from pisa_b200.distributed import local_slice
from pisa_b200.utils.resources import find_resource

__all__ = ["simple_data_loader", "read_events_file", "load_events", "apply_cut", "init_test"]


def _split(value):
    if value is None:
        return []
    if isinstance(value, (list, tuple)):
        return [str(v).strip() for v in value]
    return [v.strip() for v in str(value).split(",") if v.strip()]


def read_events_file(path):
    """(``{category: {variable: ndarray}}``, metadata dict) from one events file (``.npz`` or HDF5)."""
    if isinstance(path, Mapping):
        return OrderedDict((k, OrderedDict(v)) for k, v in path.items()), dict(getattr(path, "metadata", {}) or {})
    path = find_resource(path)
    data, meta = OrderedDict(), {}
    if path.endswith(".npz"):
        with np.load(path, allow_pickle=False) as f:
            for key in f.files:
                cat, _, var = key.partition("/")
                if not var:
                    raise ValueError('%s: key "%s" is not of the form "<category>/<variable>"' % (path, key))
                if cat == "__metadata__":
                    meta[var] = f[key].item() if f[key].ndim == 0 else f[key]
                else:
                    data.setdefault(cat, OrderedDict())[var] = f[key]
    elif path.endswith((".hdf5", ".h5", ".hdf")):
        try:
            import h5py
        except ImportError as exc:
            raise ImportError("reading %s needs h5py, which is not installed; convert the file to the .npz layout "
                              '("<category>/<variable>" keys) described in simple_data_loader.py' % path) from exc
        with h5py.File(path, "r") as f:
            meta.update({k: (v.item() if hasattr(v, "item") and np.ndim(v) == 0 else v) for k, v in f.attrs.items()})
            for cat, grp in f.items():
                if not isinstance(grp, h5py.Group):
                    raise ValueError("%s: top-level entry %s is not a group of event variables" % (path, cat))
                data[cat] = OrderedDict((var, np.asarray(ds)) for var, ds in grp.items())
    else:
        raise ValueError("unknown events file format: %s (expected .npz, .hdf5 or .h5)" % path)
    if not data:
        raise ValueError("No input data found")
    return data, meta


def _subsample_indices(n_values, fraction, subsample_index, rand):
    """events_pi.py:455-500: draw `subsample_index` + 1 disjoint random sub-samples, keep the last one."""
    desired = int(fraction * float(n_values))
    current = np.arange(n_values)
    i = 0
    while True:
        assert current.size >= desired, "Not enough events available"
        chosen = np.sort(rand.choice(current, replace=False, size=desired))
        if i == subsample_index:
            return chosen
        current = np.sort(np.setxor1d(current, chosen))
        i += 1


def load_events(events_file, data_dict=None, neutrinos=True, required_metadata=None, fraction_events_to_keep=None,
                events_subsample_index=0, seed=123456):
    """Host-side part of the loader: ``(OrderedDict category -> OrderedDict key -> FTYPE ndarray, metadata)``."""
    if fraction_events_to_keep is not None:
        fraction_events_to_keep = float(fraction_events_to_keep)
        if not 0.0 <= fraction_events_to_keep <= 1.0:
            raise ValueError("`fraction_events_to_keep` must be in [0, 1]")
        if events_subsample_index < 0 or (events_subsample_index + 1) * fraction_events_to_keep > 1.0 + 1e-12:
            raise ValueError("`events_subsample_index` = %d is not available at a fraction of %g"
                             % (events_subsample_index, fraction_events_to_keep))
    if data_dict is not None:
        if not isinstance(data_dict, Mapping):
            raise TypeError("'variable_mapping' must be a mapping (e.g., dict)")
        for dst, src in data_dict.items():
            if not isinstance(dst, str):
                raise TypeError("`variable_mapping` 'dst' (key) must be a string")
            if not isinstance(src, str) and not all(isinstance(v, str) for v in src):
                raise TypeError("`variable_mapping` 'src' (value) must be a string or an iterable of strings")
    files = [events_file] if isinstance(events_file, (str, Mapping)) else list(events_file)
    raw, metadata = OrderedDict(), {}
    for f in files:
        data, meta = read_events_file(f)
        for cat, variables in data.items():                      # several files: events are appended per variable
            dst = raw.setdefault(cat, OrderedDict())
            for var, arr in variables.items():
                dst[var] = np.concatenate([dst[var], arr]) if var in dst else np.asarray(arr)
        for k in (required_metadata or []):
            if k not in meta:
                raise AssertionError("Expected metadata '%s' not found" % k)
            if k in metadata:
                if k == "livetime":
                    metadata[k] += meta[k]
                else:
                    assert metadata[k] == meta[k]
            else:
                metadata[k] = meta[k]
    if neutrinos:
        bad = [c for c in raw if not c.endswith(("_cc", "_nc"))]
        if bad:
            raise ValueError("event categories %s are not split by flavour and interaction (<flavour>_cc / _nc); the "
                             "reference's legacy joined formats are not supported" % bad)
    out = OrderedDict()
    for cat, variables in raw.items():
        mapping = tuple(zip(variables, variables)) if data_dict is None else tuple(data_dict.items())
        chosen = None
        rand = np.random.RandomState(seed)                       # the same sample each time (events_pi.py:424)
        out[cat] = OrderedDict()
        for dst, src in mapping:
            stack = []
            for var in ([src] if isinstance(src, str) else list(src)):
                if var not in variables:
                    raise KeyError("Variable '%s' cannot be found for '%s' events" % (var, cat))
                stack.append(np.asarray(variables[var]).astype(FTYPE))
            arr = np.squeeze(np.stack(stack, axis=1))
            if fraction_events_to_keep is not None:
                if chosen is None:
                    chosen = _subsample_indices(arr.size, fraction_events_to_keep, events_subsample_index, rand)
                arr = arr[chosen]
            out[cat][dst] = np.ascontiguousarray(arr)
    return out, metadata


def apply_cut(events, keep_criteria):
    """``EventsPi.apply_cut`` (events_pi.py:510-600): keep the events of every category for which the numpy boolean
    expression over its variables holds (``np`` is available in the expression)."""
    assert isinstance(keep_criteria, str)
    out = OrderedDict()
    for cat, variables in events.items():
        # the variables of the category are the names of the expression (the reference substitutes them textually)
        mask = eval(keep_criteria, {"np": np}, dict(variables))   # noqa: S307  (same contract as the reference: a cfg string)
        mask = np.asarray(mask, dtype=bool)
        out[cat] = OrderedDict((k, np.ascontiguousarray(v[mask])) for k, v in variables.items())
    return out


class simple_data_loader(Stage):  # pylint: disable=invalid-name
    def __init__(self, events_file, mc_cuts, data_dict, neutrinos=True, required_metadata=None,
                 fraction_events_to_keep=None, events_subsample_index=0, seed=123456, output_names=None,
                 **std_kwargs):
        self.events_file = events_file if isinstance(events_file, Mapping) else _split(events_file)
        self.mc_cuts = mc_cuts
        if isinstance(data_dict, str):
            data_dict = ast.literal_eval(data_dict)
        self.data_dict = data_dict
        self.neutrinos = neutrinos
        self.required_metadata = _split(required_metadata) if required_metadata is not None else None
        self.fraction_events_to_keep = None if fraction_events_to_keep in (None, "None") else float(fraction_events_to_keep)
        self.events_subsample_index = int(events_subsample_index)
        self.seed = int(seed)
        self.output_names = _split(output_names)
        if len(self.output_names) != len(set(self.output_names)):
            raise ValueError("Found duplicates in `output_names`, but each name must be unique.")
        super().__init__(expected_params=(), expected_container_keys=(),
                         supported_reps={"calc_mode": None, "apply_mode": "events"}, **std_kwargs)
        self.evts, self.metadata = load_events(self.events_file, self.data_dict, self.neutrinos, self.required_metadata,
                                               self.fraction_events_to_keep, self.events_subsample_index, self.seed)
        if self.mc_cuts:
            self.evts = apply_cut(self.evts, self.mc_cuts)

    def setup_function(self):
        """``record_event_properties`` (:166-238)."""
        for name in (self.output_names or list(self.evts)):
            if name not in self.evts:
                raise ValueError('Output name "%s" not found in events. Only found %s.' % (name, list(self.evts)))
            container = Container(name)
            container.representation = "events"
            n = None
            for key, val in self.evts[name].items():
                sl = local_slice(len(val))                       # this rank's share when events are sharded over GPUs
                container[key] = val[sl]
                n = len(val[sl])
            if "weights" in container.keys:
                raise KeyError('Found an existing `weights` array in "%s" which would be overwritten. Consider renaming '
                               "it to `initial_weights`." % name)
            container["weights"] = np.ones(n, dtype=FTYPE)
            if "initial_weights" not in container.keys:
                scale = 1.0
                if self.fraction_events_to_keep is not None and ("nu" in name or "mu" in name):
                    scale = 1.0 / float(self.fraction_events_to_keep)   # down-sampling: keep the normalisation
                container["initial_weights"] = np.full(n, scale, dtype=FTYPE)
            if self.neutrinos:
                nubar = -1 if "bar" in name else 1
                if name.startswith("nutau"):
                    flav = 2
                elif name.startswith("numu"):
                    flav = 1
                elif name.startswith("nue"):
                    flav = 0
                else:
                    raise ValueError("Cannot determine flavour of %s" % name)
                container.set_aux_data("nubar", nubar)
                container.set_aux_data("flav", flav)
            self.data.add_container(container)
        if len(self.data.names) == 0:
            raise ValueError("No containers created during data loading for some reason.")

    def apply_function(self):
        for container in self.data:
            container["weights"] = container["initial_weights"].clone()


def init_test(**param_kwargs):
    """Initialisation example (simple_data_loader.py:255-273): the reference points at a bundled HDF5 file; here a
    10-event in-memory sample with the same variables."""
    rng = np.random.RandomState(0)
    names = ["nue_cc", "numu_cc", "nutau_cc", "nuebar_cc", "numubar_cc", "nutaubar_cc",
             "nue_nc", "numu_nc", "nutau_nc", "nuebar_nc", "numubar_nc", "nutaubar_nc"]
    variables = ["true_energy", "true_coszen", "reco_energy", "reco_coszen", "pid", "weighted_aeff", "nominal_nue_flux",
                 "nominal_numu_flux", "nominal_nuebar_flux", "nominal_numubar_flux"]
    events = OrderedDict()
    for name in names:
        ev = OrderedDict((v, rng.rand(10)) for v in variables)
        ev["true_energy"] = 1.0 + 79.0 * ev["true_energy"]
        ev["reco_energy"] = 1.0 + 79.0 * ev["reco_energy"]
        ev["true_coszen"] = 2.0 * ev["true_coszen"] - 1.0
        ev["reco_coszen"] = 2.0 * ev["reco_coszen"] - 1.0
        events[name] = ev
    return simple_data_loader(
        events_file=events, mc_cuts="(true_coszen <= 0.5) & (true_energy <= 70)",
        data_dict={"true_energy": "true_energy", "true_coszen": "true_coszen", "reco_energy": "reco_energy",
                   "reco_coszen": "reco_coszen", "pid": "pid", "weighted_aeff": "weighted_aeff",
                   "nu_flux_nominal": ["nominal_nue_flux", "nominal_numu_flux"],
                   "nubar_flux_nominal": ["nominal_nuebar_flux", "nominal_numubar_flux"]},
        output_names=names)
