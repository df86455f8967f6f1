"""``data.csv_loader`` service: CSV data-release events -> device-resident containers.

Drop-in for pisa/stages/data/csv_loader.py (reference :20-172): constructor kwargs ``events_file, data_dict,
output_names, neutrinos=True, dis_idx=None, scale_aeff=False`` (:58-66); per output name the events are
selected by PDG code ``nubar * (12 + 2 flav)`` from the ``pdg_code`` / ``pdg`` column and by ``type >= 1`` (CC) or
``type == 0`` (NC) (:121-139); ``initial_weights`` / ``weights`` are ones (:143-144); ``data_dict`` maps container
keys to CSV columns (:145-146); ``scale_aeff`` converts cm^2 -> m^2 (:148-149); ``dis_idx`` derives the ``dis``
flag from ``interaction`` (:152-153).  ``apply_function`` resets the weights (:164-167).

The CSV is parsed once on the host (pandas, like the reference); every column is then uploaded once and
stays in HBM as an SoA array of FTYPE -- the layout the kernels read.
"""
import ast

import numpy as np

from pisa_b200 import FTYPE
from pisa_b200.core.container import Container
from pisa_b200.core.stage import Stage
from pisa_b200.distributed import local_slice
from pisa_b200.utils.resources import find_resource

__all__ = ["csv_loader", "select_events", "init_test"]


def _split(value):
    if isinstance(value, (list, tuple)):
        return [str(v).strip() for v in value]
    return [v.strip() for v in str(value).split(",")]


def species_of(name):
    """(nubar, flav) from a container name like the reference does (:114-121)."""
    nubar = -1 if "bar" in name else 1
    flav = None
    if "e" in name:
        flav = 0
    if "mu" in name:
        flav = 1
    if "tau" in name:
        flav = 2
    return nubar, flav


def select_events(raw_data, name, neutrinos=True):
    """Rows of ``raw_data`` (a pandas DataFrame) that belong to container ``name`` (:112-141)."""
    if not neutrinos:
        return raw_data
    nubar, flav = species_of(name)
    pdg = nubar * (12 + 2 * flav)
    if "pdg_code" in raw_data:
        mask = raw_data["pdg_code"] == pdg
    elif "pdg" in raw_data:
        mask = raw_data["pdg"] == pdg
    else:
        raise ValueError("Either 'pdg' or 'pdg_code' must be in file.")
    if "cc" in name:
        mask = np.logical_and(mask, raw_data["type"] >= 1)
    else:
        mask = np.logical_and(mask, raw_data["type"] == 0)
    return raw_data[mask]


class csv_loader(Stage):  # pylint: disable=invalid-name
    def __init__(self, events_file, data_dict, output_names, neutrinos=True, dis_idx=None, scale_aeff=False,
                 **std_kwargs):
        self.events_file = [find_resource(f) for f in _split(events_file)]
        if isinstance(data_dict, str):
            self.data_dict = ast.literal_eval(data_dict)
        elif isinstance(data_dict, dict):
            self.data_dict = data_dict
        else:
            raise ValueError("Unsupported type %s for data_dict." % type(data_dict))
        self.output_names = _split(output_names)
        if len(self.output_names) != len(set(self.output_names)):
            raise ValueError("Found duplicates in `output_names`, but each name must be unique.")
        self.neutrinos = neutrinos
        self.dis_idx = int(dis_idx) if dis_idx is not None else None
        self.scale_aeff = scale_aeff
        super().__init__(expected_params=(), expected_container_keys=(),
                         supported_reps={"calc_mode": "events", "apply_mode": "events"}, **std_kwargs)

    def setup_function(self):
        import pandas as pd
        raw_data = pd.concat([pd.read_csv(f) for f in self.events_file])
        for name in self.output_names:
            container = Container(name)
            if self.neutrinos:
                nubar, flav = species_of(name)
                container.set_aux_data("nubar", nubar)
                container.set_aux_data("flav", flav)
            events = select_events(raw_data, name, self.neutrinos)
            events = events.iloc[local_slice(len(events))]      # this rank's share when events are sharded over GPUs
            container["initial_weights"] = np.ones(len(events), dtype=FTYPE)
            container["weights"] = np.ones(len(events), dtype=FTYPE)
            for key, val in self.data_dict.items():
                container[key] = np.ascontiguousarray(events[val].values.astype(FTYPE))
            if self.scale_aeff and "weighted_aeff" in container.keys:
                container["weighted_aeff"] = container["weighted_aeff"] * 1.0e-4
            if "dis" not in container.keys and "interaction" in container.keys and self.dis_idx is not None:
                container["dis"] = (container["interaction"] == self.dis_idx).to(container["interaction"].dtype)
            self.data.add_container(container)
        if len(self.data.names) == 0:
            raise ValueError("No containers created during data loading for some reason.")

    def apply_function(self):
        for container in self.data:
            container["weights"] = container["initial_weights"].clone()


def init_test(**param_kwargs):
    """Initialisation example (csv_loader.py:170-182)."""
    data_dict = {"true_energy": "true_energy", "true_coszen": "true_coszen", "weighted_aeff": "weight",
                 "reco_energy": "reco_energy", "reco_coszen": "reco_coszen", "pid": "pid"}
    return csv_loader(events_file="events/IceCube_3y_oscillations/neutrino_mc.csv.bz2", data_dict=data_dict,
                      output_names=["nue_cc", "numu_cc"])
