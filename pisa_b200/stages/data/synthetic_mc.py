"""``data.synthetic_mc`` service: synthetic stand-in for ``data.csv_loader`` on the IceCube-3y MC.

NOT in the reference: the bundled MC file of the IceCube-3y example
(``events/IceCube_3y_oscillations/neutrino_mc.csv.bz2``, read by pisa/stages/data/csv_loader.py
:108-170) is absent from the reference tree (.MISSING_LARGE_BLOBS), so pipelines of that shape run
on events drawn from the laws of SURVEY.md 8d with the same container keys the loader provides
(``true_energy, true_coszen, reco_energy, reco_coszen, pid, weighted_aeff, initial_weights, weights``,
aux ``nubar`` / ``flav``) plus ``nu_flux`` (the flux stages are "next" rows of the scope table).
Events are generated on the device.  ``apply_function`` resets the weights like csv_loader.py:166-170.
"""
from pisa_b200 import FTYPE
from pisa_b200.core.container import Container, default_device
from pisa_b200.core.stage import Stage
from pisa_b200.distributed import local_slice
from pisa_b200.utils import synthetic as syn

__all__ = ["synthetic_mc"]


class synthetic_mc(Stage):  # pylint: disable=invalid-name
    def __init__(self, output_names, unc_weights=False, **std_kwargs):
        self.output_names = output_names
        self.with_unc_weights = bool(unc_weights)   # also provide `unc_weights` ~ U(0.5, 1.5) (utils.hist apply_unc_weights)
        super().__init__(expected_params=("n_events", "seed"), expected_container_keys=(),
                         supported_reps={"calc_mode": ["events"], "apply_mode": ["events"]}, **std_kwargs)

    def setup_function(self):
        n_events = int(self.params.n_events.value.m)
        seed = int(self.params.seed.value.m)
        dev = default_device()
        for i, name in enumerate(self.output_names):
            container = Container(name, representation="events")
            nubar = -1 if "bar" in name else 1
            flav = 2 if "tau" in name else (1 if "mu" in name else 0)
            ev = syn.make_events_torch(n_events, seed + i, FTYPE, dev)
            keep = local_slice(n_events)                         # this rank's share when events are sharded over GPUs
            if keep != slice(0, n_events):
                ev = {k: v[keep].contiguous() for k, v in ev.items()}
                n_events_local = keep.stop - keep.start
            else:
                n_events_local = n_events
            for key in ("true_energy", "true_coszen", "reco_energy", "reco_coszen", "pid", "nu_flux"):
                container[key] = ev[key]
            # nominal fluxes as flux.honda_ip would provide them (nu and nubar tables differ by ~20 %)
            container["nu_flux_nominal"] = ev["nu_flux"]
            container["nubar_flux_nominal"] = (ev["nu_flux"] * 0.8).contiguous()
            container["weighted_aeff"] = ev["weights"]
            if self.with_unc_weights:
                import torch
                g = torch.Generator(device=dev)
                g.manual_seed(seed + 100_003 + i)
                container["unc_weights"] = (0.5 + torch.rand(n_events_local, generator=g, device=dev,
                                                             dtype=torch.float64)).to(ev["weights"].dtype)
            container["initial_weights"] = ev["weights"].new_ones(n_events_local)
            container["weights"] = ev["weights"].new_ones(n_events_local)
            container.set_aux_data("nubar", nubar)
            container.set_aux_data("flav", flav)
            self.data.add_container(container)

    def apply_function(self):
        for container in self.data:
            container["weights"] = container["initial_weights"].clone()
