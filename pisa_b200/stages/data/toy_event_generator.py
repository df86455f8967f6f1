"""``data.toy_event_generator`` service (pisa/stages/data/toy_event_generator.py, reference :17-104).

Creates one container per ``output_names`` entry with ``nubar`` / ``flav`` aux scalars, unit
``weights`` / ``weighted_aeff``, ``initial_weights`` (random or ones) and a nominal flux of
(nue, numu) = (0, 1); in an array representation it also draws ``true_energy = 10**(3 U)`` and
``true_coszen = 2 U - 1`` from ``numpy.random.RandomState(seed)`` exactly like :75-76, then uploads
the arrays to the device.  ``apply_function`` resets ``weights`` to ``initial_weights`` (:98-102).
"""
import numpy as np

from pisa_b200 import FTYPE
from pisa_b200.core.binning import MultiDimBinning
from pisa_b200.core.container import Container
from pisa_b200.core.stage import Stage

__all__ = ["toy_event_generator", "init_test"]


class toy_event_generator(Stage):  # pylint: disable=invalid-name
    def __init__(self, output_names, **std_kwargs):
        self.output_names = output_names
        super().__init__(expected_params=("n_events", "random", "seed"), expected_container_keys=(), **std_kwargs)

    def setup_function(self):
        n_events = int(self.params.n_events.value.m)
        seed = int(self.params.seed.value.m)
        self.random_state = np.random.RandomState(seed)
        for name in self.output_names:
            container = Container(name, representation=self.calc_mode)
            nubar = -1 if "bar" in name else 1
            if "e" in name:
                flav = 0
            if "mu" in name:
                flav = 1
            if "tau" in name:
                flav = 2
            if not isinstance(self.calc_mode, MultiDimBinning):
                container["true_energy"] = np.power(10, self.random_state.rand(n_events).astype(FTYPE) * 3)
                container["true_coszen"] = self.random_state.rand(n_events).astype(FTYPE) * 2 - 1
            size = container.size
            if self.params.random.value:
                container["initial_weights"] = self.random_state.rand(size).astype(FTYPE)
            else:
                container["initial_weights"] = np.ones(size, dtype=FTYPE)
            container.set_aux_data("nubar", nubar)
            container.set_aux_data("flav", flav)
            container["weights"] = np.ones(size, dtype=FTYPE)
            container["weighted_aeff"] = np.ones(size, dtype=FTYPE)
            flux = np.stack([np.zeros(size, dtype=FTYPE), np.ones(size, dtype=FTYPE)], axis=1)
            container["nu_flux_nominal"] = flux
            container["nubar_flux_nominal"] = flux
            self.data.add_container(container)

    def apply_function(self):
        for container in self.data:
            container["weights"] = container["initial_weights"].clone()


def init_test(**param_kwargs):
    from pisa_b200.core.param import Param, ParamSet
    return toy_event_generator(output_names=["numu", "nue_bar"], params=ParamSet([
        Param(name="n_events", value=100, **param_kwargs),
        Param(name="random", value=1, **param_kwargs),
        Param(name="seed", value=666, **param_kwargs)]))
