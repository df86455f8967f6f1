"""``flux.nominal`` service: ``nu_flux`` = the nominal flux of the container's species.

Placeholder for the reference's flux stages (``flux.honda_ip`` + ``flux.barr_simple``), which are
"next" rows of the scope table: it selects ``nu_flux_nominal`` or ``nubar_flux_nominal`` by the
container's ``nubar`` -- what ``flux.barr_simple`` computes with all systematics at their nominal
values (pisa/stages/flux/barr_simple.py:136-204 with ratios = 1 and deltas = 0)."""
from pisa_b200.core.stage import Stage

__all__ = ["nominal"]


class nominal(Stage):  # pylint: disable=invalid-name
    def __init__(self, **std_kwargs):
        super().__init__(expected_params=(), expected_container_keys=("nu_flux_nominal", "nubar_flux_nominal", "nubar"),
                         **std_kwargs)

    def setup_function(self):
        for container in self.data:
            key = "nu_flux_nominal" if container["nubar"] > 0 else "nubar_flux_nominal"
            container["nu_flux"] = container[key].clone()
