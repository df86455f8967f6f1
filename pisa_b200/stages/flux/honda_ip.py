"""``flux.honda_ip`` service: nominal atmospheric fluxes from an azimuth-averaged Honda table.

Drop-in for pisa/stages/flux/honda_ip.py (reference :24-125): ``expected_params = ('flux_table',)`` (:45-47),
container keys ``true_energy, true_coszen`` read (:48-51) and ``nu_flux_nominal`` = (nue, numu),
``nubar_flux_nominal`` = (nuebar, numubar) written; in map mode all twelve containers are linked while the
fluxes are evaluated, because the nominal flux does not depend on the outgoing flavour (:63-69,82-88).

The table is read and splined on the host exactly like ``load_2d_table`` (``pisa_b200.utils.flux_weights``);
``compute_function`` then makes ONE call per container into the CUDA library (``pisab_flux_honda_2d``) for all
four primaries -- the reference evaluates ``calculate_2d_flux_weights`` four times per container with a
per-event spline fit in Python (23 s of setup in its published profile).
"""
from pisa_b200 import ops
from pisa_b200.core.stage import Stage
from pisa_b200.utils.flux_weights import HondaTable2D

__all__ = ["honda_ip", "init_test"]

_ALL = ["nue_cc", "numu_cc", "nutau_cc", "nue_nc", "numu_nc", "nutau_nc",
        "nuebar_cc", "numubar_cc", "nutaubar_cc", "nuebar_nc", "numubar_nc", "nutaubar_nc"]


class honda_ip(Stage):  # pylint: disable=invalid-name
    def __init__(self, **std_kwargs):
        super().__init__(expected_params=("flux_table",), expected_container_keys=("true_energy", "true_coszen"),
                         **std_kwargs)
        self.flux_table = None

    def _link(self):
        if self.data.is_map:
            self.data.link_containers("nu", [n for n in _ALL if n in self.data.names])

    def setup_function(self):
        self.flux_table = HondaTable2D(self.params.flux_table.value)
        self._link()
        for container in self.data:
            e = container["true_energy"]
            container["nu_flux_nominal"] = e.new_empty((container.size, 2))
            container["nubar_flux_nominal"] = e.new_empty((container.size, 2))
        self.data.unlink_containers()

    def compute_function(self):
        self._link()
        for container in self.data:
            ops.flux_honda_2d(self.flux_table, container["true_energy"], container["true_coszen"],
                              container["nu_flux_nominal"], container["nubar_flux_nominal"])
            container.mark_changed("nu_flux_nominal")
            container.mark_changed("nubar_flux_nominal")
        self.data.unlink_containers()


def init_test(**param_kwargs):
    """Instantiation example (honda_ip.py:118-125)."""
    from pisa_b200.core.param import Param, ParamSet
    return honda_ip(params=ParamSet([Param(name="flux_table", value="flux/honda-2015-spl-solmin-aa.d", **param_kwargs)]))
