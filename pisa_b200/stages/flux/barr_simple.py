"""``flux.barr_simple`` service: Barr-style flux systematics on the nominal fluxes.

Drop-in for pisa/stages/flux/barr_simple.py (reference :23-226): ``expected_params``
``nue_numu_ratio, nu_nubar_ratio, delta_index, Barr_uphor_ratio, Barr_nu_nubar_ratio`` (:52-58), container
keys ``true_energy, true_coszen, nu_flux_nominal, nubar_flux_nominal, nubar`` read (:59-65) and ``nu_flux``
written (``setup_function`` allocates it :73-75, ``compute_function`` fills it and marks it changed
:78-104).  Like the reference it implements no ``apply_function``.

Everything transcendental in the reference's ``apply_sys_kernel`` (the parameterisations of
pisa/utils/barr_parameterization.py) depends on the event only, not on the five systematic parameters: those
terms are evaluated once in ``setup_function`` (``pisab_flux_barr_terms``, FP64), and ``compute_function`` makes
one HBM-bound call per container (``pisab_flux_barr_apply``) whenever a flux parameter changes.
"""
from pisa_b200 import ops
from pisa_b200.core.stage import Stage

__all__ = ["barr_simple", "init_test"]


class barr_simple(Stage):  # pylint: disable=invalid-name
    def __init__(self, **std_kwargs):
        expected_params = ("nue_numu_ratio", "nu_nubar_ratio", "delta_index", "Barr_uphor_ratio",
                           "Barr_nu_nubar_ratio")
        expected_container_keys = ("true_energy", "true_coszen", "nu_flux_nominal", "nubar_flux_nominal", "nubar")
        super().__init__(expected_params=expected_params, expected_container_keys=expected_container_keys,
                         **std_kwargs)

    def setup_function(self):
        for container in self.data:
            e = container["true_energy"]
            container["nu_flux"] = e.new_empty((container.size, 2))
            container["flux_barr_terms"] = ops.flux_barr_terms(e, container["true_coszen"])

    def compute_function(self):
        p = self.params
        nue_numu_ratio = p.nue_numu_ratio.value.m_as("dimensionless")
        nu_nubar_ratio = p.nu_nubar_ratio.value.m_as("dimensionless")
        delta_index = p.delta_index.value.m_as("dimensionless")
        uphor = p.Barr_uphor_ratio.value.m_as("dimensionless")
        nubar_sys = p.Barr_nu_nubar_ratio.value.m_as("dimensionless")
        for container in self.data:
            ops.flux_barr_apply(container["flux_barr_terms"], container["nu_flux_nominal"],
                                container["nubar_flux_nominal"], int(container["nubar"]), nue_numu_ratio,
                                nu_nubar_ratio, delta_index, uphor, nubar_sys, out=container["nu_flux"])
            container.mark_changed("nu_flux")


def init_test(**param_kwargs):
    """Instantiation example (barr_simple.py:229-239)."""
    from pisa_b200.core.param import Param, ParamSet
    return barr_simple(params=ParamSet([
        Param(name="nue_numu_ratio", value=1.0, **param_kwargs),
        Param(name="nu_nubar_ratio", value=1.0, **param_kwargs),
        Param(name="delta_index", value=0.0, **param_kwargs),
        Param(name="Barr_uphor_ratio", value=0.0, **param_kwargs),
        Param(name="Barr_nu_nubar_ratio", value=0.0, **param_kwargs),
    ]))
