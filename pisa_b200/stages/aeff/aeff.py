"""``aeff.aeff`` service: ``weights *= weighted_aeff * livetime * norms`` (pisa/stages/aeff/aeff.py:22-88).

A "next" row of the scope table: one multiply per event, ``pisab_scale_weights`` (24 B/event,
HBM-bound); in the fit loop the per-event factor is folded into the weights once and the scalar part travels as the
per-container ``scale`` of the fused kernel (``pisa_b200.fused``).
"""
from pisa_b200 import ops
from pisa_b200.core.stage import Stage

__all__ = ["aeff"]


class aeff(Stage):  # pylint: disable=invalid-name
    def __init__(self, **std_kwargs):
        super().__init__(expected_params=("livetime", "aeff_scale", "nutau_cc_norm", "nutau_norm", "nu_nc_norm"),
                         expected_container_keys=("weights", "weighted_aeff"), **std_kwargs)

    def container_scale(self, name):
        """The scalar part of the weight factor of container `name` (aeff.py:68-88): livetime, overall scale
        and the flavour / interaction norms."""
        p = self.params
        scale = p.aeff_scale.m_as("dimensionless") * p.livetime.m_as("sec")
        if name in ["nutau_cc", "nutaubar_cc"]:
            scale *= p.nutau_cc_norm.m_as("dimensionless")
        if "nutau" in name:
            scale *= p.nutau_norm.m_as("dimensionless")
        if "nc" in name:
            scale *= p.nu_nc_norm.m_as("dimensionless")
        return scale

    def apply_function(self):
        for container in self.data:
            ops.scale_weights(container["weights"], container["weighted_aeff"], self.container_scale(container.name))
            container.mark_changed("weights")
