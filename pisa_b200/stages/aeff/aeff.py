"""``aeff.aeff`` service: ``weights *= weighted_aeff * livetime * norms`` (pisa/stages/aeff/aeff.py:22-88).

A "next" row of the scope table: one multiply per event.  It runs as a single
in-place device operation through torch (plumbing) until it is folded into the fused kernel.
"""
from pisa_b200.core.stage import Stage

__all__ = ["aeff"]


class aeff(Stage):  # pylint: disable=invalid-name
    def __init__(self, **std_kwargs):
        super().__init__(expected_params=("livetime", "aeff_scale", "nutau_cc_norm", "nutau_norm", "nu_nc_norm"),
                         expected_container_keys=("weights", "weighted_aeff"), **std_kwargs)

    def apply_function(self):
        aeff_scale = self.params.aeff_scale.m_as("dimensionless")
        livetime_s = self.params.livetime.m_as("sec")
        nutau_cc_norm = self.params.nutau_cc_norm.m_as("dimensionless")
        nutau_norm = self.params.nutau_norm.m_as("dimensionless")
        nu_nc_norm = self.params.nu_nc_norm.m_as("dimensionless")
        for container in self.data:
            scale = aeff_scale * livetime_s
            if container.name in ["nutau_cc", "nutaubar_cc"]:
                scale *= nutau_cc_norm
            if "nutau" in container.name:
                scale *= nutau_norm
            if "nc" in container.name:
                scale *= nu_nc_norm
            w = container["weights"]
            w *= container["weighted_aeff"] * scale
            container.mark_changed("weights")
