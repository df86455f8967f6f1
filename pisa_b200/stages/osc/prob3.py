"""``osc.prob3`` service: three-flavour oscillation probabilities through the Earth on a B200.

Drop-in for pisa/stages/osc/prob3.py (reference :37-641): same constructor kwargs
(``include_nlo, nsi_type, reparam_mix_matrix, neutrino_decay, tomography_type, lri_type``), same
``expected_params`` (:178-191 plus the 9 ``eps_*`` for ``nsi_type='standard'`` :244-254), same
container keys read (``true_energy, true_coszen, nubar, flav, nu_flux, weights``) and written
(``probability, prob_e, prob_mu``; ``weights`` multiplied in ``apply_function`` :621-622), same
linking of the flavour containers in grid mode (:398-404,415-419,456-459).

What runs where: ``compute_function`` builds the PMNS / dm / matter-potential matrices on the host
exactly like :485-567 and then makes ONE call per container into the CUDA library
(``pisab_prob3_propagate_earth``), which evaluates the Earth layers (reference: ``Layers.calcLayers``
at setup, :406-409) and the propagation (``propagate_array`` :439-450) in one kernel and writes
``probability``, ``prob_e`` and ``prob_mu`` (``fill_probs`` :593-605).  The ``densities`` /
``distances`` arrays of the reference are therefore not materialised by default
(``store_layers=True`` restores them for inspection).

Both NSI parameterisations (:232-254,341-346), the long-range-interaction potential (:272-275,519-520,567-575)
and the Earth-tomography density scalings (:278-297,368-395,521-537) are host-side parameter -> matrix / density
work and follow the reference's sequence of calls.  ``neutrino_decay=True`` (:224-230,256-259,349-351,516-517,561-563) adds
the ``decay_alpha3`` parameter and sends ``decay_flag = 1`` with ``diag(0, 0, -i alpha3)`` to the library, whose decay
kernels (csrc/prob3_decay.cuh) evaluate the non-Hermitian layers that the reference hands to ``numpy.linalg.eigvals``.
"""
import numpy as np

from pisa_b200 import ops
from pisa_b200.core.stage import Stage
from pisa_b200.stages.osc.decay_params import DecayParams
from pisa_b200.stages.osc.layers import Layers
from pisa_b200.stages.osc.lri_params import LRI_TYPES, LRIParams
from pisa_b200.stages.osc.nsi_params import StdNSIParams, VacuumLikeNSIParams
from pisa_b200.stages.osc.scaling_params import (FIVE_LAYER_RADII, FIVE_LAYER_RHOS, TOMOGRAPHY_ERROR_MSG,
                                                 TOMOGRAPHY_TYPES, Core_scaling_w_constrain,
                                                 Core_scaling_wo_constrain, Mass_scaling)
from pisa_b200.stages.osc.osc_params import OscParams
from pisa_b200.utils.resources import find_resource

__all__ = ["prob3", "init_test", "NSI_TYPES", "LRI_TYPES", "TOMOGRAPHY_TYPES"]

NSI_TYPES = ["standard", "vacuum-like"]

_NU = ["nue_cc", "numu_cc", "nutau_cc", "nue_nc", "numu_nc", "nutau_nc"]
_NUBAR = ["nuebar_cc", "numubar_cc", "nutaubar_cc", "nuebar_nc", "numubar_nc", "nutaubar_nc"]


class prob3(Stage):  # pylint: disable=invalid-name
    def __init__(self, include_nlo=False, nsi_type=None, reparam_mix_matrix=False, neutrino_decay=False,
                 tomography_type=None, lri_type=None, store_layers=False, **std_kwargs):
        expected_params = ("detector_depth", "earth_model", "prop_height", "YeI", "YeO", "YeM", "theta12",
                           "theta13", "theta23", "deltam21", "deltam31", "deltacp")
        expected_container_keys = ("true_energy", "true_coszen", "nubar", "flav", "nu_flux", "weights")
        self.include_nlo = include_nlo
        if nsi_type is not None:
            nsi_type = nsi_type.strip().lower()
            if nsi_type not in NSI_TYPES:
                raise ValueError('Chosen NSI type "%s" not available! Choose one of %s.' % (nsi_type, NSI_TYPES))
        self.nsi_type = nsi_type
        self.reparam_mix_matrix = reparam_mix_matrix
        self.neutrino_decay = bool(neutrino_decay)
        self.decay_flag = 1 if neutrino_decay else -1   # :227-230
        decay_params = ("decay_alpha3",) if neutrino_decay else ()
        lri_params = ()
        if lri_type is not None:
            lri_type = lri_type.strip().lower()
            if lri_type not in LRI_TYPES:
                raise ValueError('Chosen LRI symmetry type "%s" not available! Choose one of %s.' % (lri_type, LRI_TYPES))
            lri_params = ("v_lri",)
        self.lri_type = lri_type
        tomography_params = ()
        if tomography_type is not None:
            tomography_type = tomography_type.strip().lower()
            if tomography_type not in TOMOGRAPHY_TYPES:
                raise ValueError('Chosen tomography type "%s" not available! Choose one of %s.'
                                 % (tomography_type, TOMOGRAPHY_TYPES))
            tomography_params = {"mass_of_earth": ("density_scale",),
                                 "mass_of_core_w_constrain": ("core_density_scale",),
                                 "mass_of_core_wo_constrain": ("core_density_scale", "innermantle_density_scale",
                                                               "middlemantle_density_scale")}[tomography_type]
        self.tomography_type = tomography_type
        nsi_params = ()
        if nsi_type == "standard":
            nsi_params = ("eps_ee", "eps_emu_magn", "eps_emu_phase", "eps_etau_magn", "eps_etau_phase",
                          "eps_mumu", "eps_mutau_magn", "eps_mutau_phase", "eps_tautau")
        elif nsi_type == "vacuum-like":
            nsi_params = ("eps_scale", "eps_prime", "phi12", "phi13", "phi23", "alpha1", "alpha2", "deltansi")
        super().__init__(expected_params=expected_params + nsi_params + decay_params + lri_params + tomography_params,
                         expected_container_keys=expected_container_keys, **std_kwargs)
        self.store_layers = store_layers
        self.layers = None
        self.osc_params = None
        self.nsi_params = None
        self.lri_params = None
        self.decay_params = None
        self.tomography_params = None
        self.gen_mat_pot_matrix_complex = None
        self.decay_matrix = np.zeros((3, 3), dtype=np.complex128)
        self.lri_pot = np.zeros((3, 3), dtype=np.float64)
        self.YeI = self.YeO = self.YeM = None
        self._earth = None
        self._orders = {}

    # ------------------------------------------------------------------------------------
    def _link(self):
        if self.is_map:
            self.data.link_containers("nu", _NU)
            self.data.link_containers("nubar", _NUBAR)

    def _store_layers(self):
        if not self.store_layers:
            return
        for container in self.data:
            self.layers.calcLayers(container["true_coszen"])
            container["densities"] = self.layers.density
            container["distances"] = self.layers.distance

    def setup_function(self):
        self.osc_params = OscParams()
        if self.nsi_type == "standard":
            self.nsi_params = StdNSIParams()
        elif self.nsi_type == "vacuum-like":
            self.nsi_params = VacuumLikeNSIParams()
        if self.lri_type is not None:
            self.lri_params = LRIParams()
        if self.neutrino_decay:
            self.decay_params = DecayParams()
        earth_model = find_resource(self.params.earth_model.value)
        self.YeI = self.params.YeI.value.m_as("dimensionless")
        self.YeO = self.params.YeO.value.m_as("dimensionless")
        self.YeM = self.params.YeM.value.m_as("dimensionless")
        prop_height = self.params.prop_height.value.m_as("km")
        detector_depth = self.params.detector_depth.value.m_as("km")
        self.layers = Layers(earth_model, detector_depth, prop_height)
        self.layers.setElecFrac(self.YeI, self.YeO, self.YeM)
        self._earth = self.layers.earth_struct()
        if self.tomography_type == "mass_of_earth":
            self.tomography_params = Mass_scaling()
        elif self.tomography_type is not None:
            # the external Earth model must be the hard-coded 5-layer one (:378-389)
            radii_ext = self.layers.radii[::-1][:-1]
            rhos_ext = self.layers.rhos_unweighted[::-1][:-1]
            if len(radii_ext) != len(FIVE_LAYER_RADII) or len(rhos_ext) != len(FIVE_LAYER_RHOS):
                raise ValueError(TOMOGRAPHY_ERROR_MSG)
            if not (np.allclose(radii_ext + 1, FIVE_LAYER_RADII + 1) and np.allclose(rhos_ext + 1, FIVE_LAYER_RHOS + 1)):
                raise ValueError(TOMOGRAPHY_ERROR_MSG)
            self.tomography_params = (Core_scaling_w_constrain() if self.tomography_type == "mass_of_core_w_constrain"
                                      else Core_scaling_wo_constrain())

        # layers do not care about flavour: link everything while touching true_coszen (:398-412)
        if self.is_map:
            self.data.link_containers("nu", _NU + _NUBAR)
        self._store_layers()
        # thread order grouping events by crossed shells: depends on true_coszen only
        self._orders = {}
        for container in self.data:
            order = ops.layer_order(self._earth, container["true_coszen"])
            for c in (container.containers if hasattr(container, "containers") else [container]):
                self._orders[c.name] = order
        self.data.unlink_containers()

        # output arrays (:414-427)
        self._link()
        for container in self.data:
            n = container.size
            tc = container["true_coszen"]
            container["probability"] = tc.new_empty((n, 3, 3))
        self.data.unlink_containers()
        for container in self.data:
            tc = container["true_coszen"]
            container["prob_e"] = tc.new_empty(container.size)
            container["prob_mu"] = tc.new_empty(container.size)

    def _update_matrices(self):
        p = self.params
        for angle in (p.theta12, p.theta13, p.theta23, p.deltacp):
            if angle.value.units == "dimensionless":
                raise ValueError("%s is dimensionless, but needs units rad or deg!" % angle.name)
        o = self.osc_params
        o.theta12 = p.theta12.value.m_as("rad")
        o.theta13 = p.theta13.value.m_as("rad")
        o.theta23 = p.theta23.value.m_as("rad")
        o.dm21 = p.deltam21.value.m_as("eV**2")
        o.dm31 = p.deltam31.value.m_as("eV**2")
        o.deltacp = p.deltacp.value.m_as("rad")
        std = np.zeros((3, 3), dtype=np.complex128)
        std[0, 0] += 1.020 if self.include_nlo else 1.0   # :540-546
        if self.nsi_type == "standard":
            n = self.nsi_params
            n.eps_ee = p.eps_ee.value.m_as("dimensionless")
            n.eps_emu = (p.eps_emu_magn.value.m_as("dimensionless"), p.eps_emu_phase.value.m_as("rad"))
            n.eps_etau = (p.eps_etau_magn.value.m_as("dimensionless"), p.eps_etau_phase.value.m_as("rad"))
            n.eps_mumu = p.eps_mumu.value.m_as("dimensionless")
            n.eps_mutau = (p.eps_mutau_magn.value.m_as("dimensionless"), p.eps_mutau_phase.value.m_as("rad"))
            n.eps_tautau = p.eps_tautau.value.m_as("dimensionless")
            self.gen_mat_pot_matrix_complex = std + n.eps_matrix
        elif self.nsi_type == "vacuum-like":
            n = self.nsi_params
            n.eps_scale = p.eps_scale.value.m_as("dimensionless")
            n.eps_prime = p.eps_prime.value.m_as("dimensionless")
            for name in ("phi12", "phi13", "phi23", "alpha1", "alpha2", "deltansi"):
                setattr(n, name, p[name].value.m_as("rad"))
            self.gen_mat_pot_matrix_complex = std + n.eps_matrix
        else:
            self.gen_mat_pot_matrix_complex = std
        if self.lri_type is not None:                        # :519-520,567-575
            self.lri_params.v_lri = p.v_lri.value.m_as("eV")
            self.lri_pot = self.lri_params.potential_matrix(self.lri_type)
        if self.neutrino_decay:                              # :516-517,561-563
            self.decay_params.decay_alpha3 = p.decay_alpha3.value.m_as("eV**2")
            self.decay_matrix = self.decay_params.decay_matrix
        mix = o.mix_matrix_reparam_complex if self.reparam_mix_matrix else o.mix_matrix_complex
        return ops.OscConsts.from_matrices(o.dm_matrix, mix, self.gen_mat_pot_matrix_complex, self.decay_flag,
                                           self.decay_matrix, self.lri_pot)

    def _apply_tomography(self):
        """prob3.py:521-537, call for call: ``Layers.scaling`` with the type's factors, then ``setElecFrac`` and new
        layer densities.  (``setElecFrac`` weights ``rhos_unweighted`` -- layers.py:411-439 -- which ``scaling``
        does not touch, so in this reference revision the factors do not reach the propagated densities; the
        sequence is kept as it is so that results stay identical.)"""
        p, t = self.params, self.tomography_params
        if self.tomography_type == "mass_of_earth":
            t.density_scale = p.density_scale.value.m_as("dimensionless")
            self.layers.scaling(scaling_array=t.density_scale)
        elif self.tomography_type == "mass_of_core_w_constrain":
            t.core_density_scale = p.core_density_scale.value.m_as("dimensionless")
            self.layers.scaling(scaling_array=t.scaling_array)
        else:
            t.core_density_scale = p.core_density_scale.value.m_as("dimensionless")
            t.innermantle_density_scale = p.innermantle_density_scale.value.m_as("dimensionless")
            t.middlemantle_density_scale = p.middlemantle_density_scale.value.m_as("dimensionless")
            self.layers.scaling(scaling_array=t.scaling_factor_array)
        self.layers.setElecFrac(self.YeI, self.YeO, self.YeM)
        self._earth = self.layers.earth_struct()
        self._store_layers()

    def update_hypothesis(self):
        """Host part of ``compute_function`` (:463-567): Earth densities for the current electron fractions /
        tomography scalings and the oscillation matrices.  Returns (OscConsts, Earth struct); also used by
        ``pisa_b200.fused.FusedPipeline``."""
        YeI = self.params.YeI.value.m_as("dimensionless")
        YeO = self.params.YeO.value.m_as("dimensionless")
        YeM = self.params.YeM.value.m_as("dimensionless")
        if YeI != self.YeI or YeO != self.YeO or YeM != self.YeM:   # :466-474
            self.YeI, self.YeO, self.YeM = YeI, YeO, YeM
            self.layers.setElecFrac(YeI, YeO, YeM)
            self._earth = self.layers.earth_struct()
            self._store_layers()
        if self.tomography_type is not None:
            self._apply_tomography()
        return self._update_matrices(), self._earth

    def compute_function(self):
        consts, _ = self.update_hypothesis()

        self._link()
        for container in self.data:
            first = container.containers[0] if hasattr(container, "containers") else container
            ops.propagate_earth(consts, self._earth, int(container["nubar"]), container["true_energy"],
                                container["true_coszen"], probability=container["probability"],
                                order=self._orders.get(first.name))
            container.mark_changed("probability")
        self.data.unlink_containers()   # the following is flavour specific (:590-608)
        for container in self.data:
            flav = int(container["flav"])
            ops.fill_probs(container["probability"], 0, flav, out=container["prob_e"])
            ops.fill_probs(container["probability"], 1, flav, out=container["prob_mu"])
            container.mark_changed("prob_e")
            container.mark_changed("prob_mu")

    def apply_function(self):
        for container in self.data:
            w = container["weights"]
            ops.apply_osc_weights(container["nu_flux"], container["prob_e"], container["prob_mu"], w)
            container.mark_changed("weights")


def init_test(**param_kwargs):
    """Initialisation example (prob3.py:625-641)."""
    from pisa_b200.core.param import Param, ParamSet
    from pisa_b200.utils.units import ureg
    return prob3(include_nlo=True, params=ParamSet([
        Param(name="detector_depth", value=10 * ureg.km, **param_kwargs),
        Param(name="prop_height", value=18 * ureg.km, **param_kwargs),
        Param(name="earth_model", value="osc/PREM_4layer.dat", **param_kwargs),
        Param(name="YeI", value=0.5, **param_kwargs),
        Param(name="YeO", value=0.5, **param_kwargs),
        Param(name="YeM", value=0.5, **param_kwargs),
        Param(name="theta12", value=33 * ureg.degree, **param_kwargs),
        Param(name="theta13", value=8 * ureg.degree, **param_kwargs),
        Param(name="theta23", value=50 * ureg.degree, **param_kwargs),
        Param(name="deltam21", value=8e-5 * ureg.eV ** 2, **param_kwargs),
        Param(name="deltam31", value=3e-3 * ureg.eV ** 2, **param_kwargs),
        Param(name="deltacp", value=180 * ureg.degree, **param_kwargs),
    ]))
