"""Host-side parameter builders vs matrices produced by the reference classes (CPU only)."""
import os

import numpy as np
import pytest

from conftest import ROOT, load_golden
from pisa_b200.stages.osc.layers import Layers
from pisa_b200.stages.osc.nsi_params import StdNSIParams
from pisa_b200.stages.osc.osc_params import OscParams

import sys
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import PARAM_SETS  # noqa: E402  (pure-python dict; no reference import at module level)


def _osc(ps):
    op = OscParams()
    op.theta12, op.theta13, op.theta23 = (np.deg2rad(ps[k]) for k in ("t12", "t13", "t23"))
    op.deltacp = np.deg2rad(ps["dcp"])
    op.dm21, op.dm31 = ps["dm21"], ps["dm31"]
    return op


def test_mix_and_dm_matrices_bitwise():
    g = load_golden("ref_params_f8.npz")
    for name, ps in PARAM_SETS.items():
        op = _osc(ps)
        assert np.array_equal(op.mix_matrix_complex, g[name + "/mix"]), name
        assert np.array_equal(op.mix_matrix_reparam_complex, g[name + "/mix_reparam"]), name
        assert np.array_equal(op.dm_matrix, g[name + "/dm"]), name
    op = OscParams()
    assert np.array_equal(op.dm_matrix, g["degenerate/dm"])


def test_std_nsi_eps_matrix_bitwise():
    g = load_golden("ref_params_f8.npz")
    ps = PARAM_SETS["nufit20_nh_dcp306_stdnsi"]["nsi"]
    nsi = StdNSIParams()
    nsi.eps_ee = ps["eps_ee"]
    nsi.eps_emu = (ps["eps_emu"][0], np.deg2rad(ps["eps_emu"][1]))
    nsi.eps_etau = (ps["eps_etau"][0], np.deg2rad(ps["eps_etau"][1]))
    nsi.eps_mumu = ps["eps_mumu"]
    nsi.eps_mutau = (ps["eps_mutau"][0], np.deg2rad(ps["eps_mutau"][1]))
    nsi.eps_tautau = ps["eps_tautau"]
    assert np.array_equal(nsi.eps_matrix, g["nufit20_nh_dcp306_stdnsi/eps"])
    std = np.zeros((3, 3), dtype=np.complex128)
    std[0, 0] += 1.0
    assert np.array_equal(std + nsi.eps_matrix, g["nufit20_nh_dcp306_stdnsi/mat_pot"])


def test_layers_host_tables_bitwise():
    from pisa_b200 import FTYPE
    g = load_golden("ref_layers_f8.npz" if FTYPE == np.float64 else "ref_layers_f4.npz")   # PISA_FTYPE-dependent tables
    for key in sorted({k.rsplit("/", 1)[0] for k in g.files}):
        model = key.split("/")[0]
        depth, height, yei, yeo, yem = g[key + "/params"]
        L = Layers(os.path.join(ROOT, "pisa_b200", "resources", "osc", model + ".dat"), depth, height)
        L.setElecFrac(yei, yeo, yem)
        assert np.array_equal(L.radii, g[key + "/radii"])
        assert np.array_equal(L.rhos, g[key + "/rhos"])
        assert np.array_equal(L.coszen_limit, g[key + "/coszen_limit"])
        assert L.max_layers == int(g[key + "/max_layers"]) and L.r_detector == float(g[key + "/r_detector"])
        # idempotence of setElecFrac (test_layers_4, layers.py:669-772)
        rhos = L.rhos.copy()
        L.setElecFrac(yei, yeo, yem)
        assert np.array_equal(rhos, L.rhos)


def test_vectorised_osc_consts_equal_the_scalar_builder():
    """scan.osc_consts_array (one numpy pass for P hypotheses) == scan.osc_consts per hypothesis, bit for bit."""
    import ctypes
    from pisa_b200 import scan
    rng = np.random.default_rng(3)
    n = 40
    t23, dm31 = rng.uniform(0.5, 1.0, n), rng.uniform(-3e-3, 3e-3, n)
    dm31[3] = 0.0    # degeneracy nudge branch of dm_matrix
    fixed = dict(theta12=0.58, theta13=0.148, deltacp=4.1, dm21=7.5e-5)
    mp = np.array([[1.0, 0.07 + 0.02j, 0], [0.07 - 0.02j, 0, 0.003j], [0, -0.003j, 0.1]])
    for mat_pot in (None, mp):
        arr = scan.osc_consts_array(fixed["theta12"], fixed["theta13"], t23, fixed["deltacp"], fixed["dm21"], dm31, mat_pot)
        assert len(arr) == n
        for k in range(n):
            one = scan.osc_consts(fixed["theta12"], fixed["theta13"], t23[k], fixed["deltacp"], fixed["dm21"], dm31[k], mat_pot)
            assert bytes(one) == bytes(arr[k]), k
    with pytest.raises(AssertionError):
        scan.osc_consts_array(0.5, 0.1, 0.7, 7.0, 7e-5, 2e-3)
    # the optional seventh parameter (neutrino decay): decay_flag = 1 and mat_decay = diag(0, 0, -i alpha3), the record
    # DecayParams + OscConsts.from_matrices give; per-hypothesis values and a broadcast scalar
    from pisa_b200 import ops
    from pisa_b200.stages.osc.decay_params import DecayParams
    alpha = rng.uniform(0, 1e-3, n)
    arr = scan.osc_consts_array(fixed["theta12"], fixed["theta13"], t23, fixed["deltacp"], fixed["dm21"], dm31, mp,
                                decay_alpha3=alpha)
    arr_s = scan.osc_consts_array(fixed["theta12"], fixed["theta13"], t23, fixed["deltacp"], fixed["dm21"], dm31, mp,
                                  decay_alpha3=2.0e-4)
    for k in range(n):
        one = scan.osc_consts(fixed["theta12"], fixed["theta13"], t23[k], fixed["deltacp"], fixed["dm21"], dm31[k], mp,
                              decay_alpha3=alpha[k])
        assert bytes(one) == bytes(arr[k]) and int(arr[k].decay_flag) == 1 and int(arr_s[k].decay_flag) == 1
        d = DecayParams()
        d.decay_alpha3 = alpha[k]
        op = OscParams()
        op.theta12, op.theta13, op.theta23, op.deltacp, op.dm21, op.dm31 = 0.58, 0.148, t23[k], 4.1, 7.5e-5, dm31[k]
        ref = ops.OscConsts.from_matrices(op.dm_matrix, op.mix_matrix_complex, mp, 1, d.decay_matrix)
        assert bytes(ref) == bytes(one), k
        assert list(arr_s[k].mat_decay)[16:18] == [0.0, -2.0e-4]


def test_scalar_osc_consts_equal_oscparams_matrices():
    """scan.osc_consts (scalar fast path of the fit loop) == OscConsts.from_matrices(OscParams ...), bit for bit."""
    from pisa_b200 import ops, scan
    rng = np.random.default_rng(11)
    for _ in range(200):
        t12, t13, t23 = rng.uniform(0.1, 1.4, 3)
        dcp = rng.uniform(0, 2 * np.pi)
        m21, m31 = rng.uniform(1e-5, 1e-4), rng.choice([-1, 1]) * rng.uniform(1e-3, 4e-3)
        op = OscParams()
        op.theta12, op.theta13, op.theta23, op.deltacp, op.dm21, op.dm31 = t12, t13, t23, dcp, m21, m31
        mp = np.zeros((3, 3), dtype=np.complex128)
        mp[0, 0] = 1.0
        ref = ops.OscConsts.from_matrices(op.dm_matrix, op.mix_matrix_complex, mp)
        got = scan.osc_consts(t12, t13, t23, dcp, m21, m31)
        assert bytes(ref) == bytes(got)


def test_vacuum_like_nsi_matrix_equals_reference():
    """VacuumLikeNSIParams.eps_matrix (nsi_params.py:326-384) on seeded parameter sets: bit-identical to the
    unmodified reference (tests/golden/ref_params_f8.npz, generated by make_golden.gen_params)."""
    from pisa_b200.stages.osc.nsi_params import VacuumLikeNSIParams
    g = load_golden("ref_params_f8.npz")
    names = [str(x) for x in g["vacuum_nsi/names"]]
    for vals, ref in zip(g["vacuum_nsi/values"], g["vacuum_nsi/eps"]):
        v = VacuumLikeNSIParams()
        for k, x in zip(names, vals):
            setattr(v, k, float(x))
        assert np.array_equal(v.eps_matrix, ref)
        assert v.eps_emu == ref[0, 1] and v.eps_tautau == ref[2, 2].real and v.eps_mumu == 0.0
    v = VacuumLikeNSIParams()
    assert np.array_equal(v.eps_matrix, np.zeros((3, 3)))           # defaults: the standard potential only
    with pytest.raises(AssertionError):
        v.phi12 = 4.0
    with pytest.raises(AssertionError):
        v.alpha1 = -0.1
    with pytest.raises(TypeError):
        v.eps_scale = 1.0 + 1.0j


def test_lri_and_tomography_parameter_classes():
    """LRIParams (lri_params.py:24-116) and the tomography scalings (scaling_params.py): the constrained core scaling
    keeps the 5-layer Earth's mass and moment of inertia (its defining equations; parity unpinned -- the reference
    module needs pint)."""
    from pisa_b200.stages.osc.lri_params import LRIParams
    from pisa_b200.stages.osc.scaling_params import (FIVE_LAYER_RADII as R, FIVE_LAYER_RHOS as RHO,
                                                     Core_scaling_w_constrain, Core_scaling_wo_constrain, Mass_scaling)
    lri = LRIParams()
    lri.v_lri = 3e-14
    assert np.array_equal(lri.potential_matrix_emu, np.diag([3e-14, -3e-14, 0.0]))
    assert np.array_equal(lri.potential_matrix_etau, np.diag([3e-14, 0.0, -3e-14]))
    assert np.array_equal(lri.potential_matrix_mutau, np.diag([0.0, 3e-14, -3e-14]))
    with pytest.raises(AssertionError):
        lri.v_lri = 2.0
    with pytest.raises(ValueError):
        lri.potential_matrix("ee-symmetry")
    m = Mass_scaling()
    m.density_scale = 1.2
    with pytest.raises(AssertionError):
        m.density_scale = -1.0
    c = Core_scaling_w_constrain()
    c.core_density_scale = 1.0
    assert np.allclose(c.scaling_array, np.ones(6), rtol=0, atol=1e-12)
    c.core_density_scale = 1.03
    s = c.scaling_array                     # [outer mantle, middle mantle, inner mantle, core, core, core]
    assert s[0] == 1.0 and s[3] == s[4] == s[5] == 1.03
    fac = np.array([s[3], s[3], s[2], s[1], 1.0])          # shells from the centre outwards
    for power in (3, 5):
        shell = RHO[1:] * (R[1:] ** power - R[:-1] ** power)
        assert abs((fac * shell).sum() / shell.sum() - 1.0) < 1e-14
    w = Core_scaling_wo_constrain()
    w.core_density_scale, w.innermantle_density_scale, w.middlemantle_density_scale = 1.1, 0.9, 1.05
    assert w.scaling_factor_array.tolist() == [1.0, 1.05, 0.9, 1.1, 1.1, 1.1]


def test_fit_driver_argument_checks_need_no_device():
    """scan.fit_chi2 validates its parameter names before it touches the engine: the six oscillation parameters must
    each be free or fixed exactly once, decay_alpha3 is the only optional seventh name."""
    from pisa_b200 import scan
    fixed = dict(theta12=0.58, theta13=0.148, deltacp=4.1, dm21=7.5e-5)
    with pytest.raises(ValueError, match="unknown oscillation parameter"):
        scan.fit_chi2(None, None, dict(theta24=0.1), fixed)
    with pytest.raises(ValueError, match="free or fixed"):
        scan.fit_chi2(None, None, dict(theta23=0.7), fixed)                      # dm31 missing
    with pytest.raises(ValueError, match="free or fixed"):
        scan.fit_chi2(None, None, dict(theta23=0.7, dm31=2.5e-3, decay_alpha3=1e-4), dict(fixed, decay_alpha3=2e-4))
    with pytest.raises(ValueError, match="unknown oscillation parameter"):
        scan.fit_chi2(None, None, dict(theta23=0.7, dm31=2.5e-3), dict(fixed, decay_alpha4=1e-4))
    assert scan.DECAY_PARAM_NAME == "decay_alpha3" and scan.DECAY_PARAM_NAME not in scan.OSC_PARAM_NAMES
