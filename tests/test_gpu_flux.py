"""GPU parity tests of flux.barr_simple: CUDA kernel vs the reference fixtures and the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from conftest import load_golden  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def test_flux_barr_simple_vs_reference_fixture():
    """every fixture case (5 systematic settings x nu/nubar) produced by the unmodified reference."""
    from pisa_b200 import ops
    dev = _dev()
    g = load_golden("ref_flux_f8.npz")
    T = lambda k: torch.tensor(g[k], device=dev)  # noqa: E731
    e, cz, nu, nb = T("true_energy"), T("true_coszen"), T("nu_flux_nominal"), T("nubar_flux_nominal")
    for name in sorted({k.split("/")[0] for k in g.files if "/" in k}):
        pars = [float(x) for x in g[name + "/params"]]
        for nubar, tag in ((1, "nu"), (-1, "nubar")):
            out = ops.flux_barr_simple(e, cz, nu, nb, nubar, *pars).cpu().numpy()
            ref = g["%s/%s" % (name, tag)]
            assert np.allclose(out, ref, rtol=1e-10, atol=1e-300), (name, tag, np.abs(out - ref).max())
            assert np.array_equal(out == 0, ref == 0)      # the zero-flux branch of apply_ratio_scale
    # FP32 storage mode against the reference's f4 fixture
    g4 = load_golden("ref_flux_f4.npz")
    T4 = lambda k: torch.tensor(g4[k], device=dev)  # noqa: E731
    out = ops.flux_barr_simple(T4("true_energy"), T4("true_coszen"), T4("nu_flux_nominal"), T4("nubar_flux_nominal"),
                               -1, *[float(x) for x in g4["all_down/params"]]).cpu().numpy()
    assert out.dtype == np.float32 and np.allclose(out, g4["all_down/nubar"], rtol=2e-5, atol=1e-6)


def test_flux_barr_simple_large_random_vs_oracle():
    from pisa_b200 import ops
    dev = _dev()
    rng = np.random.default_rng(9)
    n = 300_000
    e = 10 ** rng.uniform(0, 4, n)
    cz = rng.uniform(-1, 1, n)
    nu, nb = rng.uniform(0.0, 2.0, (n, 2)), rng.uniform(0.0, 2.0, (n, 2))
    for pars in ((1.2, 0.8, 0.2, -2.5, 2.0), (0.7, 1.3, -0.3, 1.7, -4.0)):
        for nubar in (1, -1):
            ref = oracle.flux_barr_simple(e, cz, nu, nb, nubar, *pars)
            out = ops.flux_barr_simple(*(torch.tensor(a, device=dev) for a in (e, cz, nu, nb)), nubar, *pars)
            assert np.allclose(out.cpu().numpy(), ref, rtol=1e-10, atol=1e-300)
    with pytest.raises(ValueError):
        ops.flux_barr_simple(torch.tensor(e, device=dev), torch.tensor(cz[:-1], device=dev),
                             torch.tensor(nu, device=dev), torch.tensor(nb, device=dev), 1, 1, 1, 0, 0, 0)


def test_flux_stage_in_a_pipeline():
    """flux.barr_simple selected from a cfg: nu_flux follows the systematic parameters and feeds prob3."""
    from pisa_b200.core.pipeline import Pipeline
    from pisa_b200.utils.units import ureg
    _dev()
    pipe = Pipeline("settings/pipeline/b200_flux_events.cfg")
    assert [s.service_name for s in pipe.stages] == ["synthetic_mc", "barr_simple", "prob3", "aeff", "hist"]
    out0 = pipe.get_outputs()
    c = pipe.data["numu_cc"]
    c.representation = "events"
    got = c["nu_flux"].cpu().numpy()
    ref = oracle.flux_barr_simple(c["true_energy"].cpu().numpy(), c["true_coszen"].cpu().numpy(),
                                  c["nu_flux_nominal"].cpu().numpy(), c["nubar_flux_nominal"].cpu().numpy(), 1,
                                  1.0, 1.0, 0.0, 0.0, 0.0)
    assert np.allclose(got, ref, rtol=1e-10)
    pipe.params.delta_index = 0.1 * ureg.dimensionless
    out1 = pipe.get_outputs()
    assert not np.allclose(out1["numu_cc"].hist, out0["numu_cc"].hist, rtol=1e-4)
    c.representation = "events"
    ref1 = oracle.flux_barr_simple(c["true_energy"].cpu().numpy(), c["true_coszen"].cpu().numpy(),
                                   c["nu_flux_nominal"].cpu().numpy(), c["nubar_flux_nominal"].cpu().numpy(), 1,
                                   1.0, 1.0, 0.1, 0.0, 0.0)
    assert np.allclose(c["nu_flux"].cpu().numpy(), ref1, rtol=1e-10)


def test_honda_flux_kernel_vs_reference_fixture():
    """flux.honda_ip arithmetic: the one-pass CUDA evaluation against the unmodified reference
    (calculate_2d_flux_weights on 600 seeded events incl. extrapolated energies and coszen = +-1)."""
    from pisa_b200 import ops
    from pisa_b200.utils.flux_weights import HondaTable2D
    dev = _dev()
    g = load_golden("ref_honda_f8.npz")
    T = HondaTable2D(str(g["table"]))
    e, cz = torch.tensor(g["true_energy"], device=dev), torch.tensor(g["true_coszen"], device=dev)
    nu, nubar = ops.flux_honda_2d(T, e, cz)
    got = {"nue": nu[:, 0], "numu": nu[:, 1], "nuebar": nubar[:, 0], "numubar": nubar[:, 1]}
    for prim, t in got.items():
        out, ref = t.cpu().numpy(), g[prim]
        assert np.allclose(out, ref, rtol=1e-10, atol=0), (prim, np.abs(out / ref - 1).max())
    # FP32 storage mode
    nu4, _ = ops.flux_honda_2d(T, e.float(), cz.float())
    assert nu4.dtype == torch.float32 and np.allclose(nu4[:, 1].cpu().numpy(), g["numu"], rtol=3e-5)
    with pytest.raises(ValueError):
        ops.flux_honda_2d(T, e, cz * 1.5)


def test_honda_flux_larger_sample_vs_oracle():
    from oracle import honda
    from pisa_b200 import ops
    from pisa_b200.utils.flux_weights import HondaTable2D
    dev = _dev()
    T = HondaTable2D("flux/honda-2015-spl-solmin-aa.d")
    rng = np.random.default_rng(21)
    n = 1500
    e = 10 ** rng.uniform(0, 3, n)
    cz = rng.uniform(-1, 1, n)
    nu, nubar = ops.flux_honda_2d(T, torch.tensor(e, device=dev), torch.tensor(cz, device=dev))
    for prim, t in (("numu", nu[:, 1]), ("nuebar", nubar[:, 0])):
        ref = honda.honda_2d_flux(e, cz, T.spline_dict[prim])
        assert np.allclose(t.cpu().numpy(), ref, rtol=1e-10, atol=0)
    # size-independent properties on 2e6 events: positive, finite, nu_mu > nu_e above 10 GeV, smooth in E
    n2 = 2_000_000
    e2 = torch.tensor(10 ** rng.uniform(0, 3, n2), device=dev)
    c2 = torch.tensor(rng.uniform(-1, 1, n2), device=dev)
    nu2, nb2 = ops.flux_honda_2d(T, e2, c2)
    assert torch.isfinite(nu2).all() and torch.isfinite(nb2).all() and float(nu2.min()) > 0 and float(nb2.min()) > 0
    hi = e2 > 10
    assert bool((nu2[hi, 1] > nu2[hi, 0]).all())


def test_full_flux_chain_pipeline():
    """flux.honda_ip -> flux.barr_simple -> osc.prob3 -> aeff.aeff -> utils.hist selected from one cfg (the stage
    order of the reference's IceCube_3y_neutrinos.cfg) on synthetic events."""
    from oracle import honda
    from pisa_b200.core.pipeline import Pipeline
    from pisa_b200.utils.flux_weights import HondaTable2D
    _dev()
    pipe = Pipeline("settings/pipeline/b200_icecube3y_events.cfg")
    assert [s.service_name for s in pipe.stages] == ["synthetic_mc", "honda_ip", "barr_simple", "prob3", "aeff", "hist"]
    out = pipe.get_outputs()
    assert out["numu_cc"].hist.shape == (8, 8, 2) and np.isfinite(out["numu_cc"].hist).all()
    c = pipe.data["nuebar_nc"]
    c.representation = "events"
    T = HondaTable2D("flux/honda-2015-spl-solmin-aa.d")
    sel = slice(0, 200)
    e, cz = c["true_energy"].cpu().numpy()[sel], c["true_coszen"].cpu().numpy()[sel]
    assert np.allclose(c["nubar_flux_nominal"].cpu().numpy()[sel, 1], honda.honda_2d_flux(e, cz, T.spline_dict["numubar"]),
                       rtol=1e-10)
    ref = oracle.flux_barr_simple(c["true_energy"].cpu().numpy(), c["true_coszen"].cpu().numpy(),
                                  c["nu_flux_nominal"].cpu().numpy(), c["nubar_flux_nominal"].cpu().numpy(), -1,
                                  1.0, 1.0, 0.0, 0.0, 0.0)
    assert np.allclose(c["nu_flux"].cpu().numpy(), ref, rtol=1e-10)


def test_flux_terms_plus_apply_equal_the_direct_kernel_and_the_oracle():
    """fit-loop form (terms once, cheap apply per template) vs the direct kernel and vs the oracle, all five
    fixture parameter sets, nu and nubar, incl. the zero-flux rows."""
    from pisa_b200 import ops
    dev = _dev()
    g = load_golden("ref_flux_f8.npz")
    T = lambda k: torch.tensor(g[k], device=dev)  # noqa: E731
    e, cz, nu, nb = T("true_energy"), T("true_coszen"), T("nu_flux_nominal"), T("nubar_flux_nominal")
    terms = ops.flux_barr_terms(e, cz)
    assert terms.shape == (e.numel(), 4) and torch.isfinite(terms).all()
    for name in sorted({k.split("/")[0] for k in g.files if "/" in k}):
        pars = [float(x) for x in g[name + "/params"]]
        for nubar, tag in ((1, "nu"), (-1, "nubar")):
            out = ops.flux_barr_apply(terms, nu, nb, nubar, *pars).cpu().numpy()
            ref = g["%s/%s" % (name, tag)]
            assert np.allclose(out, ref, rtol=1e-10, atol=1e-300), (name, tag, np.abs(out - ref).max())
            direct = ops.flux_barr_simple(e, cz, nu, nb, nubar, *pars).cpu().numpy()
            assert np.allclose(out, direct, rtol=1e-13, atol=1e-300)
    # float32 storage of the fluxes, terms stay float64
    out4 = ops.flux_barr_apply(terms, nu.float(), nb.float(), -1, 0.95, 1.1, -0.1, -1.0, 1.0)
    assert out4.dtype == torch.float32 and np.allclose(out4.cpu().numpy(), g["all_down/nubar"], rtol=3e-5, atol=1e-6)
