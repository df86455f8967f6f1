"""GPU tests of the callers and loaders around the hot path (SURVEY 8f.4 and the fit-loop caller of 3.5):
``data.simple_data_loader``, ``DistributionMaker.get_outputs(return_sum=True)``, and ``utils.hist``'s
``apply_unc_weights`` / ``unweighted`` options through the fused template kernel."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from conftest import ROOT  # noqa: E402

CFG_DIR = os.path.join(ROOT, "pisa_b200", "resources", "settings", "pipeline")
NAMES = ["nue_cc", "numu_cc", "nutau_cc", "nue_nc", "numu_nc", "nutau_nc", "nuebar_cc", "numubar_cc", "nutaubar_cc",
         "nuebar_nc", "numubar_nc", "nutaubar_nc"]


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _events_cfg(tmp_path, loader_section, hist_extra="", order_head="data.simple_data_loader"):
    base = open(os.path.join(CFG_DIR, "b200_events.cfg")).read()
    head, rest = base.split("[data.synthetic_mc]")
    rest = rest[rest.index("[osc.prob3]"):]
    text = head.replace("data.synthetic_mc", order_head) + loader_section + "\n" + rest
    if hist_extra:
        text = text.replace("error_method = sumw2", "error_method = sumw2\n" + hist_extra)
    cfg = tmp_path / "pipeline.cfg"
    cfg.write_text(text)
    return str(cfg)


def test_simple_data_loader_pipeline_matches_oracle_chain(tmp_path):
    """data.simple_data_loader (PISA-style events file in the .npz layout, variable mapping with a stacked flux,
    mc_cuts, down-sampling) -> osc.prob3 -> aeff.aeff -> utils.hist, against the oracle chain on the loaded events."""
    _need_gpu()
    from pisa_b200.core.pipeline import Pipeline
    from pisa_b200.stages.data.simple_data_loader import apply_cut, load_events
    from pisa_b200.utils import synthetic as syn
    rng = np.random.default_rng(11)
    arrays = {}
    for name in NAMES:
        n = 4000 + 100 * len(name)
        e = 10 ** rng.uniform(0, 3, n)
        cz = rng.uniform(-1, 1, n)
        cols = dict(true_energy=e, true_coszen=cz, reco_energy=np.clip(e * rng.lognormal(0, 0.3, n), 5.7, 56.0),
                    reco_coszen=np.clip(cz + rng.normal(0, 0.2, n), -1, 0.999), pid=rng.integers(0, 2, n).astype(float),
                    weighted_aeff=rng.uniform(0, 1e-4, n), nominal_nue_flux=rng.uniform(0.5, 1.5, n),
                    nominal_numu_flux=rng.uniform(0.5, 1.5, n))
        for k, v in cols.items():
            arrays["%s/%s" % (name, k)] = v
    arrays["__metadata__/livetime"] = np.float64(2.5)
    path = tmp_path / "events.npz"
    np.savez(path, **arrays)
    data_dict = ("{'true_energy': 'true_energy', 'true_coszen': 'true_coszen', 'reco_energy': 'reco_energy', "
                 "'reco_coszen': 'reco_coszen', 'pid': 'pid', 'weighted_aeff': 'weighted_aeff', "
                 "'nu_flux': ['nominal_nue_flux', 'nominal_numu_flux']}")
    cut = "(true_coszen <= 0.5) & (true_energy <= 700)"
    loader = ("[data.simple_data_loader]\napply_mode = events\noutput_names = %s\nevents_file = %s\nmc_cuts = %s\n"
              "data_dict = %s\nfraction_events_to_keep = 0.5\nevents_subsample_index = 1\nrequired_metadata = livetime\n"
              % (", ".join(NAMES), path, cut, data_dict))
    pipe = Pipeline(_events_cfg(tmp_path, loader))
    assert pipe.stages[0].service_name == "simple_data_loader" and pipe.stages[0].metadata["livetime"] == 2.5
    out = pipe.get_outputs()
    import ast
    host, _ = load_events(str(path), ast.literal_eval(data_dict), fraction_events_to_keep=0.5, events_subsample_index=1)
    host = apply_cut(host, cut)
    o = pipe["prob3"].osc_params
    dm, mix, mat_pot = o.dm_matrix, o.mix_matrix_complex, pipe["prob3"].gen_mat_pot_matrix_complex
    L = oracle.OracleLayers(np.loadtxt(os.path.join(ROOT, "pisa_b200", "resources", "osc", "PREM_12layer.dat")), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    zc, zf = np.zeros((3, 3), dtype=complex), np.zeros((3, 3))
    livetime = 2.5 * 365 * 86400.0
    for c in pipe.data.containers:
        ev = host[c.name]
        c.representation = "events"
        assert c.size == len(ev["true_energy"]) and 0 < c.size < 0.5 * (4000 + 100 * len(c.name)) + 1
        assert float(ev["true_coszen"].max()) <= 0.5 and float(ev["true_energy"].max()) <= 700
        assert np.array_equal(c["nu_flux"].cpu().numpy(), ev["nu_flux"]) and ev["nu_flux"].shape[1] == 2
        # down-sampled: initial weights carry the inverse fraction (simple_data_loader.py:218-224)
        assert np.all(c["initial_weights"].cpu().numpy() == 2.0)
        nubar, flav = int(c["nubar"]), int(c["flav"])
        _, den, dis = L.calcLayers(ev["true_coszen"])
        prob = oracle.propagate_array(dm, mix, mat_pot, -1, zc, zf, nubar, ev["true_energy"], den, dis)
        w = 2.0 * (ev["nu_flux"][:, 0] * prob[:, 0, flav] + ev["nu_flux"][:, 1] * prob[:, 1, flav])
        w = w * (ev["weighted_aeff"] * (1.0 * livetime))
        ie = oracle.digitize_irregular(ev["reco_energy"], syn.DRAGON_E_EDGES)
        i2, _ = oracle.regular_index([ev["reco_coszen"], ev["pid"]], [-1.0, -0.5], [1.0, 1.5], [8, 2])
        idx = np.where((ie >= 0) & (ie < 8) & (i2 >= 0), ie * 16 + i2, -1)
        assert np.allclose(out[c.name].hist, oracle.accumulate(idx, w, 128).reshape(8, 8, 2), rtol=1e-10, atol=0), c.name
    # the two halves of a 50 % split are disjoint and together cover the file (independent sub-samples)
    a, _ = load_events(str(path), {"e": "true_energy"}, fraction_events_to_keep=0.5, events_subsample_index=0)
    b, _ = load_events(str(path), {"e": "true_energy"}, fraction_events_to_keep=0.5, events_subsample_index=1)
    full, _ = load_events(str(path), {"e": "true_energy"})
    for name in NAMES:
        assert len(np.intersect1d(a[name]["e"], b[name]["e"])) == 0
        assert np.array_equal(np.sort(np.concatenate([a[name]["e"], b[name]["e"]])), np.sort(full[name]["e"]))


def test_simple_data_loader_errors(tmp_path):
    _need_gpu()
    from pisa_b200.stages.data.simple_data_loader import simple_data_loader
    ev = {"numu_cc": {"true_energy": np.ones(4), "weights": np.ones(4)}}
    stage = simple_data_loader(events_file=ev, mc_cuts=None, data_dict=None, output_names=["numu_cc"])
    from pisa_b200.core.container import ContainerSet
    stage.data = ContainerSet("x")
    with pytest.raises(KeyError):          # a `weights` field in the file would be overwritten
        stage.setup()
    with pytest.raises(ValueError):        # duplicate output names
        simple_data_loader(events_file=ev, mc_cuts=None, data_dict=None, output_names=["numu_cc", "numu_cc"])
    with pytest.raises(ValueError):        # categories not split by flavour / interaction
        simple_data_loader(events_file={"numu": {"true_energy": np.ones(4)}}, mc_cuts=None, data_dict=None,
                           output_names=["numu"])
    with pytest.raises(ImportError):       # HDF5 needs h5py
        try:
            import h5py  # noqa: F401
            raise ImportError("h5py present: nothing to check")
        except ImportError:
            p = tmp_path / "x.hdf5"
            p.write_bytes(b"")
            simple_data_loader(events_file=str(p), mc_cuts=None, data_dict=None, output_names=["numu_cc"])


def test_distribution_maker_sum_and_param_updates():
    """DistributionMaker.get_outputs(return_sum=True) (distribution_maker.py:251-294) over two pipelines, fused and
    staged evaluation, update_params / select_params across pipelines."""
    _need_gpu()
    from pisa_b200.core.distribution_maker import DistributionMaker
    from pisa_b200.core.pipeline import Pipeline
    from pisa_b200.fused import FusedPipeline
    from pisa_b200.utils.units import ureg
    cfgs = ["settings/pipeline/b200_events.cfg", "settings/pipeline/b200_flux_events.cfg"]
    fused, staged = DistributionMaker(cfgs), DistributionMaker(cfgs, fused=False)
    assert all(isinstance(e, FusedPipeline) for e in fused.evaluators)
    assert all(isinstance(e, Pipeline) for e in staged.evaluators)

    def check():
        tot_f, tot_s = fused.get_outputs(return_sum=True), staged.get_outputs(return_sum=True)
        assert len(tot_f) == 1 and tot_f.names == ["total"]
        parts = staged.get_outputs()
        manual = sum(m.hist for ms in parts for m in ms)
        assert np.allclose(tot_s["total"].hist, manual, rtol=1e-13, atol=0)
        assert np.allclose(tot_f["total"].hist, tot_s["total"].hist, rtol=1e-10, atol=0)
        err = np.sqrt(sum(m.std_devs ** 2 for ms in parts for m in ms))
        assert np.allclose(tot_f["total"].std_devs, err, rtol=1e-10, atol=0)
        return tot_f["total"].hist

    first = check()
    for dm in (fused, staged):
        p = dm.params.theta23
        p.value = 48.0 * ureg.deg
        dm.update_params(p)
    second = check()
    assert not np.allclose(first, second, rtol=1e-6)
    for dm in (fused, staged):
        dm.select_params("ih")
        assert all(pl.params.deltam31.value.m < 0 for pl in dm)
    check()
    with pytest.raises(KeyError):
        fused.select_params("no_such_selection")
    fused.select_params("no_such_selection", error_on_missing=False)


@pytest.mark.parametrize("hist_extra,loader_extra", [("apply_unc_weights = True", "unc_weights = True"),
                                                      ("unweighted = True", ""),
                                                      ("unweighted = True\napply_unc_weights = True", "unc_weights = True")])
def test_fused_pipeline_with_hist_options(tmp_path, hist_extra, loader_extra):
    """utils.hist's apply_unc_weights / unweighted (hist.py:141-145,198-209) through FusedPipeline: weights, errors and
    bin_unc2 equal the staged pipeline's."""
    _need_gpu()
    from pisa_b200.core.pipeline import Pipeline
    from pisa_b200.fused import FusedPipeline
    from pisa_b200.utils.units import ureg
    base = open(os.path.join(CFG_DIR, "b200_events.cfg")).read()
    base = base.replace("error_method = sumw2", "error_method = sumw2\n" + hist_extra)
    if loader_extra:
        base = base.replace("param.n_events = 20000", loader_extra + "\nparam.n_events = 20000")
    cfg = tmp_path / "pipeline.cfg"
    cfg.write_text(base)
    staged, fused = Pipeline(str(cfg)), FusedPipeline(Pipeline(str(cfg)))
    for theta in (42.3, 49.0):
        for p in (staged, fused.pipeline):
            p.params.theta23 = theta * ureg.deg
        staged.run()
        fused.run()
        for cs, cf in zip(staged.data.containers, fused.pipeline.data.containers):
            cs.representation = cf.representation = staged.output_binning
            for key in ("weights", "errors", "bin_unc2"):
                a, b = cs[key].cpu().numpy(), cf[key].cpu().numpy()
                assert np.allclose(b, a, rtol=1e-10, atol=0), (cs.name, key, hist_extra)
            if "unweighted" in hist_extra and not loader_extra:
                assert float(cs["weights"].sum()) == cs["weights"].sum().round().item()   # plain counts


def test_fused_pipeline_floating_flux_systematics_and_astro_weights():
    """flux.barr_simple in front of osc.prob3: FusedPipeline evaluates the five flux systematics inside the template
    kernel (no engine rebuild, nu_flux not rewritten) and matches the staged pipeline; an additive ``astro_weights``
    term (hist.py:141-145) on the containers goes through the kernel as well."""
    _need_gpu()
    from pisa_b200.core.pipeline import Pipeline
    from pisa_b200.fused import FusedPipeline
    from pisa_b200.utils.units import ureg
    cfg = "settings/pipeline/b200_flux_events.cfg"
    staged, fused = Pipeline(cfg), FusedPipeline(Pipeline(cfg))
    assert fused.barr is not None and all(s.service_name != "barr_simple" for s in fused.pre)

    def compare(tag):
        staged.run()
        fused.run()
        for cs, cf in zip(staged.data.containers, fused.pipeline.data.containers):
            cs.representation = cf.representation = staged.output_binning
            for key in ("weights", "errors"):
                a, b = cs[key].cpu().numpy(), cf[key].cpu().numpy()
                assert np.allclose(b, a, rtol=1e-10, atol=0), (tag, cs.name, key)
        return np.stack([c["weights"].cpu().numpy() for c in fused.pipeline.data.containers])

    first = compare("nominal")
    engine = fused._engine
    for p in (staged, fused.pipeline):
        p.params.delta_index = 0.08
        p.params.Barr_uphor_ratio = -0.7
        p.params.nu_nubar_ratio = 1.06
        p.params.theta23 = 47.0 * ureg.deg
    second = compare("flux systematics moved")
    assert fused._engine is engine and engine._flux_stale     # no rebuild, nu_flux not rewritten
    assert not np.allclose(first, second, rtol=1e-4)
    # additive astrophysical term: both pipelines get the same per-event array
    for p in (staged, fused.pipeline):
        for c in p.data.containers:
            c.representation = "events"
            g = torch.Generator(device="cpu").manual_seed(len(c.name))
            c["astro_weights"] = (torch.rand(c.size, generator=g, dtype=torch.float64) * 1e-3).to(c["weights"])
    fused._engine = None                                       # event arrays changed: new engine
    third = compare("astro_weights")
    assert (third.sum(axis=1) > second.sum(axis=1)).all()
    for p in (staged, fused.pipeline):
        p.params.Barr_nu_nubar_ratio = 0.9
    compare("astro_weights + flux systematics")


def test_hypersurface_stage_and_epilogue_vs_reference_golden():
    """discr_sys.hypersurfaces against the UNMODIFIED reference stage (compute_function + apply_function run on binned
    containers by tests/golden/make_golden_hypersurfaces.py): weights (clipped at 0), errors, bin_unc2 after the stage,
    and the same per-bin scales applied by the fit-loop epilogue kernel (pisab_reweight_hist_chi2's ``bin_scales``:
    sum w -> max(s sum w, 0), sum w^2 -> s^2 sum w^2, i.e. errors -> |s| errors)."""
    _need_gpu()
    from pisa_b200 import ops
    from pisa_b200.core.container import Container, ContainerSet
    from pisa_b200.core.param import Param, ParamSet
    from pisa_b200.stages.discr_sys.hypersurfaces import hypersurfaces
    from pisa_b200.utils.config_parser import parse_pipeline_config
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_hypersurfaces_f8.npz"))
    names, maps = list(gold["param_names"]), list(gold["map_names"])
    binning = parse_pipeline_config("settings/pipeline/b200_icecube3y_full.cfg")[("discr_sys", "hypersurfaces")]["calc_mode"]
    dev = torch.device("cuda:0")
    for ci, values in enumerate(gold["param_values"]):
        data = ContainerSet("hs")
        for m in maps:
            c = Container(m, representation=binning)
            for key in ("weights", "errors", "bin_unc2"):
                c[key] = gold["in_%s_%d_%s" % (key, ci, m)]
            data.add_container(c)
        st = hypersurfaces(fit_results_file="events/IceCube_3y_oscillations/hyperplanes_*.csv.bz2",
                           params=ParamSet([Param(name=n, value=float(v)) for n, v in zip(names, values)]),
                           data=data, calc_mode=binning, apply_mode=binning, error_method="sumw2")
        st.setup()
        st.run()
        for m, c in zip(maps, data.containers):
            c.representation = binning
            for key in ("weights", "errors", "bin_unc2", "hs_scales"):
                want = gold["out_%s_%d_%s" % (key, ci, m)]
                got = c[key].cpu().numpy() if isinstance(c[key], torch.Tensor) else np.asarray(c[key])
                assert np.array_equal(got, want), (ci, m, key)
        # the fit-loop epilogue with the same scales: feed the input maps as one-block "partials" of 4 containers
        partial = torch.stack([torch.stack([torch.tensor(gold["in_weights_%d_%s" % (ci, m)], device=dev),
                                            torch.tensor(gold["in_errors_%d_%s" % (ci, m)], device=dev) ** 2]) for m in maps])
        scales = torch.tensor(np.stack([gold["out_hs_scales_%d_%s" % (ci, m)] for m in maps]), device=dev)
        out = ops.hist_reduce_chi2(partial.contiguous(), 1, scales)
        for k, m in enumerate(maps):
            assert np.array_equal(out[k, 0].cpu().numpy(), gold["out_weights_%d_%s" % (ci, m)])
            assert np.allclose(np.sqrt(out[k, 1].cpu().numpy()), np.abs(gold["out_errors_%d_%s" % (ci, m)]),
                               rtol=4e-16, atol=0)


def test_one_call_template_chi2_with_bin_scales():
    """pisab_reweight_hist_chi2: template kernel + ONE epilogue kernel (reduce, per-bin hypersurface scales, container
    sum, mod_chi2 in the last-arriving block) against the separate launches; repeated calls reuse the arrival counter."""
    _need_gpu()
    from pisa_b200 import ops
    from pisa_b200.engine import ReweightEngine
    from pisa_b200.stages.osc.layers import Layers
    from pisa_b200.utils import synthetic as syn
    dev = torch.device("cuda:0")
    L = Layers(os.path.join(ROOT, "pisa_b200", "resources", syn.EARTH["earth_model"]), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = L.earth_struct()
    binning, _keep = ops.make_binning(syn.DRAGON_DIMS, dev)
    eng = ReweightEngine(earth, 128, np.float64, dev)
    for c, (name, nubar, flav) in enumerate(syn.CONTAINERS):
        ev = syn.make_events_torch(9000 + 500 * c, seed=60 + c, dtype=np.float64, device=dev)
        idx = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
        eng.add_container(name, nubar, flav, ev["true_energy"], ev["true_coszen"], ev["nu_flux"], ev["weights"], idx)

    def consts(theta23):
        dm, mix, mp = syn.osc_matrices(dict(syn.NUFIT20_NH, theta23=theta23))
        return ops.OscConsts.from_matrices(dm, mix, mp)
    observed = eng.evaluate(consts(45.0))[:, 0].sum(dim=0).contiguous()
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    scales = (0.7 + 0.6 * torch.rand((12, 128), generator=g, device=dev, dtype=torch.float64))
    scales[3, 5] = -0.2                                    # a negative scale: weights clipped at 0, errors scaled by |s|
    for theta in (40.0, 45.0, 51.0, 40.0):
        c = consts(theta)
        plain = eng.evaluate(c).clone()
        want = ops.template_chi2(plain, observed).clone()
        hist, chi2 = eng.evaluate_chi2(c, observed)
        assert torch.equal(hist, plain) and torch.equal(chi2, want), theta      # same reductions, same order
        hist_s, chi2_s = eng.evaluate_chi2(c, observed, bin_scales=scales)
        ref = plain.clone()
        ref[:, 0] = torch.clamp(plain[:, 0] * scales, min=0.0)
        ref[:, 1] = plain[:, 1] * scales * scales
        assert torch.allclose(hist_s, ref, rtol=1e-15, atol=0)
        assert torch.allclose(chi2_s, ops.template_chi2(ref, observed), rtol=1e-13, atol=0)
    assert float(eng.evaluate_chi2(consts(45.0), observed)[1]) < 1e-20           # Asimov point
