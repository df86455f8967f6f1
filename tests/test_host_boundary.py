"""CPU-only tests of the host-side boundary: units, params, binning, cfg parsing, Stage contract,
and that the C-ABI library loads and exports every symbol the header declares."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from pisa_b200.core.binning import MultiDimBinning, OneDimBinning
from pisa_b200.core.param import Param, ParamSelector, ParamSet
from pisa_b200.utils.config_parser import parse_pipeline_config
from pisa_b200.utils.units import Quantity, parse_quantity, ureg

REF_RES = "/root/reference/pisa_examples/resources"


def test_abi_library_exports_every_declared_symbol():
    from pisa_b200 import _lib
    lib = _lib.load()   # loading must work without a GPU
    header = open(os.path.join(ROOT, "include", "pisa_b200.h")).read()
    declared = sorted(set(re.findall(r"^(?:int|int64_t|double|const char \*)\s*(pisab_[a-z0-9_]+)\s*\(", header, re.M)))
    assert len(declared) >= 29
    exported = subprocess.check_output(["nm", "-D", "--defined-only", _lib.lib_path()]).decode()
    for name in declared:
        assert re.search(r"\bT %s\b" % name, exported), "symbol %s declared in the header but not exported" % name
        assert hasattr(lib, name)
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert b"sm_100a" in lib.pisab_version()
    # SASS is sm_100a only (no multi-arch fat binary, no PTX fallback for other targets)
    archs = set(re.findall(r"arch = (sm_\d+a?)", subprocess.check_output(["cuobjdump", "-lelf", _lib.lib_path()]).decode()
                           + subprocess.run(["cuobjdump", "-sass", _lib.lib_path()], capture_output=True, text=True).stdout[:200000]))
    assert archs <= {"sm_100a"} and archs


def test_no_cpu_path():
    import torch
    from pisa_b200 import ops
    with pytest.raises(TypeError):
        ops.hist_accumulate(torch.zeros(4, dtype=torch.int32), torch.zeros(4, dtype=torch.float64), 8)
    # the product never imports the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pisa_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_units_and_quantity_parsing():
    q, s = parse_quantity("8.5 +/- 0.205 units.deg")
    assert q.units == "deg" and q.m == 8.5 and s == 0.205
    assert np.isclose(q.m_as("rad"), np.deg2rad(8.5))
    q, s = parse_quantity("2.457e-3 units.eV**2")
    assert q.m_as("eV**2") == 2.457e-3 and np.isnan(s)
    q, _ = parse_quantity("42. * units.degree")
    assert q.m_as("deg") == 42.0
    assert (2.5 * ureg.common_year).m_as("sec") == 2.5 * 365 * 86400
    assert parse_quantity("1e4")[0].units == "dimensionless"
    with pytest.raises(ValueError):
        parse_quantity("osc/PREM_12layer.dat")
    with pytest.raises(ValueError):
        (1.0 * ureg.km).to("deg")
    assert isinstance(np.array([0., 90.]) * ureg.degree, Quantity)


def test_paramset_selector_and_hash():
    a = Param("theta23", 42 * ureg.deg, is_fixed=False)
    b = Param("deltam31", 2.457e-3 * ureg.eV ** 2)
    bi = Param("deltam31", -2.374e-3 * ureg.eV ** 2)
    sel = ParamSelector(regular_params=[a], selector_param_sets={"nh": [b], "ih": [bi]}, selections=["nh"])
    ps = sel.params
    assert ps.names == ("theta23", "deltam31") and ps.deltam31.value.m == 2.457e-3
    h0 = ps.values_hash
    ps.theta23.value = 45 * ureg.deg
    assert ps.values_hash != h0 and ps.free.names == ("theta23",)
    ps.theta23.value = 42 * ureg.deg
    assert ps.values_hash == h0
    sel.select_params(["ih"])
    assert ps.deltam31.value.m == -2.374e-3     # same ParamSet object the stage holds
    with pytest.raises(ValueError):
        ps.theta23.value = 3 * ureg.km


def test_param_selector_update_survives_reselection():
    """Pipeline.update_params goes through ParamSelector.update (pipeline.py:579-596 / param.py:1708-1730 of the
    reference): the regular set, the current set and the ACTIVE selector sets take the new value, so selecting again
    does not revert it; selections are applied in order and a missing one only raises when asked to."""
    t23 = Param("theta23", 42 * ureg.deg, is_fixed=False)
    nh = Param("deltam31", 2.457e-3 * ureg.eV ** 2)
    ih = Param("deltam31", -2.374e-3 * ureg.eV ** 2)
    sel = ParamSelector(regular_params=[t23], selector_param_sets={"nh": [nh], "ih": [ih]}, selections=["nh"])
    sel.update([Param("theta23", 47 * ureg.deg, is_fixed=False), Param("deltam31", 2.6e-3 * ureg.eV ** 2)])
    assert sel.params.theta23.value.m == 47 and sel.params.deltam31.value.m == 2.6e-3
    sel.select_params(["ih"])
    assert sel.params.deltam31.value.m == -2.374e-3 and sel.params.theta23.value.m == 47   # regular update kept
    sel.select_params(["nh"])
    assert sel.params.deltam31.value.m == 2.6e-3            # the active selector set was updated too
    sel.update(Param("not_mine", 1.0), extend=False)        # excess params of other stages are ignored
    assert "not_mine" not in sel.params.names
    sel.select_params(["ih", "no_such"], error_on_missing=False)
    assert sel.params.deltam31.value.m == -2.374e-3         # the available selection was applied
    with pytest.raises(KeyError):
        sel.select_params(["nh", "no_such"], error_on_missing=True)
    assert sel.params.deltam31.value.m == 2.6e-3            # ... in order, before the missing one raised


def test_simple_data_loader_host_logic(tmp_path):
    """File reading (.npz layout), variable mapping with stacking, cuts and the reference's sub-sampling algorithm
    (events_pi.py:175-505), without a device."""
    from pisa_b200 import FTYPE
    from pisa_b200.stages.data.simple_data_loader import apply_cut, load_events
    rng = np.random.RandomState(3)
    arrays = {}
    for cat in ("nue_cc", "numubar_nc"):
        arrays[cat + "/true_energy"] = 1 + 79 * rng.rand(1000)
        arrays[cat + "/true_coszen"] = 2 * rng.rand(1000) - 1
        arrays[cat + "/fa"], arrays[cat + "/fb"] = rng.rand(1000), rng.rand(1000)
    arrays["__metadata__/livetime"] = np.float64(3.0)
    path = str(tmp_path / "ev.npz")
    np.savez(path, **arrays)
    mapping = {"true_energy": "true_energy", "true_coszen": "true_coszen", "flux": ["fa", "fb"]}
    ev, meta = load_events([path, path], mapping, required_metadata=["livetime"])
    assert meta == {"livetime": 6.0}                                   # livetimes of several files add up
    assert ev["nue_cc"]["flux"].shape == (2000, 2) and ev["nue_cc"]["flux"].dtype == FTYPE
    assert np.array_equal(ev["nue_cc"]["flux"][:1000, 1], arrays["nue_cc/fb"].astype(FTYPE))
    cut = apply_cut(ev, "(true_coszen <= 0.5) & (np.log10(true_energy) < 1.5)")
    keep = (ev["nue_cc"]["true_coszen"] <= 0.5) & (ev["nue_cc"]["true_energy"] < 10 ** 1.5)
    assert np.array_equal(cut["nue_cc"]["true_energy"], ev["nue_cc"]["true_energy"][keep])
    # sub-samples: reproducible, disjoint, the same choice for every variable of a category
    a, _ = load_events(path, mapping, fraction_events_to_keep=0.25, events_subsample_index=0)
    a2, _ = load_events(path, mapping, fraction_events_to_keep=0.25, events_subsample_index=0)
    b, _ = load_events(path, mapping, fraction_events_to_keep=0.25, events_subsample_index=2)
    assert len(a["nue_cc"]["true_energy"]) == 250 and np.array_equal(a["nue_cc"]["flux"], a2["nue_cc"]["flux"])
    assert len(np.intersect1d(a["nue_cc"]["true_energy"], b["nue_cc"]["true_energy"])) == 0
    full = arrays["nue_cc/true_energy"].astype(FTYPE)
    pos = np.searchsorted(np.sort(full), a["nue_cc"]["true_energy"])
    order = np.argsort(full)[pos]
    assert np.array_equal(a["nue_cc"]["true_coszen"], arrays["nue_cc/true_coszen"].astype(FTYPE)[order])
    with pytest.raises(KeyError):
        load_events(path, {"x": "no_such_variable"})
    with pytest.raises(ValueError):
        load_events(path, mapping, fraction_events_to_keep=0.5, events_subsample_index=2)


def test_binning_classification_is_ftype_dependent_rule():
    """dragon_datarelease.reco_energy has 9-digit edges whose ratios differ by 5.5e-10,
    so in FP64 (rtol 1e-12) it is IRREGULAR -> searchsorted on the real edges; under PISA_FTYPE=fp32 (rtol 1e-5) the
    same axis is regular-log."""
    from pisa_b200 import FTYPE
    e = OneDimBinning("reco_energy", is_log=True, bin_edges=[5.62341325, 7.49894209, 10.0, 13.33521432, 17.7827941,
                                                             23.71373706, 31.6227766, 42.16965034, 56.23413252] * ureg.GeV)
    assert e.is_irregular == (FTYPE == np.float64) and e.is_log and e.num_bins == 8
    t = OneDimBinning("true_energy", num_bins=200, is_log=True, domain=[1., 1000] * ureg.GeV)
    c = OneDimBinning("true_coszen", num_bins=200, is_lin=True, domain=[-1, 1])
    assert not t.is_irregular and not c.is_irregular
    assert np.array_equal(t.bin_edges.m, np.logspace(0, 3, 201, dtype=FTYPE))
    assert np.allclose(t.weighted_centers.m, np.sqrt(t.bin_edges.m[:-1] * t.bin_edges.m[1:]), rtol=0, atol=0)
    m = MultiDimBinning([t, c], name="calc_grid")
    assert m.shape == (200, 200) and m.size == 40000 and m.names == ["true_energy", "true_coszen"]
    assert hash(m) == hash(MultiDimBinning([t, c])) and m == MultiDimBinning([t, c])
    g = m.meshgrid("weighted_centers", attach_units=False)
    assert g[0].shape == (200, 200) and g[0][3, 0] == g[0][3, 7] and g[1][0, 5] == g[1][9, 5]  # 'ij', row-major
    pid = OneDimBinning("pid", bin_edges=[-np.inf, 0.55, np.inf])
    assert pid.is_irregular


def test_reference_cfg_copies_are_byte_identical():
    """tests/golden/ref_cfg holds the cfg TEXT of the reference's README example (config text, not code) so that the
    GPU box, which has no reference tree, can run it; here -- where the tree exists -- the copies are checked byte for
    byte, and they must parse to the reference's stage order on their own."""
    cfg_root = os.path.join(ROOT, "tests", "golden", "ref_cfg")
    rels = ["settings/pipeline/osc_example.cfg", "settings/binning/example.cfg", "settings/osc/nufitv20.cfg",
            "settings/osc/earth.cfg"]
    for rel in rels:
        assert os.path.exists(os.path.join(cfg_root, rel)), rel
        if os.path.isdir(REF_RES):
            assert open(os.path.join(cfg_root, rel), "rb").read() == open(os.path.join(REF_RES, rel), "rb").read(), rel
    old = os.environ.get("PISA_RESOURCES")
    os.environ["PISA_RESOURCES"] = cfg_root
    try:
        d = parse_pipeline_config("settings/pipeline/osc_example.cfg")
        assert list(d.keys())[1:] == [("data", "toy_event_generator"), ("flux", "barr_simple"), ("osc", "prob3")]
        assert d[("osc", "prob3")]["calc_mode"].shape == (200, 200)
    finally:
        if old is None:
            del os.environ["PISA_RESOURCES"]
        else:
            os.environ["PISA_RESOURCES"] = old


def test_parse_own_and_reference_pipeline_cfgs():
    d = parse_pipeline_config("settings/pipeline/b200_events.cfg")
    assert list(d.keys())[1:] == [("data", "synthetic_mc"), ("osc", "prob3"), ("aeff", "aeff"), ("utils", "hist")]
    assert d["pipeline"]["output_key"] == ("weights", "errors") and d["pipeline"]["output_binning"].shape == (8, 8, 2)
    p = d[("osc", "prob3")]["params"].params
    assert set(p.names) == {"earth_model", "YeI", "YeM", "YeO", "detector_depth", "prop_height", "theta12", "theta13",
                            "theta23", "deltam21", "deltam31", "deltacp"}
    assert p.theta13.value.m_as("deg") == 8.5 and p.deltam31.value.m_as("eV**2") == 2.457e-3  # nh selected
    assert not p.theta23.is_fixed and p.theta12.is_fixed and p.earth_model.value == "osc/PREM_12layer.dat"
    if not os.path.isdir(REF_RES):
        pytest.skip("reference tree not present")
    old = os.environ.get("PISA_RESOURCES")
    os.environ["PISA_RESOURCES"] = REF_RES
    try:
        # byte-identical reference cfgs (README example + the IceCube 3y pipeline)
        d = parse_pipeline_config("settings/pipeline/osc_example.cfg")
        assert list(d.keys())[1:] == [("data", "toy_event_generator"), ("flux", "barr_simple"), ("osc", "prob3")]
        assert d[("osc", "prob3")]["calc_mode"].shape == (200, 200)
        p = d[("osc", "prob3")]["params"].params
        assert p.theta23.value.m_as("deg") == 42.0 and p.deltacp.value.m_as("deg") == 0.0
        assert np.allclose(p.theta13.range.m, [7.85, 9.1], rtol=1e-6) and p.theta13.prior["kind"] == "gaussian"
        d = parse_pipeline_config("settings/pipeline/IceCube_3y_neutrinos.cfg")
        assert d[("utils", "hist")]["error_method"] == "sumw2"
        assert d[("osc", "prob3")]["calc_mode"].shape == (200, 200) and d[("osc", "prob3")]["apply_mode"] == "events"
        assert d[("aeff", "aeff")]["params"].params.livetime.value.m_as("common_year") == 2.5
    finally:
        if old is None:
            del os.environ["PISA_RESOURCES"]
        else:
            os.environ["PISA_RESOURCES"] = old


def test_stage_contract_errors():
    from pisa_b200.stages.osc.prob3 import init_test, prob3
    from pisa_b200.stages.utils.hist import hist
    s = init_test()
    assert (s.stage_name, s.service_name) == ("osc", "prob3")
    assert s.has_setup and s.has_compute and s.has_apply and s.param_hash is None
    good = s.params
    with pytest.raises(ValueError, match="Missing params"):
        prob3(params=ParamSet([p for p in good if p.name != "theta12"]))
    with pytest.raises(ValueError, match="Excess params"):
        prob3(params=ParamSet(list(good) + [Param("bogus", 1.0)]))
    with pytest.raises(ValueError):
        prob3(params=good, nsi_type="quantum")
    # neutrino_decay=True adds decay_alpha3 (prob3.py:256-259) and selects the decay branch (:227-230)
    with pytest.raises(ValueError, match="decay_alpha3"):
        prob3(params=good, neutrino_decay=True)
    from pisa_b200.utils.units import ureg
    d = prob3(params=ParamSet(list(good) + [Param("decay_alpha3", 1e-4 * ureg.eV ** 2)]), neutrino_decay=True)
    assert d.decay_flag == 1 and d.neutrino_decay
    with pytest.raises(ValueError, match="not supported"):
        hist(calc_mode="log_events").setup()
    h = hist(calc_mode="events")
    h.data = "not a container set"
    with pytest.raises(TypeError):
        h.setup()
    # nsi_type='standard' adds the nine eps_* names (prob3.py:244-254), 'vacuum-like' its eight (:234-243)
    with pytest.raises(ValueError, match="eps_ee"):
        prob3(params=good, nsi_type="standard")
    with pytest.raises(ValueError, match="eps_scale"):
        prob3(params=good, nsi_type="vacuum-like")
    with pytest.raises(ValueError, match="v_lri"):
        prob3(params=good, lri_type="emu-symmetry")
    with pytest.raises(ValueError, match="not available"):
        prob3(params=good, lri_type="ee-symmetry")
    with pytest.raises(ValueError, match="density_scale"):
        prob3(params=good, tomography_type="mass_of_earth")
    with pytest.raises(ValueError, match="not available"):
        prob3(params=good, tomography_type="mass_of_moon")


def test_csv_loader_event_selection(tmp_path):
    """data.csv_loader host logic (pisa/stages/data/csv_loader.py:112-141): PDG / interaction-type masks and
    constructor validation; the upload to device containers is covered by the GPU pipeline test."""
    import pandas as pd
    from pisa_b200.stages.data.csv_loader import csv_loader, select_events, species_of
    rng = np.random.default_rng(0)
    n = 400
    pdg = rng.choice([12, -12, 14, -14, 16, -16], n)
    df = pd.DataFrame(dict(pdg=pdg, type=rng.integers(0, 3, n), true_energy=rng.uniform(1, 100, n),
                           true_coszen=rng.uniform(-1, 1, n), weight=rng.uniform(0, 1, n)))
    total = 0
    for name in ["nue_cc", "numu_cc", "nutau_cc", "nue_nc", "numu_nc", "nutau_nc", "nuebar_cc", "numubar_cc",
                 "nutaubar_cc", "nuebar_nc", "numubar_nc", "nutaubar_nc"]:
        nubar, flav = species_of(name)
        ev = select_events(df, name)
        assert (ev["pdg"] == nubar * (12 + 2 * flav)).all()
        assert (ev["type"] >= 1).all() if "cc" in name else (ev["type"] == 0).all()
        total += len(ev)
    assert total == n                                        # the twelve containers partition the file
    with pytest.raises(ValueError):
        select_events(df.drop(columns=["pdg"]), "nue_cc")
    f = tmp_path / "events.csv"
    df.to_csv(f, index=False)
    with pytest.raises(ValueError):
        csv_loader(events_file=str(f), data_dict=3, output_names="nue_cc")
    with pytest.raises(ValueError):
        csv_loader(events_file=str(f), data_dict="{'true_energy': 'true_energy'}", output_names="nue_cc, nue_cc")
    st = csv_loader(events_file=str(f), data_dict="{'true_energy': 'true_energy', 'weighted_aeff': 'weight'}",
                    output_names="nue_cc, numubar_nc", calc_mode="events", apply_mode="events")
    assert st.output_names == ["nue_cc", "numubar_nc"] and st.data_dict["weighted_aeff"] == "weight"


def test_hyperplane_loader_and_formula():
    """discr_sys.hypersurfaces host side: data-release CSV loader (hypersurface.py:2065-2172) and the linear
    hyperplane scale = offset + sum_p gradient_p * value_p (:421-428), checked against pandas directly."""
    import pandas as pd
    from pisa_b200.stages.discr_sys.hypersurfaces import evaluate_hyperplane, load_hypersurfaces_data_release
    from pisa_b200.utils.config_parser import parse_pipeline_config
    from pisa_b200.utils.resources import find_resource
    cfg = parse_pipeline_config("settings/pipeline/b200_icecube3y_full.cfg")
    binning = cfg[("discr_sys", "hypersurfaces")]["calc_mode"]
    hs, names = load_hypersurfaces_data_release("events/IceCube_3y_oscillations/hyperplanes_*.csv.bz2", binning)
    assert names == ["ice_absorption", "ice_scattering", "opt_eff_headon", "opt_eff_lateral", "opt_eff_overall"]
    assert list(hs) == ["nue_cc+nuebar_cc", "numu_cc+numubar_cc", "nutau_cc+nutaubar_cc", "nu_nc+nubar_nc"]
    vals = dict(opt_eff_overall=1.07, opt_eff_lateral=31.0, opt_eff_headon=-0.4, ice_scattering=2.5, ice_absorption=-3.0)
    raw = pd.read_csv(find_resource("events/IceCube_3y_oscillations/hyperplanes_numu_cc.csv.bz2"))
    ref = raw["offset"].values.copy()
    for p in names:
        ref += raw[p].values * vals[p]
    got = evaluate_hyperplane(hs["numu_cc+numubar_cc"], vals)
    assert got.shape == (8, 8, 2) and np.allclose(got.ravel(), ref, rtol=1e-15)
    # rows are ordered (reco_energy, reco_coszen, pid) row-major: the first rows of the file are the first energy bin
    assert raw["reco_energy"].values[:16].std() == 0 and set(raw["pid"].values[:2]) == {0, 1}
    other = parse_pipeline_config("settings/pipeline/b200_oscillogram.cfg")[("osc", "prob3")]["calc_mode"]
    with pytest.raises((AssertionError, KeyError)):      # a binning the files were not made for
        load_hypersurfaces_data_release("events/IceCube_3y_oscillations/hyperplanes_*.csv.bz2", other)


def test_hyperplanes_vs_reference_golden():
    """Loader + evaluation against outputs of the UNMODIFIED reference (``_load_hypersurfaces_data_release`` +
    ``Hypersurface.evaluate`` run by tests/golden/make_golden_hypersurfaces.py): bit-identical scale factors for all
    four maps and seven parameter points, and the stage's ``hs_scales`` (non-finite -> 1)."""
    from pisa_b200.stages.discr_sys.hypersurfaces import evaluate_hyperplane, load_hypersurfaces_data_release
    from pisa_b200.utils.config_parser import parse_pipeline_config
    gold = np.load(os.path.join(ROOT, "tests", "golden", "ref_hypersurfaces_f8.npz"))
    cfg = parse_pipeline_config("settings/pipeline/b200_icecube3y_full.cfg")
    binning = cfg[("discr_sys", "hypersurfaces")]["calc_mode"]
    hs, names = load_hypersurfaces_data_release("events/IceCube_3y_oscillations/hyperplanes_*.csv.bz2", binning)
    assert names == list(gold["param_names"]) and list(hs) == list(gold["map_names"])
    for ci, values in enumerate(gold["param_values"]):
        vals = dict(zip(names, values))
        for k, m in enumerate(hs):
            got = evaluate_hyperplane(hs[m], vals)
            assert np.array_equal(got, gold["scales_%d" % ci][k]), (ci, m)
            want = gold["out_hs_scales_%d_%s" % (ci, m)]
            got = got.reshape(-1).copy()
            got[~np.isfinite(got)] = 1.0
            assert np.array_equal(got, want)
    assert min(gold["scales_%d" % ci].min() for ci in range(len(gold["param_values"]))) < 0   # the clip is exercised


def test_mapset_json_interchange(tmp_path):
    """MapSet.to_json / from_json in the reference's layout (map.py:1272-1362,2206-2262; binning.py:676-694,
    1842-1859; jsons.py:196-330): state keys, nested lists, .json and .json.bz2, and a file shaped like real
    PISA's output (pint long unit names, integer hashes) reads back."""
    import bz2, json
    from pisa_b200.core.binning import MultiDimBinning, OneDimBinning
    from pisa_b200.core.map import Map, MapSet
    from pisa_b200.utils import jsons
    b = MultiDimBinning([OneDimBinning("reco_energy", num_bins=4, is_log=True, domain=[1, 100], units="GeV"),
                         OneDimBinning("reco_coszen", num_bins=3, is_lin=True, domain=[-1, 1])], name="reco")
    rng = np.random.default_rng(0)
    ms = MapSet([Map("nue_cc", rng.random((4, 3)), b, error_hist=rng.random((4, 3))),
                 Map("numu_cc", rng.random((4, 3)), b)], name="template")
    for ext in ("json", "json.bz2"):
        path = str(tmp_path / ("maps." + ext))
        ms.to_json(path)
        raw = open(path, "rb").read()
        state = json.loads(bz2.decompress(raw) if ext.endswith("bz2") else raw)
        assert list(state) == ["maps", "name", "tex", "collate_by_name"]
        assert list(state["maps"][0]) == ["name", "hist", "binning", "error_hist", "hash", "tex", "full_comparison"]
        assert list(state["maps"][0]["binning"]) == ["dimensions", "name", "mask"]
        assert list(state["maps"][0]["binning"]["dimensions"][0]) == ["name", "bin_edges", "units", "is_log", "is_lin",
                                                                     "bin_names", "tex"]
        assert state["maps"][1]["error_hist"] is None                  # all-zero errors are written as null
        back = MapSet.from_json(path)
        assert back.names == ms.names and back.name == "template"
        for m0, m1 in zip(ms, back):
            assert np.array_equal(m0.hist, m1.hist) and np.array_equal(m0.std_devs, m1.std_devs)   # repr round trip
            assert m0.binning == m1.binning and m1.binning.dimensions[0].is_log
    with pytest.raises(ValueError):
        ms.to_json(str(tmp_path / "maps.txt"))
    # a file as the reference writes it
    ref_like = {"maps": [{"name": "nutau_cc", "hist": [[1.5, 2.5], [3.5, 4.5]],
                          "binning": {"dimensions": [
                              {"name": "true_energy", "bin_edges": [1.0, 10.0, 100.0], "units": "gigaelectron_volt",
                               "is_log": True, "is_lin": False, "bin_names": None, "tex": r"E_{\rm true}"},
                              {"name": "true_coszen", "bin_edges": [-1.0, 0.0, 1.0], "units": "dimensionless",
                               "is_log": False, "is_lin": True, "bin_names": None, "tex": None}],
                              "name": None, "mask": None},
                          "error_hist": [[0.1, 0.2], [0.3, 0.4]], "hash": -1234567890123, "tex": None,
                          "full_comparison": False}],
                "name": "dist", "tex": None, "collate_by_name": True}
    path = str(tmp_path / "ref_like.json")
    open(path, "w").write(json.dumps(ref_like, indent=2))
    got = MapSet.from_json(path)
    assert got["nutau_cc"].hist.tolist() == [[1.5, 2.5], [3.5, 4.5]] and got["nutau_cc"].std_devs[1, 1] == 0.4
    assert got["nutau_cc"].binning.names == ["true_energy", "true_coszen"] and got["nutau_cc"].hash == -1234567890123
    assert jsons.from_json(path)["name"] == "dist"


def test_import_alias_serves_pisa_names_from_this_package():
    """``pisa_b200.compat.install_as_pisa``: reference-style imports (pipeline.py:284-296 resolves
    ``pisa.stages.<stage>.<service>``) land on the same module objects; missing subsystems fail loudly."""
    import importlib
    import sys
    import pisa_b200.compat as compat
    assert "pisa" not in sys.modules
    compat.install_as_pisa()
    try:
        compat.install_as_pisa()                                   # idempotent
        pipeline_mod = importlib.import_module("pisa.core.pipeline")
        prob3_mod = importlib.import_module("pisa.stages.osc.prob3")
        hist_mod = importlib.import_module("pisa.stages.utils.hist")
        import pisa_b200.core.pipeline, pisa_b200.stages.osc.prob3, pisa_b200.stages.utils.hist
        assert pipeline_mod is pisa_b200.core.pipeline and prob3_mod is pisa_b200.stages.osc.prob3
        assert hist_mod.hist is pisa_b200.stages.utils.hist.hist
        assert importlib.import_module("pisa").FTYPE is pisa_b200.FTYPE
        with pytest.raises(ModuleNotFoundError):
            importlib.import_module("pisa.analysis.analysis")      # outside the hot path: not silently replaced
    finally:
        compat.uninstall()
    assert not [m for m in sys.modules if m == "pisa" or m.startswith("pisa.")]


def test_fused_pipeline_shape_checks_without_a_gpu():
    """FusedPipeline only accepts osc.prob3 (events) -> [aeff.aeff] -> utils.hist (events; apply_unc_weights and
    unweighted are handled, tests/test_gpu_callers.py): anything else is refused before any device work (structure
    checks run on stage metadata only)."""
    from types import SimpleNamespace as NS
    from pisa_b200.fused import FusedPipeline

    def stage(stage_name, service_name, **kw):
        return NS(stage_name=stage_name, service_name=service_name, **kw)

    def pipe(*stages):
        return NS(stages=list(stages), run=lambda: (_ for _ in ()).throw(AssertionError("must not run")))

    osc = stage("osc", "prob3", calc_mode="events", apply_mode="events")
    hist_ok = dict(calc_mode="events", apply_unc_weights=False, unweighted=False, error_method="sumw2")
    cases = [
        pipe(stage("data", "csv_loader"), stage("utils", "hist", **hist_ok)),                      # no oscillation stage
        pipe(osc, osc, stage("utils", "hist", **hist_ok)),                                         # two of them
        pipe(osc, stage("aeff", "aeff")),                                                          # no histogram stage
        pipe(osc, stage("flux", "barr_simple"), stage("utils", "hist", **hist_ok)),                # something in between
        pipe(stage("osc", "prob3", calc_mode="grid", apply_mode="events"), stage("utils", "hist", **hist_ok)),
        pipe(osc, stage("utils", "hist", **dict(hist_ok, calc_mode="binned"))),
        pipe(osc, stage("aeff", "aeff"), stage("utils", "hist", **dict(hist_ok, error_method="fluctuate"))),
    ]
    for p in cases:
        with pytest.raises(NotImplementedError):
            FusedPipeline(p)
