"""GPU parity tests: bin indexing (bit-exact), weighted histogram, lookup, fused reweight+hist."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from conftest import ROOT, load_golden  # noqa: E402

DRAGON_E_EDGES = np.array([5.62341325, 7.49894209, 10.0, 13.33521432, 17.7827941, 23.71373706,
                           31.6227766, 42.16965034, 56.23413252])


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _sample(n, seed, dtype=np.float64):
    """SURVEY 8d reco variables."""
    rng = np.random.default_rng(seed)
    true_e = 10 ** rng.uniform(0, 3, n)
    true_cz = rng.uniform(-1, 1, n)
    reco_e = true_e * rng.lognormal(0, 0.3, n)
    reco_cz = true_cz + rng.normal(0, 0.2, n)
    pid = rng.integers(0, 2, n).astype(np.float64)
    return reco_e.astype(dtype), reco_cz.astype(dtype), pid.astype(dtype)


def _edge_values(edges):
    fin = edges[np.isfinite(edges)]
    return np.concatenate([edges, np.nextafter(fin, np.inf), np.nextafter(fin, -np.inf),
                           [-np.inf, np.inf, np.nan, fin.min() - 1, fin.max() + 1]])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_index_bit_exact_lin_log_edges(dtype):
    from pisa_b200 import ops
    dev = _dev()
    n = 300_000
    reco_e, reco_cz, pid = _sample(n, 0, dtype)
    # edge cases in every dimension (translation.py:868-905 spirit)
    ev = _edge_values(DRAGON_E_EDGES).astype(dtype)
    reco_e[:len(ev)] = ev
    cv = _edge_values(np.linspace(-1, 1, 9)).astype(dtype)
    reco_cz[100:100 + len(cv)] = cv
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    t = [torch.tensor(a, dtype=tdt, device=dev) for a in (reco_e, reco_cz, pid)]

    # (a) irregular energy edges (FP64 classification of dragon_datarelease)
    b, keep = ops.make_binning([dict(kind="edges", n_bins=8, edges=DRAGON_E_EDGES),
                                dict(kind="lin", n_bins=8, lo=-1.0, hi=1.0),
                                dict(kind="lin", n_bins=2, lo=0.0, hi=2.0)], dev)
    idx = ops.hist_index(b, t).cpu().numpy()
    ie = oracle.digitize_irregular(reco_e, DRAGON_E_EDGES, dtype)
    i2, _ = oracle.regular_index([reco_cz, pid], [-1.0, 0.0], [1.0, 2.0], [8, 2], dtype)
    ref = np.where((ie >= 0) & (ie < 8) & (i2 >= 0), ie * 16 + i2, -1)
    assert np.array_equal(idx, ref.astype(np.int32))

    # (b) log-regular energy axis: linear bins in log(x) (hist.py:114-121)
    with np.errstate(all="ignore"):
        log_e = np.log(reco_e)  # numpy log in FTYPE, like Container.translate
    # the binning's edges are FTYPE values, and the regularised domain is their FTYPE log (hist.py:118-120)
    e_lo, e_hi = dtype(5.62341325), dtype(56.23413252)
    b, keep = ops.make_binning([dict(kind="log", n_bins=8, lo=float(e_lo), hi=float(e_hi)),
                                dict(kind="lin", n_bins=8, lo=-1.0, hi=1.0),
                                dict(kind="lin", n_bins=2, lo=0.0, hi=2.0)], dev)
    idx = ops.hist_index(b, t).cpu().numpy()
    lo, hi = float(np.log(e_lo)), float(np.log(e_hi))
    ref, _ = oracle.regular_index([log_e, reco_cz, pid], [lo, -1.0, 0.0], [hi, 1.0, 2.0], [8, 8, 2], dtype)
    diff = np.flatnonzero(idx != ref)
    # device log vs numpy log can differ by 1 ulp: only events within 1 ulp of an edge may move
    assert len(diff) <= 2, (len(diff), reco_e[diff][:5])
    # samples sitting exactly on the ends of the logarithmic domain (clipped reconstructions pile up there): the
    # domain's log is taken by the same device function as the samples', so lo is inside and hi outside in both
    # precisions, as in the reference (np.log of the FTYPE edge and of the FTYPE sample)
    for lo_raw, hi_raw in ((5.62341325, 56.23413252), (1.0, 1000.0), (0.3, 7.7)):
        lo_t, hi_t = dtype(lo_raw), dtype(hi_raw)
        b1, _ = ops.make_binning([dict(kind="log", n_bins=8, lo=float(lo_t), hi=float(hi_t))], dev)
        ends = np.array([lo_t, hi_t, np.nextafter(lo_t, dtype(0)), np.nextafter(hi_t, dtype(0))], dtype=dtype)
        got = ops.hist_index(b1, [torch.tensor(ends, device=dev)]).cpu().numpy()
        # (a value one ulp below an end may share the end's logarithm: the in-range test is made in log space, as in
        # the reference, so it goes with the end or to the other side)
        assert got[0] == 0 and got[1] == -1 and got[2] in (-1, 0) and got[3] in (7, -1), (lo_raw, hi_raw, got)


def test_accumulate_vs_oracle_and_deterministic():
    from pisa_b200 import ops
    dev = _dev()
    n = 1_000_003
    rng = np.random.default_rng(2)
    idx = rng.integers(-1, 128, n).astype(np.int32)
    idx[rng.random(n) < 0.3] = 5  # a hot bin
    w = rng.uniform(0, 2, n)
    ti, tw = torch.tensor(idx, device=dev), torch.tensor(w, device=dev)
    h, h2 = ops.hist_accumulate(ti, tw, 128)
    ref = oracle.accumulate(idx, w, 128)
    ref2 = oracle.accumulate(idx, w * w, 128)
    assert np.allclose(h.cpu().numpy(), ref, rtol=1e-10, atol=0)
    assert np.allclose(h2.cpu().numpy(), ref2, rtol=1e-10, atol=0)
    # counts are exact
    c, _ = ops.hist_accumulate(ti, None, 128, want_w2=False)
    assert np.array_equal(c.cpu().numpy(), np.bincount(idx[idx >= 0], minlength=128).astype(np.float64))
    # run-to-run bit reproducibility (fixed accumulation order, no float atomics)
    for _ in range(3):
        hb, hb2 = ops.hist_accumulate(ti, tw, 128)
        assert torch.equal(hb, h) and torch.equal(hb2, h2)
    # ragged / tiny / empty inputs
    for m in (0, 1, 31, 33, 129):
        hh, _ = ops.hist_accumulate(ti[:m].contiguous(), tw[:m].contiguous(), 128)
        assert np.allclose(hh.cpu().numpy(), oracle.accumulate(idx[:m], w[:m], 128), rtol=1e-12, atol=0)
    # float32 weights, large binning (atomic path)
    w32 = w.astype(np.float32)
    idx_big = rng.integers(0, 3200, n).astype(np.int32)
    hb, _ = ops.hist_accumulate(torch.tensor(idx_big, device=dev), torch.tensor(w32, device=dev), 3200)
    assert np.allclose(hb.cpu().numpy(), oracle.accumulate(idx_big, w32.astype(np.float64), 3200), rtol=1e-10)


def test_lookup_exact():
    from pisa_b200 import ops
    dev = _dev()
    g = load_golden("ref_hist_f8.npz")
    x, y = g["lookup/x"], g["lookup/y"]
    x0, x1, nx, y0, y1, ny, _, _, _ = g["lookup/binning"]
    b, keep = ops.make_binning([dict(kind="lin", n_bins=int(nx), lo=x0, hi=x1),
                                dict(kind="lin", n_bins=int(ny), lo=y0, hi=y1)], dev)
    idx = ops.hist_index(b, [torch.tensor(x, device=dev), torch.tensor(y, device=dev)])
    ref_idx, _ = oracle.regular_index([x, y], [x0, y0], [x1, y1], [nx, ny])
    assert np.array_equal(idx.cpu().numpy(), ref_idx.astype(np.int32))
    for key in ("h2", "h2a"):
        out = ops.lookup(idx, torch.tensor(g["lookup/" + key], device=dev)).cpu().numpy()
        assert np.array_equal(out, oracle.lookup(ref_idx, g["lookup/" + key]))
    # and against the reference's own njit lookups, except the events it reads out of bounds
    def rounds_up(v, lo, hi, n):
        return (v >= lo) & (v < hi) & (((v - lo) * (n / (hi - lo))).astype(np.int64) >= n)
    ok = ~(rounds_up(x, x0, x1, nx) | rounds_up(y, y0, y1, ny))
    out = ops.lookup(idx, torch.tensor(g["lookup/h2"], device=dev)).cpu().numpy()
    assert np.array_equal(out[ok], g["lookup/o2"][ok])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_reweight_hist_vs_oracle_chain(dtype):
    """prob3 -> fill_probs -> weights *= flux.prob -> hist(w), hist(w^2): fused kernel vs the oracle."""
    from pisa_b200 import ops
    dev = _dev()
    g = load_golden("ref_prob3_f8.npz")
    key = "nufit20_nh_dcp306_stdnsi/nu"
    dm, mix, mat_pot = g[key + "/dm"], g[key + "/mix"], g[key + "/mat_pot"]
    consts = ops.OscConsts.from_matrices(dm, mix, mat_pot)
    L = oracle.OracleLayers(np.loadtxt(os.path.join(ROOT, "pisa_b200", "resources", "osc", "PREM_12layer.dat")), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = ops.Earth.from_arrays(L.radii, L.rhos, L.coszen_limit, L.r_detector, L.max_layers)
    n = 100_000
    rng = np.random.default_rng(1)
    energy = (10 ** rng.uniform(0, 3, n)).astype(dtype)
    coszen = rng.uniform(-1, 1, n).astype(dtype)
    flux = rng.uniform(0.5, 1.5, (n, 2)).astype(dtype)
    w0 = rng.uniform(0, 1, n).astype(dtype)
    idx = rng.integers(-1, 128, n).astype(np.int32)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    T = lambda a: torch.tensor(a, dtype=tdt, device=dev)  # noqa: E731
    zero = np.zeros((3, 3), dtype=np.complex128)
    _, den, dis = L.calcLayers(coszen.astype(np.float64))
    for nubar, flav in ((1, 1), (-1, 0), (1, 2)):
        prob = oracle.propagate_array(dm, mix, mat_pot, -1, zero, np.zeros((3, 3)), nubar,
                                      energy.astype(np.float64), den, dis, n_threads=os.cpu_count())
        pe, pmu = oracle.fill_probs(prob, 0, flav), oracle.fill_probs(prob, 1, flav)
        w = w0.astype(np.float64) * (flux[:, 0].astype(np.float64) * pe + flux[:, 1].astype(np.float64) * pmu)
        ref, ref2 = oracle.accumulate(idx, w, 128), oracle.accumulate(idx, w * w, 128)
        wout = torch.empty(n, dtype=tdt, device=dev)
        h, h2 = ops.reweight_hist(consts, earth, nubar, flav, T(energy), T(coszen), T(flux), T(w0),
                                  torch.tensor(idx, device=dev), 128, weights_out=wout)
        tol = 1e-10 if dtype == np.float64 else 2e-5
        assert np.allclose(h.cpu().numpy(), ref, rtol=tol), np.abs(h.cpu().numpy() / ref - 1).max()
        assert np.allclose(h2.cpu().numpy(), ref2, rtol=2 * tol)
        # grouped thread order: same histogram within rounding, per-event weights bit-identical,
        # and bit-reproducible run to run
        order = ops.layer_order(earth, T(coszen))
        wout_o = torch.empty(n, dtype=tdt, device=dev)
        ho, ho2 = ops.reweight_hist(consts, earth, nubar, flav, T(energy), T(coszen), T(flux), T(w0),
                                    torch.tensor(idx, device=dev), 128, weights_out=wout_o, order=order)
        assert torch.equal(wout_o, wout)
        assert np.allclose(ho.cpu().numpy(), ref, rtol=tol) and np.allclose(ho2.cpu().numpy(), ref2, rtol=2 * tol)
        ho_b, _ = ops.reweight_hist(consts, earth, nubar, flav, T(energy), T(coszen), T(flux), T(w0),
                                    torch.tensor(idx, device=dev), 128, order=order)
        assert torch.equal(ho, ho_b)
        assert np.allclose(wout.cpu().numpy(), w, rtol=tol, atol=1e-12 if dtype == np.float64 else 1e-5)
        # unfused path gives the same histogram (binned weights within 1e-10)
        _, pe_t, pmu_t = ops.propagate_earth(consts, earth, nubar, T(energy), T(coszen), flav=flav,
                                             want_probability=False)
        w_t = ops.apply_osc_weights(T(flux), pe_t, pmu_t, T(w0).clone())
        hu, hu2 = ops.hist_accumulate(torch.tensor(idx, device=dev), w_t, 128)
        assert np.allclose(hu.cpu().numpy(), ref, rtol=tol)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_fused_reweight_hist_with_neutrino_decay(dtype):
    """decay_flag = 1 through the template entry points (reweight_hist_decay_kernel): single-container and batched
    calls against the oracle chain with the eigvals branch, per-event outputs, the chi2 epilogue, run-to-run
    bit-reproducibility, and the standard kernels recovered at alpha3 = 0."""
    from pisa_b200 import ops
    from pisa_b200.engine import ReweightEngine
    dev = _dev()
    g = load_golden("ref_decay_f8.npz")
    key = "nufit20_nh_dcp306_stdnsi_a2e-4/nu"
    dm, mix, mat_pot, md = g[key + "/dm"], g[key + "/mix"], g[key + "/mat_pot"], g[key + "/mat_decay"]
    consts = ops.OscConsts.from_matrices(dm, mix, mat_pot, 1, md)
    L = oracle.OracleLayers(np.loadtxt(os.path.join(ROOT, "pisa_b200", "resources", "osc", "PREM_12layer.dat")), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = ops.Earth.from_arrays(L.radii, L.rhos, L.coszen_limit, L.r_detector, L.max_layers)
    n = 60_000
    rng = np.random.default_rng(5)
    energy = (10 ** rng.uniform(0, 3, n)).astype(dtype)
    coszen = rng.uniform(-1, 1, n).astype(dtype)
    flux = rng.uniform(0.5, 1.5, (n, 2)).astype(dtype)
    w0 = rng.uniform(0, 1, n).astype(dtype)
    idx = rng.integers(-1, 128, n).astype(np.int32)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    T = lambda a: torch.tensor(a, dtype=tdt, device=dev)  # noqa: E731
    _, den, dis = L.calcLayers(coszen.astype(np.float64))
    tol = 1e-10 if dtype == np.float64 else 2e-5
    refs = []
    for nubar, flav in ((1, 1), (-1, 0)):
        prob = oracle.propagate_array(dm, mix, mat_pot, 1, md, np.zeros((3, 3)), nubar, energy.astype(np.float64), den, dis,
                                      n_threads=os.cpu_count())
        pe, pmu = oracle.fill_probs(prob, 0, flav), oracle.fill_probs(prob, 1, flav)
        w = w0.astype(np.float64) * (flux[:, 0].astype(np.float64) * pe + flux[:, 1].astype(np.float64) * pmu)
        ref, ref2 = oracle.accumulate(idx, w, 128), oracle.accumulate(idx, w * w, 128)
        refs.append((ref, ref2))
        wout = torch.empty(n, dtype=tdt, device=dev)
        h, h2 = ops.reweight_hist(consts, earth, nubar, flav, T(energy), T(coszen), T(flux), T(w0),
                                  torch.tensor(idx, device=dev), 128, weights_out=wout)
        assert np.allclose(h.cpu().numpy(), ref, rtol=tol), np.abs(h.cpu().numpy() / ref - 1).max()
        assert np.allclose(h2.cpu().numpy(), ref2, rtol=2 * tol)
        assert np.allclose(wout.cpu().numpy(), w, rtol=tol, atol=1e-12 if dtype == np.float64 else 1e-5)
        hb, _ = ops.reweight_hist(consts, earth, nubar, flav, T(energy), T(coszen), T(flux), T(w0),
                                  torch.tensor(idx, device=dev), 128)
        assert torch.equal(h, hb)
        assert h.sum().item() < 0.95 * oracle.accumulate(idx, w0.astype(np.float64) * flux.astype(np.float64).sum(axis=1), 128).sum()
    # batched form through the engine (what FusedPipeline and the fit loop call), with and without sorting
    for sort in (True, False):
        eng = ReweightEngine(earth, 128, dtype, dev, sort_events=sort)
        for name, nubar, flav in (("numu_cc", 1, 1), ("nuebar_cc", -1, 0)):
            eng.add_container(name, nubar, flav, T(energy), T(coszen), T(flux), T(w0), torch.tensor(idx, device=dev))
        out = eng.evaluate(consts).cpu().numpy()
        for c, (ref, ref2) in enumerate(refs):
            assert np.allclose(out[c, 0], ref, rtol=tol) and np.allclose(out[c, 1], ref2, rtol=2 * tol)
        assert np.array_equal(out, eng.evaluate(consts).cpu().numpy())
        # alpha3 = 0 through the decay kernel == the standard template kernel
        zero = ops.OscConsts.from_matrices(dm, mix, mat_pot, 1, np.zeros((3, 3), dtype=complex))
        std = ops.OscConsts.from_matrices(dm, mix, mat_pot)
        a, b = eng.evaluate(zero).cpu().numpy(), eng.evaluate(std).cpu().numpy()
        assert np.allclose(a, b, rtol=1e-10 if dtype == np.float64 else 1e-5)
        # the multi-template scan: decay tables travel per template; a template without decay in the same scan goes
        # through the general-matrix kernel too
        dec2 = ops.OscConsts.from_matrices(dm, mix, mat_pot, 1, 3.0 * md)
        many = eng.evaluate_many([consts, std, dec2]).cpu().numpy()
        assert np.allclose(many[0], out, rtol=1e-12 if dtype == np.float64 else 1e-6)
        assert np.allclose(many[1], b, rtol=1e-10 if dtype == np.float64 else 1e-5)
        assert np.allclose(many[2], eng.evaluate(dec2).cpu().numpy(), rtol=1e-12 if dtype == np.float64 else 1e-6)
        assert many[2][:, 0].sum() < many[0][:, 0].sum() < many[1][:, 0].sum()
        # the fit-loop form: template + container sum + mod_chi2 in the library call
        observed = torch.tensor(refs[0][0] + refs[1][0], device=dev)
        hist, chi2 = eng.evaluate_chi2(consts, observed)
        assert np.allclose(hist.cpu().numpy(), out, rtol=1e-12)
        assert float(chi2) < (1e-12 if dtype == np.float64 else 1e-3)


def test_mod_chi2():
    from pisa_b200 import ops
    dev = _dev()
    rng = np.random.default_rng(0)
    exp = rng.uniform(0, 50, 128)
    exp[3] = 0.0
    w2 = rng.uniform(0, 5, 128)
    obs = rng.poisson(exp).astype(np.float64)
    e = np.clip(exp, 1e-10, np.inf)
    ref = ((obs - e) ** 2 / (w2 + e)).sum()  # stats.py:674-695 with sigma^2 = sumw2
    out = ops.mod_chi2(torch.tensor(exp, device=dev), torch.tensor(w2, device=dev), torch.tensor(obs, device=dev))
    assert np.isclose(float(out), ref, rtol=1e-12)


def test_batched_template_equals_per_container_calls():
    """One launch over ragged containers (incl. an empty one, per-container scale, standard-matter AND NSI
    instantiations) == the per-container calls bit for bit, and == the oracle chain within 1e-10."""
    from pisa_b200 import ops
    from pisa_b200.utils import synthetic as syn
    dev = _dev()
    L = oracle.OracleLayers(np.loadtxt(os.path.join(ROOT, "pisa_b200", "resources", "osc", "PREM_12layer.dat")), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = ops.Earth.from_arrays(L.radii, L.rhos, L.coszen_limit, L.r_detector, L.max_layers)
    binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
    sizes = [5000, 1, 0, 777, 33, 12345, 64, 31, 2048, 9, 300, 4097]
    zero = np.zeros((3, 3), dtype=np.complex128)
    for nsi in (None, syn.STD_NSI):
        dm, mix, mat_pot = syn.osc_matrices(nsi=nsi)
        consts = ops.OscConsts.from_matrices(dm, mix, mat_pot)
        desc, host = [], []
        for c, ((name, nubar, flav), n) in enumerate(zip(syn.CONTAINERS, sizes)):
            ev = syn.make_events_numpy(n, seed=50 + c)
            t = {k: torch.tensor(v, device=dev) for k, v in ev.items()}
            index = ops.hist_index(binning, [t["reco_energy"], t["reco_coszen"], t["pid"]]) if n else \
                torch.empty(0, dtype=torch.int32, device=dev)
            scale = 0.5 + 0.25 * c
            desc.append(dict(nubar=nubar, flav=flav, energy=t["true_energy"], coszen=t["true_coszen"],
                             nu_flux=t["nu_flux"], weights=t["weights"], index=index, scale=scale,
                             weights_out=torch.empty(n, dtype=torch.float64, device=dev)))
            host.append((nubar, flav, ev, index.cpu().numpy(), scale))
        batch = ops.TemplateBatch(desc, 128)
        out = ops.reweight_hist_batch(consts, earth, batch)
        out2 = ops.reweight_hist_batch(consts, earth, batch)
        assert torch.equal(out, out2)                      # run-to-run bit-reproducible
        out = out.cpu().numpy()
        for c, (nubar, flav, ev, idx, scale) in enumerate(host):
            n = len(idx)
            if n == 0:
                assert not out[c].any()
                continue
            d = desc[c]
            h, h2 = ops.reweight_hist(consts, earth, nubar, flav, d["energy"], d["coszen"], d["nu_flux"],
                                      d["weights"], d["index"], 128)
            # the single-container entry point has scale == 1; compare through the oracle instead
            _, den, dis = L.calcLayers(ev["true_coszen"])
            prob = oracle.propagate_array(dm, mix, mat_pot, -1, zero, np.zeros((3, 3)), nubar, ev["true_energy"],
                                          den, dis)
            w = ev["weights"] * (ev["nu_flux"][:, 0] * prob[:, 0, flav] + ev["nu_flux"][:, 1] * prob[:, 1, flav]) * scale
            ref, ref2 = oracle.accumulate(idx, w, 128), oracle.accumulate(idx, w * w, 128)
            assert np.allclose(out[c, 0], ref, rtol=1e-10, atol=1e-300), (c, np.abs(out[c, 0] - ref).max())
            assert np.allclose(out[c, 1], ref2, rtol=2e-10, atol=1e-300)
            assert np.allclose(h.cpu().numpy() * scale, out[c, 0], rtol=1e-13)
            assert np.allclose(d["weights_out"].cpu().numpy(), w, rtol=1e-10, atol=1e-13)
    with pytest.raises(ValueError):
        ops.TemplateBatch(desc + desc, 128)                # more than MAX_BATCH containers


def test_full_size_properties_of_the_fused_path():
    """BASELINE-size launch (2e7 events in one container, the per-container size of config C4 at 8 GPUs is
    1.3e7): properties that do not need the oracle -- checksum of checksums, exact linearity under a
    power-of-two weight scale, bit-reproducibility, invariance under event permutation (rounding only)."""
    from pisa_b200 import ops
    from pisa_b200.utils import synthetic as syn
    dev = _dev()
    L = oracle.OracleLayers(np.loadtxt(os.path.join(ROOT, "pisa_b200", "resources", "osc", "PREM_12layer.dat")), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = ops.Earth.from_arrays(L.radii, L.rhos, L.coszen_limit, L.r_detector, L.max_layers)
    dm, mix, mat_pot = syn.osc_matrices(nsi=syn.STD_NSI)
    consts = ops.OscConsts.from_matrices(dm, mix, mat_pot)
    binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
    n = 20_000_000
    ev = syn.make_events_torch(n, seed=5, dtype=np.float64, device=dev)
    # push some events out of range so that index == -1 occurs at scale
    ev["reco_coszen"][::97] = 1.5
    idx = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
    assert int((idx < 0).sum()) >= n // 97
    wout = torch.empty(n, dtype=torch.float64, device=dev)
    args = (consts, earth, -1, 1, ev["true_energy"], ev["true_coszen"], ev["nu_flux"])
    h, h2 = ops.reweight_hist(*args, ev["weights"], idx, 128, weights_out=wout)
    inside = idx >= 0
    # checksum of checksums: every in-range event landed in exactly one bin
    assert abs(float(h.sum()) / float(wout[inside].sum()) - 1) < 1e-12
    assert abs(float(h2.sum()) / float((wout[inside] ** 2).sum()) - 1) < 1e-12
    assert float(wout.min()) >= 0.0 and torch.isfinite(wout).all()
    # per-bin cross-check against an independent device reduction (torch index_add, plumbing only)
    ref = torch.zeros(128, dtype=torch.float64, device=dev).index_add_(0, idx[inside].long(), wout[inside])
    assert torch.allclose(h, ref, rtol=1e-11, atol=0)
    # bit-reproducible, exactly linear under a power-of-two scale
    hb, _ = ops.reweight_hist(*args, ev["weights"], idx, 128)
    assert torch.equal(h, hb)
    h4, h42 = ops.reweight_hist(*args, (ev["weights"] * 4.0).contiguous(), idx, 128)
    assert torch.equal(h4, 4.0 * h) and torch.equal(h42, 16.0 * h2)
    # permutation of the events: same histogram up to summation order
    perm = torch.randperm(n, device=dev)
    hp, _ = ops.reweight_hist(consts, earth, -1, 1, ev["true_energy"][perm].contiguous(),
                              ev["true_coszen"][perm].contiguous(), ev["nu_flux"][perm].contiguous(),
                              ev["weights"][perm].contiguous(), idx[perm].contiguous(), 128)
    assert torch.allclose(hp, h, rtol=1e-11, atol=0)
    # unitarity of the full matrix on 1e7 of the events
    p, _, _ = ops.propagate_earth(consts, earth, -1, ev["true_energy"][:10_000_000].contiguous(),
                                  ev["true_coszen"][:10_000_000].contiguous())
    assert float((p.sum(dim=1) - 1).abs().max()) < 5e-12 and float((p.sum(dim=2) - 1).abs().max()) < 5e-12


def test_full_size_properties_of_the_decay_path():
    """The decay template kernel at a BASELINE-size launch (1.2e7 events per container): checksum of checksums, exact
    linearity under a power-of-two weight scale, bit-reproducibility, permutation invariance, monotonic loss of
    probability in alpha3, and agreement of the template's per-event weights with the stand-alone decay kernel."""
    from pisa_b200 import ops
    from pisa_b200.utils import synthetic as syn
    dev = _dev()
    L = oracle.OracleLayers(np.loadtxt(os.path.join(ROOT, "pisa_b200", "resources", "osc", "PREM_12layer.dat")), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = ops.Earth.from_arrays(L.radii, L.rhos, L.coszen_limit, L.r_detector, L.max_layers)
    dm, mix, mat_pot = syn.osc_matrices(nsi=syn.STD_NSI)
    md = np.zeros((3, 3), dtype=complex)
    md[2, 2] = -2.0e-4j
    consts = ops.OscConsts.from_matrices(dm, mix, mat_pot, 1, md)
    binning, keep = ops.make_binning(syn.DRAGON_DIMS, dev)
    n = 12_000_000
    ev = syn.make_events_torch(n, seed=6, dtype=np.float64, device=dev)
    ev["reco_coszen"][::89] = 1.5
    idx = ops.hist_index(binning, [ev["reco_energy"], ev["reco_coszen"], ev["pid"]])
    inside = idx >= 0
    wout = torch.empty(n, dtype=torch.float64, device=dev)
    args = (consts, earth, 1, 1, ev["true_energy"], ev["true_coszen"], ev["nu_flux"])
    h, h2 = ops.reweight_hist(*args, ev["weights"], idx, 128, weights_out=wout)
    assert abs(float(h.sum()) / float(wout[inside].sum()) - 1) < 1e-12
    assert abs(float(h2.sum()) / float((wout[inside] ** 2).sum()) - 1) < 1e-12
    assert float(wout.min()) >= 0.0 and torch.isfinite(wout).all()
    ref = torch.zeros(128, dtype=torch.float64, device=dev).index_add_(0, idx[inside].long(), wout[inside])
    assert torch.allclose(h, ref, rtol=1e-11, atol=0)
    hb, _ = ops.reweight_hist(*args, ev["weights"], idx, 128)
    assert torch.equal(h, hb)
    h4, h42 = ops.reweight_hist(*args, (ev["weights"] * 4.0).contiguous(), idx, 128)
    assert torch.equal(h4, 4.0 * h) and torch.equal(h42, 16.0 * h2)
    perm = torch.randperm(n, device=dev)
    hp, _ = ops.reweight_hist(consts, earth, 1, 1, ev["true_energy"][perm].contiguous(),
                              ev["true_coszen"][perm].contiguous(), ev["nu_flux"][perm].contiguous(),
                              ev["weights"][perm].contiguous(), idx[perm].contiguous(), 128)
    assert torch.allclose(hp, h, rtol=1e-11, atol=0)
    # the stand-alone decay kernel gives the template's per-event weights
    _, pe, pmu = ops.propagate_earth(consts, earth, 1, ev["true_energy"], ev["true_coszen"], flav=1, want_probability=False)
    w_ref = ev["weights"] * (ev["nu_flux"][:, 0] * pe + ev["nu_flux"][:, 1] * pmu)
    assert torch.allclose(wout, w_ref, rtol=1e-12, atol=1e-15)
    # more decay, fewer events -- in every bin (numu appearance + survival both lose the nu3 component)
    totals = []
    for alpha in (0.0, 1.0e-4, 2.0e-4, 8.0e-4):
        md[2, 2] = -1j * alpha
        c = ops.OscConsts.from_matrices(dm, mix, mat_pot, 1, md)
        totals.append(float(ops.reweight_hist(c, earth, 1, 1, ev["true_energy"], ev["true_coszen"], ev["nu_flux"],
                                              ev["weights"], idx, 128)[0].sum()))
    assert totals[0] > totals[1] > totals[2] > totals[3]
    # no probability is created: rows and columns of the full matrix sum to at most one
    p, _, _ = ops.propagate_earth(consts, earth, 1, ev["true_energy"][:4_000_000].contiguous(),
                                  ev["true_coszen"][:4_000_000].contiguous())
    assert float(p.sum(dim=1).max()) < 1 + 1e-12 and float(p.sum(dim=2).max()) < 1 + 1e-12 and float(p.min()) >= 0.0


@pytest.mark.parametrize("dtype", [np.float64, np.float32, None])
def test_large_input_histogram(dtype):
    """3e6 events through the stand-alone histogram kernel: ragged tail, invalid and out-of-range indices,
    all storage types.  Checked against the oracle (exact for counts) and for bit-reproducibility."""
    from pisa_b200 import ops
    dev = _dev()
    n = 3_000_000 + 777
    rng = np.random.default_rng(3)
    idx = rng.integers(-1, 128, n).astype(np.int32)
    idx[::1001] = 500                       # beyond the binning: must be ignored, not written
    w = None if dtype is None else rng.uniform(0, 2, n).astype(dtype)
    ti = torch.tensor(idx, device=dev)
    tw = None if w is None else torch.tensor(w, device=dev)
    h, h2 = ops.hist_accumulate(ti, tw, 128)
    hb, h2b = ops.hist_accumulate(ti, tw, 128)
    assert torch.equal(h, hb) and torch.equal(h2, h2b)
    ok = (idx >= 0) & (idx < 128)
    wd = np.ones(n) if w is None else w.astype(np.float64)
    ref = oracle.accumulate(np.where(ok, idx, -1), wd, 128)
    ref2 = oracle.accumulate(np.where(ok, idx, -1), wd * wd, 128)
    if w is None:
        assert np.array_equal(h.cpu().numpy(), ref) and np.array_equal(h2.cpu().numpy(), ref2)
    else:
        assert np.allclose(h.cpu().numpy(), ref, rtol=1e-11, atol=0)
        assert np.allclose(h2.cpu().numpy(), ref2, rtol=1e-11, atol=0)
    # an unaligned view (offset by one element)
    h_u, _ = ops.hist_accumulate(ti[1:], None if tw is None else tw[1:], 128)
    ref_u = oracle.accumulate(np.where(ok[1:], idx[1:], -1), wd[1:], 128)
    assert np.allclose(h_u.cpu().numpy(), ref_u, rtol=1e-11, atol=0)


@pytest.mark.parametrize("n_bins", [1, 400, 1024])
def test_fused_kernel_other_bin_counts(n_bins):
    """Fused template kernel with 1, 400 and 1024 bins (one block per SM at the large counts: 225 KB of shared
    memory) against the unfused path; more than PISAB_DET_MAX_BINS bins must be refused, not mis-launched."""
    from pisa_b200 import ops
    from pisa_b200._lib import PisabError
    from pisa_b200.utils import synthetic as syn
    dev = _dev()
    L = oracle.OracleLayers(np.loadtxt(os.path.join(ROOT, "pisa_b200", "resources", "osc", "PREM_12layer.dat")), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = ops.Earth.from_arrays(L.radii, L.rhos, L.coszen_limit, L.r_detector, L.max_layers)
    dm, mix, mat_pot = syn.osc_matrices()
    consts = ops.OscConsts.from_matrices(dm, mix, mat_pot)
    n = 200_000
    ev = syn.make_events_torch(n, seed=8, dtype=np.float64, device=dev)
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    idx = torch.randint(-1, n_bins, (n,), generator=g, device=dev, dtype=torch.int32)
    h, h2 = ops.reweight_hist(consts, earth, 1, 0, ev["true_energy"], ev["true_coszen"], ev["nu_flux"], ev["weights"],
                              idx, n_bins)
    _, pe, pmu = ops.propagate_earth(consts, earth, 1, ev["true_energy"], ev["true_coszen"], flav=0,
                                     want_probability=False)
    w = ops.apply_osc_weights(ev["nu_flux"], pe, pmu, ev["weights"].clone())
    hu, hu2 = ops.hist_accumulate(idx, w, n_bins)
    assert torch.allclose(h, hu, rtol=1e-12, atol=0) and torch.allclose(h2, hu2, rtol=1e-12, atol=0)
    assert abs(float(h.sum()) / float(w[idx >= 0].sum()) - 1) < 1e-12
    if n_bins == 1024:
        big = torch.randint(-1, 1025, (n,), generator=g, device=dev, dtype=torch.int32)
        with pytest.raises(PisabError):
            ops.reweight_hist(consts, earth, 1, 0, ev["true_energy"], ev["true_coszen"], ev["nu_flux"],
                              ev["weights"], big, 1025)


@pytest.mark.parametrize("n,n_bins", [(0, 128), (1, 1), (2048, 128), (2049, 7), (4095, 129), (300_001, 128),
                                      (1_000_003, 256), (5000, 257), (400_001, 400), (300_007, 513), (700_001, 1024)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_planned_histogram_vs_oracle(n, n_bins, dtype):
    """pisab_hist_plan_build + pisab_hist_accumulate_planned (bin-sorted tiles, one thread per bin): empty, single,
    tile-boundary and ragged sizes, out-of-range indices dropped, a hot bin, both storage types, one / two / four
    bins per thread (up to 1024 bins); identical bits run to run; the plan is refused for a different index length;
    above 1024 bins ``hist_plan`` returns the sorted plan instead."""
    from pisa_b200 import ops
    dev = _dev()
    rng = np.random.default_rng(n + n_bins)
    idx = rng.integers(-2, n_bins + 2, n).astype(np.int32)
    if n > 10:
        idx[rng.random(n) < 0.2] = n_bins // 2
    w = rng.uniform(-0.5, 2, n).astype(dtype)
    ti, tw = torch.tensor(idx, device=dev), torch.tensor(w, device=dev)
    plan = ops.hist_plan(ti, n_bins)
    assert plan is not None
    h, h2 = ops.hist_accumulate(ti, tw, n_bins, plan=plan)
    ok = (idx >= 0) & (idx < n_bins)
    wd = w.astype(np.float64)
    ref = oracle.accumulate(np.where(ok, idx, -1), wd, n_bins)
    ref2 = oracle.accumulate(np.where(ok, idx, -1), wd * wd, n_bins)
    scale = max(1.0, float(np.abs(wd).sum()))
    assert np.allclose(h.cpu().numpy(), ref, rtol=1e-11, atol=1e-15 * scale)
    assert np.allclose(h2.cpu().numpy(), ref2, rtol=1e-11, atol=1e-15 * scale)
    for _ in range(2):
        hb, hb2 = ops.hist_accumulate(ti, tw, n_bins, plan=plan)
        assert torch.equal(hb, h) and torch.equal(hb2, h2)
    h_only, none = ops.hist_accumulate(ti, tw, n_bins, want_w2=False, plan=plan)
    assert none is None and torch.equal(h_only, h)
    cnt, cnt2 = ops.hist_accumulate(ti, None, n_bins, plan=plan)      # static counts, cached on the plan
    want = np.bincount(idx[ok], minlength=n_bins).astype(np.float64)
    assert np.array_equal(cnt.cpu().numpy(), want) and np.array_equal(cnt2.cpu().numpy(), want)
    if n > 1:
        with pytest.raises(ValueError):
            ops.hist_accumulate(ti[1:].contiguous(), tw[1:].contiguous(), n_bins, plan=plan)
    big = ops.hist_plan(ti, 1025)
    assert big is not None and big.perm is not None and big.buf is None


def test_large_binning_is_exact_and_order_independent():
    """More than PISAB_DET_MAX_BINS bins: 128-bit fixed-point accumulation with integer atomics.  The result is the
    exact sum rounded once, so it is bit-identical run to run AND for any permutation of the events (float atomics,
    round 1, differed in the last bits from run to run), and it agrees with math.fsum."""
    import math
    from pisa_b200 import ops
    dev = _dev()
    rng = np.random.default_rng(12)
    n, n_bins = 600_000, 3200
    idx = rng.integers(-1, n_bins + 1, n).astype(np.int32)
    w = (rng.uniform(0, 1, n) * 10.0 ** rng.uniform(-6, 3, n)) * rng.choice([1.0, 1.0, 1.0, -1.0], n)
    ti, tw = torch.tensor(idx, device=dev), torch.tensor(w, device=dev)
    h, h2 = ops.hist_accumulate(ti, tw, n_bins)
    perm = rng.permutation(n)
    hp, hp2 = ops.hist_accumulate(torch.tensor(idx[perm], device=dev), torch.tensor(w[perm], device=dev), n_bins)
    assert torch.equal(h, hp) and torch.equal(h2, hp2)
    for _ in range(2):
        hb, hb2 = ops.hist_accumulate(ti, tw, n_bins)
        assert torch.equal(hb, h) and torch.equal(hb2, h2)
    hh, hh2 = h.cpu().numpy(), h2.cpu().numpy()
    for b in (0, 17, 1599, 3199):
        sel = idx == b
        assert hh[b] == math.fsum(w[sel]), b                       # exactly rounded
        assert abs(hh2[b] - math.fsum(w[sel] ** 2)) <= 2e-16 * hh2[b]   # (w*w is rounded before it is accumulated)
    ok = (idx >= 0) & (idx < n_bins)
    assert np.allclose(hh, oracle.accumulate(np.where(ok, idx, -1), w, n_bins), rtol=1e-9, atol=1e-12 * np.abs(w).sum())
    # counts and float32 weights
    c, _ = ops.hist_accumulate(ti, None, n_bins, want_w2=False)
    assert np.array_equal(c.cpu().numpy(), np.bincount(idx[ok], minlength=n_bins).astype(np.float64))
    w32 = np.abs(w).astype(np.float32)
    h32, _ = ops.hist_accumulate(ti, torch.tensor(w32, device=dev), n_bins)
    assert h32.cpu().numpy()[5] == math.fsum(w32[idx == 5].astype(np.float64))
    # the sorted plan (static indices, fit loop): the same exact sums from one atomic pair per warp and plane
    plan = ops.hist_plan(ti, n_bins)
    assert plan is not None and plan.perm is not None
    sk = plan.sorted_index.cpu().numpy()
    assert np.all(np.diff(sk) >= 0) and np.array_equal(np.where(ok, idx, n_bins)[plan.perm.cpu().numpy()], sk)
    for weights in (tw, torch.tensor(w32, device=dev)):
        a, a2 = ops.hist_accumulate(ti, weights, n_bins)
        b_, b2 = ops.hist_accumulate(ti, weights, n_bins, plan=plan)
        assert torch.equal(a, b_) and torch.equal(a2, b2)
    cp, _ = ops.hist_accumulate(ti, None, n_bins, want_w2=False, plan=plan)
    assert torch.equal(cp, c)
    # all-zero weights and an empty input
    z, z2 = ops.hist_accumulate(ti, torch.zeros(n, dtype=torch.float64, device=dev), n_bins)
    assert float(z.abs().sum()) == 0.0 and float(z2.abs().sum()) == 0.0
    e, _ = ops.hist_accumulate(ti[:0].contiguous(), tw[:0].contiguous(), n_bins)
    assert float(e.abs().sum()) == 0.0


def test_sort_order_is_stable():
    """pisab_sort_order_i32 (setup-time radix sort behind layer_order and the sorted histogram plan) against a stable
    numpy argsort: ascending, descending, limited key bits, with the sorted keys."""
    from pisa_b200 import ops
    dev = _dev()
    rng = np.random.default_rng(5)
    for n, hi, bits in ((0, 10, 0), (1, 10, 0), (1000, 7, 3), (300_001, 60, 8), (200_000, 2_000_000, 0)):
        k = rng.integers(0, hi, n).astype(np.int32)
        t = torch.tensor(k, device=dev)
        for desc in (False, True):
            order, sk = ops.sort_order(t, descending=desc, key_bits=bits, want_sorted=True)
            want = np.argsort(-k.astype(np.int64) if desc else k, kind="stable")
            assert np.array_equal(order.cpu().numpy(), want), (n, hi, desc)
            assert np.array_equal(sk.cpu().numpy(), k[want])
    with pytest.raises(TypeError):
        ops.sort_order(torch.zeros(4, dtype=torch.int64, device=dev))


def test_fused_template_with_3200_bins_is_deterministic():
    """The fused template kernel above PISAB_DET_MAX_BINS (the 40 x 40 x 2 stress binning of SURVEY 8d): same launch,
    exact fixed-point accumulators; equal to propagate + reweight + histogram, bit-identical run to run."""
    from pisa_b200 import ops
    from pisa_b200.engine import ReweightEngine
    from pisa_b200.utils import synthetic as syn
    dev = _dev()
    L = oracle.OracleLayers(np.loadtxt(os.path.join(ROOT, "pisa_b200", "resources", "osc", "PREM_12layer.dat")), 2.0, 20.0)
    L.setElecFrac(0.4656, 0.4656, 0.4957)
    earth = ops.Earth.from_arrays(L.radii, L.rhos, L.coszen_limit, L.r_detector, L.max_layers)
    n_bins, n = 3200, 150_000
    g = torch.Generator(device=dev)
    g.manual_seed(9)
    for dtype, nsi, tol in ((np.float64, False, 1e-12), (np.float64, True, 1e-12), (np.float32, False, 2e-5)):
        dm, mix, mat_pot = syn.osc_matrices(nsi=syn.STD_NSI if nsi else None)
        consts = ops.OscConsts.from_matrices(dm, mix, mat_pot)
        eng = ReweightEngine(earth, n_bins, dtype, dev)
        parts = []
        for c, (name, nubar, flav) in enumerate(syn.CONTAINERS[:3] + syn.CONTAINERS[7:9]):
            ev = syn.make_events_torch(n + 1000 * c, seed=30 + c, dtype=dtype, device=dev)
            idx = torch.randint(-1, n_bins, (n + 1000 * c,), generator=g, device=dev, dtype=torch.int32)
            eng.add_container(name, nubar, flav, ev["true_energy"], ev["true_coszen"], ev["nu_flux"], ev["weights"], idx)
            parts.append((nubar, flav, ev, idx))
        eng.set_scales([1.0, 2.5, 0.5, 1.0, 3.0])
        out = eng.evaluate(consts).clone()
        for _ in range(2):
            assert torch.equal(eng.evaluate(consts), out)
        for c, ((nubar, flav, ev, idx), scale) in enumerate(zip(parts, [1.0, 2.5, 0.5, 1.0, 3.0])):
            _, pe, pmu = ops.propagate_earth(consts, earth, nubar, ev["true_energy"], ev["true_coszen"], flav=flav,
                                             want_probability=False)
            w = (ev["weights"].double() * (ev["nu_flux"][:, 0].double() * pe.double() + ev["nu_flux"][:, 1].double() * pmu.double())
                 * scale).contiguous()
            hu, hu2 = ops.hist_accumulate(idx, w, n_bins)
            nz = hu != 0
            assert float(((out[c, 0] - hu).abs()[nz] / hu.abs()[nz]).max()) < tol, (dtype, nsi, c)
            assert float(((out[c, 1] - hu2).abs()[nz] / hu2.abs()[nz]).max()) < 2 * tol, (dtype, nsi, c)
