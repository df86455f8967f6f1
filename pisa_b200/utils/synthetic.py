"""Synthetic Monte-Carlo events of the shapes named in BASELINE.json / SURVEY.md 8d.

The bundled IceCube MC file is absent from the reference tree (.MISSING_LARGE_BLOBS), so every
configuration runs on synthetic events with the same columns as
settings/pipeline/IceCube_3y_neutrinos.cfg:39-46 and the laws of
pisa/stages/data/toy_event_generator.py:75-76:
    true_energy = 10**U(0,3) GeV, true_coszen = U(-1,1), nu_flux[N,2] = U(0.5,1.5),
    weighted_aeff = U(0,1), reco_energy = true_energy*logN(0,0.3) clipped into the binning range,
    reco_coszen = clip(true_coszen + N(0,0.2), -1, nextafter(1,0)), pid in {0,1}.
"""
import numpy as np

# (name, nubar, flav) in the order of settings/pipeline/osc_example.cfg:41
CONTAINERS = [
    ("nue_cc", 1, 0), ("numu_cc", 1, 1), ("nutau_cc", 1, 2),
    ("nue_nc", 1, 0), ("numu_nc", 1, 1), ("nutau_nc", 1, 2),
    ("nuebar_cc", -1, 0), ("numubar_cc", -1, 1), ("nutaubar_cc", -1, 2),
    ("nuebar_nc", -1, 0), ("numubar_nc", -1, 1), ("nutaubar_nc", -1, 2),
]

# settings/binning/IceCube_3y_oscillations.cfg:11-19 `dragon_datarelease`
DRAGON_E_EDGES = np.array([5.62341325, 7.49894209, 10.0, 13.33521432, 17.7827941, 23.71373706,
                           31.6227766, 42.16965034, 56.23413252])
DRAGON_DIMS = [
    dict(name="reco_energy", kind="edges", n_bins=8, edges=DRAGON_E_EDGES),
    dict(name="reco_coszen", kind="lin", n_bins=8, lo=-1.0, hi=1.0),
    dict(name="pid", kind="lin", n_bins=2, lo=-0.5, hi=1.5),
]
DRAGON_NBINS = 128

# settings/osc/nufitv20.cfg (NH) + settings/osc/earth.cfg
NUFIT20_NH = dict(theta12=33.48, theta13=8.5, theta23=42.3, deltacp=0.0, deltam21=7.5e-5, deltam31=2.457e-3)
EARTH = dict(earth_model="osc/PREM_12layer.dat", YeI=0.4656, YeO=0.4656, YeM=0.4957, detector_depth=2.0,
             prop_height=20.0)
# standard-NSI point the reference test intended (numba_osc_tests.py:130-135)
STD_NSI = dict(eps_ee=0.0, eps_emu=(0.07, 340.0), eps_etau=(0.06, 35.0), eps_mumu=0.0, eps_mutau=(0.003, 175.0),
               eps_tautau=0.0)


def osc_matrices(params=NUFIT20_NH, nsi=None, include_nlo=False):
    """(dm_matrix, mix_matrix, mat_pot) like prob3.compute_function builds them (prob3.py:485-559)."""
    from pisa_b200.stages.osc.nsi_params import StdNSIParams
    from pisa_b200.stages.osc.osc_params import OscParams
    op = OscParams()
    op.theta12, op.theta13, op.theta23 = (np.deg2rad(params[k]) for k in ("theta12", "theta13", "theta23"))
    op.deltacp = np.deg2rad(params["deltacp"])
    op.dm21, op.dm31 = params["deltam21"], params["deltam31"]
    mat_pot = np.zeros((3, 3), dtype=np.complex128)
    mat_pot[0, 0] += 1.020 if include_nlo else 1.0
    if nsi is not None:
        p = StdNSIParams()
        p.eps_ee, p.eps_mumu, p.eps_tautau = nsi["eps_ee"], nsi["eps_mumu"], nsi["eps_tautau"]
        for k in ("eps_emu", "eps_etau", "eps_mutau"):
            setattr(p, k, (nsi[k][0], np.deg2rad(nsi[k][1])))
        mat_pot = mat_pot + p.eps_matrix
    return op.dm_matrix, op.mix_matrix_complex, mat_pot


def make_events_numpy(n, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    true_energy = 10 ** rng.uniform(0, 3, n)
    true_coszen = rng.uniform(-1, 1, n)
    nu_flux = rng.uniform(0.5, 1.5, (n, 2))
    weights = rng.uniform(0, 1, n)
    lo, hi = DRAGON_E_EDGES[0], DRAGON_E_EDGES[-1]
    reco_energy = np.clip(true_energy * rng.lognormal(0, 0.3, n), lo, np.nextafter(hi, 0))
    reco_coszen = np.clip(true_coszen + rng.normal(0, 0.2, n), -1, np.nextafter(1.0, 0))
    pid = rng.integers(0, 2, n).astype(np.float64)
    out = dict(true_energy=true_energy, true_coszen=true_coszen, nu_flux=nu_flux, weights=weights,
               reco_energy=reco_energy, reco_coszen=reco_coszen, pid=pid)
    out = {k: np.ascontiguousarray(v.astype(dtype)) for k, v in out.items()}
    _clip_in_storage_type(out, np.minimum, np.dtype(dtype).type)
    return out


def _clip_in_storage_type(ev, minimum, ftype):
    """Rounding to float32 can move a clipped value onto the (excluded) upper edge: clip once more in the storage
    type, so that the reconstructed coordinates stay inside the half-open binning range in both FTYPE modes."""
    if ftype == np.float64:
        return
    # (on the logarithmic energy axis the in-range test is made on float32 logs, which cannot resolve one ulp of
    # the edge itself: stay 2^-20 below it)
    top_e = float(np.float32(DRAGON_E_EDGES[-1]) * np.float32(1.0 - 2.0 ** -20))
    top_cz = float(np.nextafter(np.float32(1.0), np.float32(0)))
    ev["reco_energy"] = minimum(ev["reco_energy"], ev["reco_energy"] * 0 + top_e)
    ev["reco_coszen"] = minimum(ev["reco_coszen"], ev["reco_coszen"] * 0 + top_cz)
    # ... and the LOWER energy edge rounds DOWN in float32 (5.62341309 < 5.62341325): every event clipped onto it
    # would fall out of the binning in FP32 mode only.  Clip to the smallest float32 not below the edge.
    lo32 = np.float32(DRAGON_E_EDGES[0])
    if float(lo32) < float(DRAGON_E_EDGES[0]):
        lo32 = np.nextafter(lo32, np.float32(np.inf))
    ev["reco_energy"] = -minimum(-ev["reco_energy"], ev["reco_energy"] * 0 - float(lo32))


def make_events_torch(n, seed, dtype, device):
    """Same laws, generated on the device (for the 1e8-event bench workload)."""
    import torch
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    u = lambda *shape: torch.rand(*shape, generator=g, device=device, dtype=torch.float64)  # noqa: E731
    nrm = lambda *shape: torch.randn(*shape, generator=g, device=device, dtype=torch.float64)  # noqa: E731
    true_energy = torch.pow(10.0, 3.0 * u(n))
    true_coszen = 2.0 * u(n) - 1.0
    nu_flux = 0.5 + u(n, 2)
    weights = u(n)
    lo, hi = float(DRAGON_E_EDGES[0]), float(np.nextafter(DRAGON_E_EDGES[-1], 0))
    reco_energy = torch.clamp(true_energy * torch.exp(0.3 * nrm(n)), lo, hi)
    reco_coszen = torch.clamp(true_coszen + 0.2 * nrm(n), -1.0, float(np.nextafter(1.0, 0)))
    pid = torch.randint(0, 2, (n,), generator=g, device=device).to(torch.float64)
    out = dict(true_energy=true_energy, true_coszen=true_coszen, nu_flux=nu_flux, weights=weights,
               reco_energy=reco_energy, reco_coszen=reco_coszen, pid=pid)
    out = {k: v.to(tdt).contiguous() for k, v in out.items()}
    _clip_in_storage_type(out, torch.minimum, np.dtype(dtype).type)
    return out


def layer_counts(coszen_limit, coszen, idx_first_inner=2):
    """(L_active, L_distinct, L_cached) per event under the reference's layer cache rule
    (numba_osc_kernels.py:236-241) for the Earth geometry of layers.py:94-159; used for the
    algorithmic-FLOP count of SURVEY.md 8d.  Works on numpy arrays or torch tensors."""
    lim = np.asarray(coszen_limit, dtype=np.float64)
    try:
        import torch
        is_t = isinstance(coszen, torch.Tensor)
    except ImportError:
        is_t = False
    if is_t:
        import torch
        limt = torch.as_tensor(lim, device=coszen.device, dtype=torch.float64)
        k = (limt[None, :] > coszen.to(torch.float64)[:, None]).sum(dim=1)
        up = coszen.to(torch.float64) < float(lim[idx_first_inner])
        active = torch.where(up, 2 * k - 2, torch.full_like(k, idx_first_inner))
        distinct = torch.where(up, k + 1, torch.full_like(k, idx_first_inner))
        return active, distinct, active - distinct
    cz = np.asarray(coszen, dtype=np.float64)
    k = (lim[None, :] > cz[:, None]).sum(axis=1)
    up = cz < lim[idx_first_inner]
    active = np.where(up, 2 * k - 2, idx_first_inner)
    distinct = np.where(up, k + 1, idx_first_inner)
    return active, distinct, active - distinct


def flops_per_event(l_distinct_mean, l_cached_mean):
    """F_event = 1329 + 2367*L_distinct + 226*L_cached (reference arithmetic)."""
    return 1329.0 + 2367.0 * float(l_distinct_mean) + 226.0 * float(l_cached_mean)
