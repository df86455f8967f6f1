"""Service smoke test in the style of the reference's pisa_tests/test_services.py (:136-183): every service is built by
its module's ``init_test()``, given two 10-event containers with the reference's test inputs (``linspace(0.1, 1, 10)``
for per-event keys, a random ``(10, 2)`` flux, ``nubar`` / ``flav`` as aux scalars, the 3 x 3 x 3 test binning) and run
through ``setup(); run()``.  As in the reference this only checks that the services run; numbers are pinned by the
parity tests."""
import importlib

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SERVICES = ["flux.barr_simple", "flux.honda_ip", "osc.prob3", "utils.hist"]


def _inputs(service):
    from pisa_b200 import FTYPE
    from pisa_b200.core.binning import MultiDimBinning, OneDimBinning
    from pisa_b200.core.container import Container, ContainerSet
    from pisa_b200.utils.units import ureg
    binning = MultiDimBinning([OneDimBinning(name="reco_energy", is_log=True, num_bins=3, domain=[0.1, 1] * ureg.GeV),
                               OneDimBinning(name="reco_coszen", is_lin=True, num_bins=3, domain=[0.1, 1]),
                               OneDimBinning(name="pid", is_lin=True, num_bins=3, domain=[0.1, 1])])
    rng = np.random.RandomState(0)
    containers = []
    for name in ("test1_cc", "test2_nc"):
        c = Container(name)
        for k in sorted(set(list(service.expected_container_keys) + ["reco_energy", "reco_coszen", "pid", "weights"])):
            if k in ("nubar", "flav"):
                c.set_aux_data(k, 1)
            elif k in ("nu_flux", "nu_flux_nominal", "nubar_flux_nominal"):
                c[k] = rng.random_sample((10, 2)).astype(FTYPE)
            else:
                c[k] = np.linspace(0.1, 1, 10, dtype=FTYPE)
        containers.append(c)
    data = ContainerSet("data", containers)
    data["output_binning"] = binning
    data["regularized_output_binning"] = binning
    return data


@pytest.mark.parametrize("name", SERVICES)
def test_service_sets_up_and_runs(name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    module = importlib.import_module("pisa_b200.stages." + name)
    service = module.init_test()
    assert (service.stage_name, service.service_name) == tuple(name.split("."))
    service.data = _inputs(service)
    for mode in ("calc_mode", "apply_mode"):          # first supported representation, "events" where possible
        if getattr(service, mode) is None and None not in service.supported_reps[mode]:
            setattr(service, mode, "events")
    service.setup()
    service.run()
    for c in service.data:
        for key in ("nu_flux", "nu_flux_nominal", "prob_e", "prob_mu", "weights"):
            if key in c.all_keys and c.find_valid_representation(key) is not None:
                c.representation = c.find_valid_representation(key)
                assert bool(torch.isfinite(c[key]).all()), (name, c.name, key)
