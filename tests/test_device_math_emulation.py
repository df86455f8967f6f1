"""The per-event DEVICE math (pisa_b200/csrc/prob3_device.cuh: eigenvalues, Cayley-Hamilton transition matrices,
shell-twin propagation, vacuum shortcut, standard-matter specialisation) compiled for the HOST with g++ through a small
shim (tests/hostemu/) and compared with the oracle on seeded events.  This is a CPU-side guard for numerics changes in
the CUDA headers -- it runs where there is no GPU; the GPU parity tests (tests/test_gpu_prob3.py) remain the parity
tests proper.  Nothing here is part of the product: the library is only ever built by nvcc without PISAB_HOST_EMU."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

HOSTEMU = os.path.join(ROOT, "tests", "hostemu")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("hostemu") / "libemu.so")
    cmd = [gxx, "-O2", "-fopenmp", "-shared", "-fPIC", "-std=c++17", "-DPISAB_HOST_EMU", "-include",
           os.path.join(HOSTEMU, "cuda_shim.h"), "-I" + os.path.join(ROOT, "pisa_b200", "csrc"),
           "-I" + os.path.join(ROOT, "include"), "-o", out, os.path.join(HOSTEMU, "emu.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-3000:]
    sys.path.insert(0, HOSTEMU)
    import check
    check.load(out)
    return check


@pytest.mark.parametrize("nsi,nubar", [(False, 1), (True, 1), (False, -1), (True, -1)])
def test_device_math_on_host_matches_oracle(emu, nsi, nubar):
    assert emu.run(n=30000, nsi=nsi, nubar=nubar, seed=5) < 1e-10     # measured ~4e-13


def test_device_math_on_host_lri_and_other_earth_models(emu):
    assert emu.run(n=20000, lri=np.diag([1e-14, -1e-14, 0.0]), seed=6) < 1e-10
    assert emu.run(n=20000, model="PREM_10layer.dat", seed=7) < 1e-10
    assert emu.run(n=20000, model="PREM_4layer.dat", depth=10.0, seed=8) < 1e-10


def test_device_math_on_host_prem59_maximum_depth(emu):
    """PREM_59layer: 61 shells, layer arrays 122 wide with up to 118 active slots -- the largest Earth model the reference
    ships and the limit of its 120-slot layer cache (numba_osc_kernels.py:173-177,227)."""
    assert emu.run(n=6000, model="PREM_59layer.dat", seed=9) < 1e-10
    assert emu.run(n=4000, nsi=True, nubar=-1, model="PREM_59layer.dat", seed=10) < 1e-10
    assert emu.run_decay(n=3000, model="PREM_59layer.dat", seed=11) < 1e-10


@pytest.mark.parametrize("nsi,nubar", [(False, 1), (True, -1)])
def test_fp32_mode_math_on_host_within_1e5_of_fp64_oracle(emu, nsi, nubar):
    """The mixed-precision FP32 mode (prob3_mp.cuh: FP64 eigenvalues / phases, float matrices and state) against the
    FP64 oracle on float32-rounded inputs, energies 1 GeV .. 1 TeV: BASELINE tolerance 1e-5 absolute.  Measured on
    1e6 events: max 6.5e-6, 99.9 % below 2.6e-6, mean 3e-7 (the floor of float matrices times ~50 rad phases)."""
    assert emu.run_mp(n=60000, nsi=nsi, nubar=nubar, seed=11, verbose=False) < 1e-5


def test_fp32_mode_math_on_host_lri_and_small_earth(emu):
    assert emu.run_mp(n=20000, lri=np.diag([1e-14, -1e-14, 0.0]), seed=12, verbose=False) < 1e-5
    assert emu.run_mp(n=20000, model="PREM_4layer.dat", depth=10.0, seed=13, verbose=False) < 1e-5


@pytest.mark.parametrize("nsi,nubar", [(False, 1), (True, -1)])
def test_fp32_mode_pairs_are_bit_identical_to_single_events(emu, nsi, nubar):
    """Two events per thread (float part in the two lanes of the packed FP32 instructions, emulated lane-wise on the
    host) give bit for bit the probabilities of the one-event form; a pair whose events cross different shells is
    flagged."""
    identical, flagged_ok, n_unequal = emu.run_pairs(n=40000, nsi=nsi, nubar=nubar, seed=21)
    assert identical and flagged_ok and n_unequal > 0


@pytest.mark.parametrize("nsi,nubar", [(False, 1), (True, -1)])
def test_decay_math_on_host_matches_oracle(emu, nsi, nubar):
    """The decay branch (prob3_decay.cuh: complex Cardano + Newton eigenvalues of the non-Hermitian layer, complex
    Cayley-Hamilton coefficients, damped amplitudes) through the Earth walk against the oracle's restatement of the
    reference's numpy.linalg.eigvals branch.  Measured 2e-13."""
    assert emu.run_decay(n=20000, nsi=nsi, nubar=nubar, seed=31) < 1e-11


def test_decay_math_on_host_edge_parameters(emu):
    assert emu.run_decay(n=10000, alpha3=0.0, seed=32) < 1e-11                    # Hermitian input to the general solver
    assert emu.run_decay(n=10000, alpha3=5e-4, lri=np.diag([1e-14, -1e-14, 0.0]), seed=33) < 1e-11
    assert emu.run_decay(n=10000, alpha3=1e-3, e_max=4.0, seed=34) < 1e-10         # up to 10 TeV (measured 1.8e-12)
    general = [[0, 0, 0], [0, -2e-5j, 1e-5 - 1e-5j], [0, 1e-5 + 1e-5j, -1e-4j]]    # a full complex decay matrix
    assert emu.run_decay(n=10000, mat_decay=general, nubar=-1, seed=35) < 1e-11
    assert emu.run_decay(n=10000, model="PREM_4layer.dat", depth=10.0, seed=36) < 1e-11
