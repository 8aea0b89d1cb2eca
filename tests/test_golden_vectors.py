"""Committed known-answer vectors (tests/golden/wilson_4x4x4x4.npz and wilson_8x4x2x4.npz, made by tests/golden/make_golden.py;
the first lattice runs in k_force_fused, the second in the t-marching kernel).

CPU leg: the oracle still reproduces them (guards the checker against drift) and they contain the reference's own golden
values.  GPU leg: the CUDA path, through the C ABI, reproduces them within the tolerances of BASELINE.json."""
import os

import numpy as np
import pytest

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BETA = 5.7


@pytest.fixture(scope="module", params=[(4, 4, 4, 4), (8, 4, 2, 4)], ids=["4x4x4x4", "8x4x2x4"])
def gold(request):
    g = dict(np.load(os.path.join(GOLD_DIR, "wilson_%s.npz" % "x".join(map(str, request.param)))))
    g["dims"] = request.param
    return g


def test_oracle_reproduces_golden(oracle, gold):
    DIMS = gold["dims"]
    U, P = gold["U0"].copy(), gold["P0"].copy()
    assert np.array_equal(oracle.hot_start_philox(DIMS, 1234), U)
    assert np.array_equal(oracle.gaussian_momenta(DIMS, 0x5678, 0), P)
    assert abs(oracle.plaquette_sum(U, DIMS) - gold["plaquette_sum"]) < 1e-12
    assert np.abs(oracle.force(U, DIMS, BETA) - gold["force"]).max() < 1e-13
    H0, H1 = oracle.md_trajectory(U, P, DIMS, BETA, 20, 1.0, 0)
    assert abs(H0 - gold["H0_qpq"]) < 1e-9 and abs((H1 - H0) - gold["dH_qpq"]) < 1e-9
    if "ref_hot_plaquette" in gold:  # the reference's golden values (test/init.jl:276-283, test/gradientflow_test.jl:129-139)
        assert abs(gold["oracle_hot_plaquette"] - gold["ref_hot_plaquette"]) < 1e-8 * gold["ref_hot_plaquette"]
        assert abs(gold["oracle_flow_plaquette"] - gold["ref_flow_plaquette"]) < 1e-11


@pytest.mark.gpu
def test_cuda_reproduces_golden(backend, gold):
    import gfb200

    DIMS = gold["dims"]
    U = gfb200.gauge_configuration(DIMS, backend=backend).upload(gold["U0"])
    P = gfb200.gauge_momenta(U).upload(gold["P0"])
    # the device RNG regenerates the stored start fields
    assert np.abs(gfb200.gauge_configuration(DIMS, backend=backend, start="hot", seed=1234).to_host() - gold["U0"]).max() < 1e-14
    assert np.abs(gfb200.gaussian_momenta(U, seed=0x5678, sweep=0).to_host() - gold["P0"]).max() < 1e-13
    ps = float(gold["plaquette_sum"])
    assert abs(gfb200.calculate_Plaquette(U) - ps) <= 1e-12 * abs(ps) + 1e-13
    loops = gfb200.make_loops_fromname("plaquette")
    action = gfb200.GaugeAction(U).push(BETA / 2, loops + loops.adjoint())
    F = gfb200.gauge_momenta(U)
    gfb200.md_force_(F, action, U)
    assert np.abs(F.to_host() - gold["force"]).max() < 1e-12 * np.abs(gold["force"]).max()
    assert abs(P * P - gold["kinetic"]) < 1e-12 * gold["kinetic"]
    assert abs(gfb200.energy_density(U) - gold["energy_clover"]) < 1e-11
    for name, integ in (("qpq", gfb200.QPQ), ("pqp", gfb200.PQP)):
        for fused in (False, True):
            U.upload(gold["U0"]); P.upload(gold["P0"])
            md = gfb200.md_driver(U, action, steps=20, trajectory_length=1.0, integrator=integ, fused=fused)
            res = gfb200.md_trajectory_(U, P, md)
            assert abs(res.initial_hamiltonian - gold["H0_" + name]) < 1e-12 * abs(gold["H0_" + name])
            assert abs(res.delta_hamiltonian - gold["dH_" + name]) < 1e-9
            if name == "qpq":
                assert np.abs(U.to_host() - gold["U_after_qpq"]).max() < 1e-11
            else:
                assert np.abs(U.to_host()[0, 0] - gold["U_after_pqp_mu0_t0"]).max() < 1e-11
    # flow: E(t) series within 1e-11
    U.upload(gold["U0"])
    g = gfb200.gradient_flow(U, steps=1, step_size=0.01)
    for k in range(10):
        gfb200.flow_(U, g)
        assert abs(gfb200.energy_density(U, "clover") - gold["flow_E_clover"][k]) < 1e-11
        assert abs(gfb200.energy_density(U, "plaquette") - gold["flow_E_plaquette"][k]) < 1e-11
    assert np.abs(U.to_host()[3, 3] - gold["U_flowed_mu3_t3"]).max() < 1e-12
    # stout forward (2 layers) and the smeared-action force through back_prop
    U.upload(gold["U0"])
    sm = gfb200.stout_smearing(U, rho=0.1, layers=2)
    Uout, multi = gfb200.calc_smearedU(U, sm)
    assert np.abs(Uout.to_host() - gold["U_stout2"]).max() < 1e-12
    P.clear_()
    gfb200.stout_force_(P, U, action, sm, -1.0)  # P = +1/3 TA(U dSdU) * ... sign folded: compare to -force
    assert np.abs(-P.to_host() - gold["stout_force"]).max() < 1e-12 * np.abs(gold["stout_force"]).max()
