"""GPU parity (through the C ABI) of RK3 Wilson flow, E(t) and the stout forward layer against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DIMS = (4, 4, 4, 4)
DIMS_ANISO = (4, 6, 4, 8)


def test_flow_golden_plaquette(backend, oracle):
    """Reference golden value: 4^4 SU(3), StableRNG(123) hot start, 100 x eps=0.01 RK3 steps ->
    plaquette 0.8786515255315753 (test/gradientflow_test.jl:129-139; the reference asserts 10 %)."""
    import gfb200

    U = gfb200.gauge_configuration(DIMS, backend=backend).upload(oracle.hot_start_stable123(DIMS))
    g = gfb200.gradient_flow(U, steps=1, step_size=0.01)
    for _ in range(100):
        gfb200.flow_(U, g)
    plaq = gfb200.measure_plaquette(U)
    assert abs(plaq - 0.8786515255315753) / 0.8786515255315753 < 1e-1  # the reference's own bar
    assert abs(plaq - 0.8786515255315753) < 1e-11  # ours


def test_flow_energy_matches_oracle(backend, oracle):
    """E(t) within 1e-11 of the oracle along the flow (BASELINE.json), clover and plaquette definitions."""
    import gfb200

    dims = DIMS_ANISO
    Uh = oracle.hot_start_philox(dims, 77)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    g = gfb200.gradient_flow(U, steps=5, step_size=0.01)
    for it in range(6):
        e_gpu = gfb200.energy_density(U, "clover")
        e_cpu = oracle.energy_density_clover(Uh, dims)
        assert abs(e_gpu - e_cpu) < 1e-11 * max(1.0, abs(e_cpu))
        p_gpu = gfb200.energy_density(U, "plaquette")
        p_cpu = 2.0 * (18.0 - oracle.plaquette_sum(Uh, dims) / np.prod(dims))
        assert abs(p_gpu - p_cpu) < 1e-11 * max(1.0, abs(p_cpu))
        gfb200.flow_(U, g)
        for _ in range(5):
            oracle.flow_step(Uh, dims, 0.01)
    assert np.abs(U.to_host() - Uh).max() < 1e-12


def test_flow_force_and_exp_primitives(backend, oracle):
    """add_force!(plaqonly) and exp_aF_U! used by the unfused flow (AbstractGaugefields.jl:2717-2762, 2810-2841)."""
    import ctypes

    import gfb200

    dims = DIMS_ANISO
    Uh = oracle.hot_start_philox(dims, 5)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    F = gfb200.gauge_momenta(U)
    backend.call("gfb_flow_force", F._h, U._h)
    Fw = oracle.flow_force(Uh, dims)
    assert np.abs(F.to_host() - Fw).max() / np.abs(Fw).max() < 1e-12
    W = gfb200.GaugeConfiguration(backend, dims)
    backend.call("gfb_exp_aF_U", W._h, ctypes.c_double(-0.0025), F._h, U._h)
    Ww = oracle.update_links(Uh, Fw, dims, -0.0025)
    assert np.abs(W.to_host() - Ww).max() < 5e-15
    with pytest.raises(ValueError):
        backend.call("gfb_exp_aF_U", W._h, ctypes.c_double(0.0), F._h, U._h)


def test_stout_forward_matches_oracle(backend, oracle):
    """STOUT_Layer forward!, < 1e-11 like the reference's cross-backend check (test/latticematrices_compat.jl:391-465)."""
    import gfb200

    dims = DIMS_ANISO
    Uh = oracle.hot_start_philox(dims, 99)
    for _ in range(2):
        oracle.flow_step(Uh, dims, 0.02)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    sm = gfb200.stout_smearing(U, rho=0.1, layers=2)
    out = gfb200.smear(U, sm)
    want = oracle.stout_forward(oracle.stout_forward(Uh, dims, 0.1), dims, 0.1)
    assert np.abs(out.to_host() - want).max() < 1e-11
    assert np.abs(out.to_host() - want).max() < 1e-13
    assert np.array_equal(U.to_host(), Uh)  # input untouched
    # tape: Q coefficients
    Q = gfb200.gauge_momenta(U)
    tmp = gfb200.GaugeConfiguration(backend, dims)
    backend.call("gfb_stout_forward", tmp._h, U._h, 0.1, Q._h)
    _, Qw = oracle.stout_forward(Uh, dims, 0.1, want_q=True)
    assert np.abs(Q.to_host() - Qw).max() < 1e-13
    assert gfb200.measure_plaquette(out) > gfb200.measure_plaquette(U)


def test_polyakov_loop(backend, oracle):
    import gfb200

    dims = DIMS_ANISO
    Uh = oracle.hot_start_philox(dims, 3)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    got = gfb200.measure_polyakov_loop(U, normalize=False)
    want = oracle.polyakov(Uh, dims)
    assert abs(got - want) < 1e-13


def test_stout_backward_matches_oracle(backend, oracle):
    """back_prop through two stout layers (Abstractsmearing.jl:352-411, stout_fast.jl:317-407) and the resulting
    HMC force of the smeared Wilson action, against the oracle's chain rule (itself pinned by finite differences
    and by the reference's closed-form exp pull-back in tests/test_oracle_pins.py)."""
    import gfb200

    dims = DIMS_ANISO
    beta, rho = 5.7, 0.1
    Uh = oracle.hot_start_philox(dims, 2024)
    for _ in range(2):
        oracle.flow_step(Uh, dims, 0.02)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    loops = gfb200.make_loops_fromname("plaquette")
    action = gfb200.GaugeAction(U).push(beta / 2, loops + loops.adjoint())
    sm = gfb200.stout_smearing(U, rho=rho, layers=2)
    Uout, multi = gfb200.calc_smearedU(U, sm)
    U1 = oracle.stout_forward(Uh, dims, rho)
    U2 = oracle.stout_forward(U1, dims, rho)
    d2 = oracle.wilson_dSdU(U2, dims, beta)
    dS = gfb200.calc_dSdU(action, Uout)
    assert np.abs(dS.to_host() - d2).max() < 1e-12 * np.abs(d2).max()
    bare = gfb200.back_prop(dS, sm, multi, U)
    d1 = oracle.stout_backward(d2, U1, dims, rho)
    d0 = oracle.stout_backward(d1, Uh, dims, rho)
    assert np.abs(bare.to_host() - d0).max() < 1e-11 * np.abs(d0).max()
    assert np.abs(bare.to_host() - d0).max() < 1e-12 * np.abs(d0).max()
    # one layer from an arbitrary (non-derivative) input field exercises every term with generic matrices
    rng = np.random.default_rng(4)
    dr = rng.normal(size=Uh.shape) + 1j * rng.normal(size=Uh.shape)
    D = gfb200.GaugeConfiguration(backend, dims).upload(dr)
    out = gfb200.GaugeConfiguration(backend, dims)
    backend.call("gfb_stout_backward", out._h, D._h, U._h, 0.12)
    want = oracle.stout_backward(dr, Uh, dims, 0.12)
    assert np.abs(out.to_host() - want).max() < 1e-12 * np.abs(want).max()
    # the composed kick (test/HMCstout_test_nowing.jl:99-118)
    Ph = oracle.gaussian_momenta(dims, 5, 0)
    P = gfb200.gauge_momenta(U).upload(Ph)
    gfb200.stout_force_(P, U, action, sm, 0.05)
    want_p = oracle.kick_from_dSdU(Ph.copy(), Uh, d0, dims, -0.05 / 3.0)
    assert np.abs(P.to_host() - want_p).max() < 1e-12 * np.abs(want_p).max()
    with pytest.raises(ValueError):
        backend.call("gfb_stout_backward", out._h, out._h, U._h, 0.1)
