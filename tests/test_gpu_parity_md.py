"""GPU parity tests (through the C ABI) of the HMC hot path against the CPU oracle.

Tolerances are the ones BASELINE.json states: plaquette and per-link force 1e-12 relative,
Delta H 1e-9 per trajectory, identical accept/reject over 20 trajectories.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DIMS = (4, 4, 4, 4)
DIMS_ANISO = (6, 4, 8, 4)  # distinct extents catch any x/y/z/t mix-up
REL = 1e-12


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def make(backend, oracle, dims, seed=1234):
    import gfb200

    Uh = oracle.hot_start_philox(dims, seed)
    U = gfb200.gauge_configuration(dims, backend=backend, start="cold")
    U.upload(Uh)
    return U, Uh


def test_upload_download_roundtrip(backend, oracle):
    U, Uh = make(backend, oracle, DIMS_ANISO)
    assert np.array_equal(U.to_host(), Uh)
    import gfb200

    P = gfb200.gauge_momenta(U)
    Ph = oracle.gaussian_momenta(DIMS_ANISO, 7, 3)
    P.upload(Ph)
    assert np.array_equal(P.to_host(), Ph)


def test_cold_start_invariants(backend, oracle):
    """cold start => plaquette 1, p*p 0, Delta H 0 (test/md_driver.jl:371-395, test/init.jl:190-211)."""
    import gfb200

    U = gfb200.gauge_configuration(DIMS, backend=backend, start="cold")
    assert gfb200.measure_plaquette(U) == 1.0
    p = gfb200.gauge_momenta(U)
    assert p * p == 0.0
    action = gfb200.GaugeAction(U).push(5.7 / 2, gfb200.make_loops_fromname("plaquette") + gfb200.make_loops_fromname("plaquette").adjoint())
    for fused in (False, True):
        md = gfb200.md_driver(U, action, steps=5, fused=fused)
        res = gfb200.md_trajectory_(U, p, md)
        assert res.delta_hamiltonian == 0.0
        assert gfb200.measure_plaquette(U) == 1.0
        assert p * p == 0.0


@pytest.mark.parametrize("dims", [DIMS, DIMS_ANISO])
def test_plaquette_matches_oracle(backend, oracle, dims):
    import gfb200

    U, Uh = make(backend, oracle, dims)
    got = gfb200.calculate_Plaquette(U)
    want = oracle.plaquette_sum(Uh, dims)
    assert abs(got - want) <= REL * abs(want) + 1e-12 * np.prod(dims) * 1e-3


def test_golden_hot_start_plaquette(backend, oracle):
    """Reference golden value 0.008449494077606137 (test/init.jl:276-283) through the CUDA plaquette kernel."""
    import gfb200

    Uh = oracle.hot_start_stable123(DIMS)
    U = gfb200.gauge_configuration(DIMS, backend=backend).upload(Uh)
    assert abs(gfb200.measure_plaquette(U) - 0.008449494077606137) / 0.008449494077606137 < 1e-8
    assert abs(gfb200.measure_plaquette(U) - 0.008449494077606137) < 1e-14


@pytest.mark.parametrize("dims", [DIMS, DIMS_ANISO])
def test_force_matches_oracle(backend, oracle, dims):
    """md_force! per link, 1e-12 relative (BASELINE.json)."""
    import gfb200

    U, Uh = make(backend, oracle, dims)
    F = gfb200.gauge_momenta(U)
    action = gfb200.GaugeAction(U).push(6.0 / 2, gfb200.make_loops_fromname("plaquette") + gfb200.make_loops_fromname("plaquette").adjoint())
    gfb200.md_force_(F, action, U)
    want = oracle.force(Uh, dims, 6.0)
    assert relerr(F.to_host(), want) < REL


def test_momentum_kick_and_link_update(backend, oracle):
    import gfb200

    dims = DIMS_ANISO
    U, Uh = make(backend, oracle, dims)
    Ph = oracle.gaussian_momenta(dims, 0x5678, 0)
    P = gfb200.gauge_momenta(U).upload(Ph)
    action = gfb200.GaugeAction(U).push(5.7 / 2, gfb200.make_loops_fromname("plaquette") + gfb200.make_loops_fromname("plaquette").adjoint())
    md = gfb200.md_driver(U, action, steps=10)
    gfb200.update_momenta_(P, U, 0.05, md)
    want = oracle.update_momenta(Ph.copy(), Uh, dims, 0.05, 5.7)
    assert relerr(P.to_host(), want) < REL
    gfb200.update_gaugefields_(U, P, 0.05, md)
    wantU = oracle.update_links(Uh, want, dims, 0.05)
    assert np.abs(U.to_host() - wantU).max() < 5e-15
    # kinetic energy and Hamiltonian
    assert abs(P * P - oracle.momentum_norm2(want, dims)) <= REL * oracle.momentum_norm2(want, dims)
    H = gfb200.md_hamiltonian(U, P, md)
    Hw = oracle.hamiltonian(wantU, want, dims, 5.7)
    assert abs(H - Hw) <= 1e-12 * abs(Hw)


def test_exp_large_and_small_arguments(backend, oracle):
    """exp(t P) over 12 orders of magnitude of |t P|, including the scaling-and-squaring branch."""
    import gfb200

    dims = (4, 4, 4, 4)
    U = gfb200.gauge_configuration(dims, backend=backend, start="cold")
    Uh = oracle.set_cold(dims)
    rng = np.random.default_rng(5)
    Ph = rng.normal(size=oracle.p_shape(dims))
    scale = 10.0 ** rng.uniform(-10, 1.3, size=Ph.shape[:-1])
    Ph *= scale[..., None]
    Ph[0, 0, 0, 0, 0] = 0.0  # exact zero -> exact identity
    P = gfb200.gauge_momenta(U).upload(Ph)
    gfb200.update_gaugefields_(U, P, 1.0)
    want = oracle.update_links(Uh, Ph, dims, 1.0)
    got = U.to_host()
    assert np.abs(got - want).max() < 2e-13
    assert np.array_equal(got[0, 0, 0, 0, 0], np.eye(3))
    small = scale < 0.5
    assert np.abs(got - want)[small].max() < 4e-15


@pytest.mark.parametrize("integrator", ["QPQ", "PQP"])
@pytest.mark.parametrize("fused", [False, True])
def test_md_trajectory_delta_h(backend, oracle, integrator, fused):
    """Delta H within 1e-9 of the oracle's trajectory on the same seeded start (BASELINE.json)."""
    import gfb200

    dims = DIMS
    U, Uh = make(backend, oracle, dims)
    Ph = oracle.gaussian_momenta(dims, 0x5678, 0)
    P = gfb200.gauge_momenta(U).upload(Ph)
    action = gfb200.GaugeAction(U).push(5.7 / 2, gfb200.make_loops_fromname("plaquette") + gfb200.make_loops_fromname("plaquette").adjoint())
    integ = getattr(gfb200, integrator)
    md = gfb200.md_driver(U, action, steps=20, trajectory_length=1.0, integrator=integ, fused=fused)
    res = gfb200.md_trajectory_(U, P, md)
    H0, H1 = oracle.md_trajectory(Uh, Ph, dims, 5.7, 20, 1.0, integ.code)
    assert abs(res.initial_hamiltonian - H0) <= 1e-12 * abs(H0)
    assert abs(res.delta_hamiltonian - (H1 - H0)) < 1e-9
    assert np.abs(U.to_host() - Uh).max() < 1e-11
    assert np.abs(P.to_host() - Ph).max() < 1e-10


def test_md_reversibility(backend, oracle):
    """forward then backward trajectory restores U and P to < 2e-12 (test/md_driver.jl:417-482)."""
    import gfb200

    dims = DIMS
    U, Uh = make(backend, oracle, dims)
    Ph = oracle.gaussian_momenta(dims, 11, 0)
    P = gfb200.gauge_momenta(U).upload(Ph)
    action = gfb200.GaugeAction(U).push(5.7 / 2, gfb200.make_loops_fromname("plaquette") + gfb200.make_loops_fromname("plaquette").adjoint())
    for integ in (gfb200.QPQ, gfb200.PQP):
        for fused in (False, True):
            fwd = gfb200.md_driver(U, action, steps=8, trajectory_length=0.4, integrator=integ, fused=fused)
            bwd = gfb200.md_driver(U, action, steps=8, trajectory_length=-0.4, integrator=integ, fused=fused)
            gfb200.md_trajectory_(U, P, fwd, diagnostics=False)
            gfb200.md_trajectory_(U, P, bwd, diagnostics=False)
            assert np.abs(U.to_host() - Uh).max() < 2e-12
            assert np.abs(P.to_host() - Ph).max() < 2e-12


@pytest.mark.parametrize("dims", [DIMS, (8, 4, 4, 4)])  # 4^4: k_force_fused; 8x4x4x4: the t-marching kernel
def test_hmc_accept_reject_sequence(backend, oracle, dims):
    """20 trajectories of the docs/src/hmc.md:128-190 loop: identical accept/reject sequence, Delta H within 1e-9
    when each trajectory starts from the oracle's state (re-synchronised), and the free-running chain's sequence."""
    import gfb200

    beta, steps = 5.7, 20
    Uh = oracle.hot_start_philox(dims, 0x1234)
    # a few thermalisation flow steps put Delta H in the O(1) regime so accept and reject both occur
    for _ in range(3):
        oracle.flow_step(Uh, dims, 0.02)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    action = gfb200.GaugeAction(U).push(beta / 2, gfb200.make_loops_fromname("plaquette") + gfb200.make_loops_fromname("plaquette").adjoint())
    md = gfb200.md_driver(U, action, steps=steps, trajectory_length=1.0, integrator=gfb200.QPQ, fused=True)
    P = gfb200.gauge_momenta(U)
    rng_gpu = np.random.default_rng(0x9ABC)
    rng_cpu = np.random.default_rng(0x9ABC)
    old = gfb200.copy_configuration(U)
    seq_gpu, seq_cpu, margins = [], [], []
    for traj in range(20):
        U.upload(Uh)  # every trajectory starts from the oracle's state (MD is chaotic)
        gfb200.gaussian_momenta_(P, seed=0x5678, sweep=traj)
        Ph = oracle.gaussian_momenta(dims, 0x5678, traj)
        assert np.abs(P.to_host() - Ph).max() < 1e-13
        gfb200.copy_configuration_(old, U)
        res = gfb200.md_trajectory_(U, P, md)
        Uo = Uh.copy()
        H0, H1 = oracle.md_trajectory(Uo, Ph, dims, beta, steps, 1.0, 0)
        assert abs(res.delta_hamiltonian - (H1 - H0)) < 1e-9
        lu_g, lu_c = np.log(rng_gpu.random()), np.log(rng_cpu.random())
        acc_g = lu_g < min(0.0, -res.delta_hamiltonian)
        acc_c = lu_c < min(0.0, -(H1 - H0))
        margins.append(abs(lu_c + (H1 - H0)))
        seq_gpu.append(bool(acc_g))
        seq_cpu.append(bool(acc_c))
        if acc_c:
            Uh = Uo
        if not acc_g:
            gfb200.copy_configuration_(U, old)
        if acc_g == acc_c:
            assert np.abs(U.to_host() - Uh).max() < 1e-10
    assert seq_gpu == seq_cpu
    assert any(seq_cpu) and not all(seq_cpu)
    assert min(margins) > 1e-7


def test_gaussian_momenta_stream(backend, oracle):
    """same (seed, sweep) -> same field; matches the oracle's site streams to 2e-12 (random_fields_site_rng.jl:225-235)."""
    import gfb200

    dims = DIMS_ANISO
    U = gfb200.gauge_configuration(dims, backend=backend)
    a = gfb200.gaussian_momenta(U, sigma=1.5, seed=300, sweep=4).to_host()
    b = gfb200.gaussian_momenta(U, sigma=1.5, seed=300, sweep=4).to_host()
    c = gfb200.gaussian_momenta(U, sigma=1.5, seed=300, sweep=5).to_host()
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert not np.array_equal(a[0], a[1])
    want = oracle.gaussian_momenta(dims, 300, 4, 1.5)
    assert np.allclose(a, want, rtol=2e-12, atol=2e-12)
    assert abs(a.mean()) < 0.04 * 1.5 and abs(a.std() - 1.5) < 0.04 * 1.5


def test_hot_start(backend, oracle):
    import gfb200

    dims = DIMS_ANISO
    U = gfb200.gauge_configuration(dims, backend=backend, start="hot", seed=1234)
    got = U.to_host()
    want = oracle.hot_start_philox(dims, 1234)
    assert np.abs(got - want).max() < 1e-14
    m = oracle.mats(got)
    assert np.abs(m @ m.conj().swapaxes(-1, -2) - np.eye(3)).max() < 2e-12
    assert not np.array_equal(got[0], got[1])


def test_argument_errors(backend):
    import gfb200

    with pytest.raises(ValueError):
        gfb200.gauge_configuration((4, 4, 4, 0), backend=backend)
    U = gfb200.gauge_configuration((4, 4, 4, 4), backend=backend)
    action = gfb200.GaugeAction(U).push(3.0, gfb200.make_loops_fromname("plaquette") + gfb200.make_loops_fromname("plaquette").adjoint())
    with pytest.raises(ValueError):
        gfb200.md_driver(U, action, steps=0)
    with pytest.raises(ValueError):
        gfb200.md_driver(U, action, steps=4, trajectory_length=0.0)
    p = gfb200.gauge_momenta(U)
    md = gfb200.md_driver(U, action, steps=4)
    with pytest.raises(ValueError):
        gfb200.update_gaugefields_(U, p, float("nan"), md)
    with pytest.raises(ValueError):
        backend.call("gfb_md_trajectory", U._h, p._h, 6.0, -1, 1.0, 0, 0, None)


def test_non_unitary_links_take_the_full_product_path(backend, oracle):
    """The fused kernels form staples from two rows of their (unitary) factors.  A configuration that is NOT unitary to 1e-12
    -- e.g. a single-precision ILDG file, or any array a user uploads -- must still give the reference's numbers: the
    reference's staples (src/AbstractGaugefields.jl:2041-2066) hold for arbitrary 3x3 matrices.  The library checks unitarity
    once per uploaded configuration and switches every pass on it to full 3x3 products; gfb_reunitarize switches back."""
    import gfb200

    dims, beta = (8, 4, 2, 4), 5.8
    rng = np.random.default_rng(7)
    Uh = oracle.hot_start_philox(dims, 11)
    Uh = Uh + 1e-7 * (rng.standard_normal(Uh.shape) + 1j * rng.standard_normal(Uh.shape))
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    loops = gfb200.make_loops_fromname("plaquette")
    action = gfb200.GaugeAction(U).push(beta / 2, loops + loops.adjoint())
    F = gfb200.gauge_momenta(U)
    gfb200.md_force_(F, action, U)
    want = oracle.force(Uh, dims, beta)
    assert np.abs(F.to_host() - want).max() < 1e-12 * np.abs(want).max()
    # the two-row shortcut would be off by ~1e-7 here: make sure this test can tell
    Uu = Uh.copy()
    oracle.reunitarize(Uu, dims)
    assert np.abs(oracle.force(Uu, dims, beta) - want).max() > 1e-9 * np.abs(want).max()
    # a fused trajectory on the non-unitary field follows the oracle (which multiplies whatever it is given)
    Ph = oracle.gaussian_momenta(dims, 0x5678, 4)
    P = gfb200.gauge_momenta(U).upload(Ph)
    md = gfb200.md_driver(U, action, steps=4, trajectory_length=0.2, integrator=gfb200.QPQ, fused=True)
    res = gfb200.md_trajectory_(U, P, md)
    Uo, Po = Uh.copy(), Ph.copy()
    H0, H1 = oracle.md_trajectory(Uo, Po, dims, beta, 4, 0.2, 0)
    assert abs(res.delta_hamiltonian - (H1 - H0)) < 1e-9
    assert np.abs(U.to_host() - Uo).max() < 1e-11
    # the plaquette-staple stout layer and the flow as well
    U.upload(Uh)
    out = gfb200.smear(U, gfb200.stout_smearing(U, rho=0.1, layers=1))
    assert np.abs(out.to_host() - oracle.stout_forward(Uh, dims, 0.1)).max() < 1e-12
    # normalize_U! (reunitarize_) puts the configuration back on the fast path and on the group manifold
    gfb200.reunitarize_(U)
    got = U.to_host()
    assert np.abs(got - Uu).max() < 1e-13
    gfb200.md_force_(F, action, U)
    wu = oracle.force(Uu, dims, beta)
    assert np.abs(F.to_host() - wu).max() < 1e-12 * np.abs(wu).max()
