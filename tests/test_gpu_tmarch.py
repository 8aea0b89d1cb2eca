"""GPU parity of the persistent t-marching kernel (csrc/tmarch.cu), which takes over the fused passes whenever the
8x4x2 spatial tile divides the lattice (all BASELINE.json bench lattices).  Same tolerances as the generic kernel
(BASELINE.json north_star): force 1e-12 relative, dH 1e-9, links 1e-11, flow E(t) 1e-11.  The t-segment length is forced
through GFB200_TMARCH_SEGLEN so that segment starts (backward-t staple from global memory, ring prologue), the ring
rotation (segments longer than 3) and ragged last segments are all exercised."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture
def seglen():
    def set_(n):
        if n is None:
            os.environ.pop("GFB200_TMARCH_SEGLEN", None)
        else:
            os.environ["GFB200_TMARCH_SEGLEN"] = str(n)

    yield set_
    os.environ.pop("GFB200_TMARCH_SEGLEN", None)


@pytest.mark.parametrize("dims,seg", [((8, 4, 2, 4), None), ((8, 4, 2, 7), 3), ((16, 8, 4, 5), 5), ((8, 8, 6, 6), 1), ((24, 4, 4, 4), 2)])
def test_tmarch_force_and_fused_step(backend, oracle, seglen, dims, seg):
    import gfb200

    seglen(seg)
    beta = 5.9
    Uh = oracle.hot_start_philox(dims, 31)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    loops = gfb200.make_loops_fromname("plaquette")
    action = gfb200.GaugeAction(U).push(beta / 2, loops + loops.adjoint())
    F = gfb200.gauge_momenta(U)
    gfb200.md_force_(F, action, U)
    want = oracle.force(Uh, dims, beta)
    assert np.abs(F.to_host() - want).max() < 1e-12 * np.abs(want).max()
    Ph = oracle.gaussian_momenta(dims, 0x5678, 1)
    for integ in (gfb200.QPQ, gfb200.PQP):
        U.upload(Uh)
        P = gfb200.gauge_momenta(U).upload(Ph)
        md = gfb200.md_driver(U, action, steps=6, trajectory_length=0.3, integrator=integ, fused=True)
        res = gfb200.md_trajectory_(U, P, md)
        Uo, Po = Uh.copy(), Ph.copy()
        H0, H1 = oracle.md_trajectory(Uo, Po, dims, beta, 6, 0.3, integ.code)
        assert abs(res.initial_hamiltonian - H0) <= 1e-12 * abs(H0)
        assert abs(res.delta_hamiltonian - (H1 - H0)) < 1e-9
        assert np.abs(U.to_host() - Uo).max() < 1e-11
        assert np.abs(P.to_host() - Po).max() < 1e-10


def test_tmarch_matches_generic_kernel_bitwise_tolerance(backend, oracle, seglen):
    """Same lattice through both kernels: the results may differ only at rounding level (different summation order)."""
    import gfb200

    dims = (16, 8, 4, 8)
    Uh = oracle.hot_start_philox(dims, 5)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    loops = gfb200.make_loops_fromname("plaquette")
    action = gfb200.GaugeAction(U).push(3.0, loops + loops.adjoint())
    F = gfb200.gauge_momenta(U)
    seglen(4)
    gfb200.md_force_(F, action, U)
    a = F.to_host().copy()
    seglen(None)
    os.environ["GFB200_TMARCH"] = "0"
    try:
        gfb200.md_force_(F, action, U)
    finally:
        os.environ.pop("GFB200_TMARCH", None)
    b = F.to_host().copy()
    assert np.abs(a - b).max() < 1e-13 * np.abs(a).max()
    want = oracle.force(Uh, dims, 6.0)
    assert np.abs(a - want).max() < 1e-12 * np.abs(want).max()


def test_tmarch_flow_and_stout(backend, oracle, seglen):
    import gfb200

    seglen(3)
    dims = (8, 8, 4, 5)
    Uh = oracle.hot_start_philox(dims, 8)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    gfb200.flow_(U, gfb200.gradient_flow(U, steps=3, step_size=0.01))
    Uo = Uh.copy()
    for _ in range(3):
        oracle.flow_step(Uo, dims, 0.01)
    assert np.abs(U.to_host() - Uo).max() < 1e-12
    assert abs(gfb200.energy_density(U) - oracle.energy_density_clover(Uo, dims)) < 1e-11
    out = gfb200.smear(U, gfb200.stout_smearing(U, rho=0.1, layers=2))
    want = oracle.stout_forward(oracle.stout_forward(Uo, dims, 0.1), dims, 0.1)
    assert np.abs(out.to_host() - want).max() < 1e-12


def test_tmarch_reversibility_full_size_property(backend):
    """Size-independent property at a bench-like lattice the oracle cannot reach: a QPQ trajectory followed by the
    momentum-flipped trajectory returns to the start (test/md_driver.jl:328-368 in the reference)."""
    import gfb200

    dims = (32, 32, 16, 8)
    U = gfb200.gauge_configuration(dims, backend=backend, start="hot", seed=77)
    U0 = U.to_host().copy()
    P = gfb200.gaussian_momenta(U, seed=3, sweep=0)
    loops = gfb200.make_loops_fromname("plaquette")
    action = gfb200.GaugeAction(U).push(3.0, loops + loops.adjoint())
    md = gfb200.md_driver(U, action, steps=10, trajectory_length=0.5, integrator=gfb200.QPQ, fused=True)
    r1 = gfb200.md_trajectory_(U, P, md)
    P.upload(-P.to_host())
    r2 = gfb200.md_trajectory_(U, P, md)
    assert np.abs(U.to_host() - U0).max() < 2e-12
    assert abs(r1.delta_hamiltonian + r2.delta_hamiltonian) < 1e-7 * abs(r1.initial_hamiltonian)


def test_tmarch_round_barrier_does_not_change_results(backend):
    """The grid-wide round barrier of the persistent grid (TmArgs::round_ctr) only aligns the CTAs in time: a fused trajectory with
    the barrier forced on (GFB200_TMARCH_ROUNDSYNC=2; by default it is on for marches of >= 32 slices, i.e. the 64^4 benchmark) must
    be bitwise identical to the one without it.  32x32x32x8: 512 tiles = 3.5 rounds of the 148 CTAs."""
    import gfb200

    dims = (32, 32, 32, 8)
    U = gfb200.gauge_configuration(dims, backend=backend, start="hot", seed=92)
    P = gfb200.gaussian_momenta(U, seed=5, sweep=0)
    U0, P0 = gfb200.copy_configuration(U), P.to_host().copy()
    loops = gfb200.make_loops_fromname("plaquette")
    action = gfb200.GaugeAction(U).push(3.0, loops + loops.adjoint())
    md = gfb200.md_driver(U, action, steps=4, trajectory_length=0.2, integrator=gfb200.QPQ, fused=True)
    out = []
    for mode in ("0", "2"):
        gfb200.copy_configuration_(U, U0)
        P.upload(P0)
        os.environ["GFB200_TMARCH_ROUNDSYNC"] = mode
        try:
            r = gfb200.md_trajectory_(U, P, md)
        finally:
            os.environ.pop("GFB200_TMARCH_ROUNDSYNC", None)
        out.append((U.to_host().copy(), P.to_host().copy(), r.delta_hamiltonian))
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
    assert out[0][2] == out[1][2]
