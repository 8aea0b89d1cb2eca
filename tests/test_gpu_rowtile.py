"""GPU parity of the persistent TMA row-tile kernel (csrc/rowtile.cu), which takes over the fused passes when NX is
32 or 64.  Same tolerances as the generic kernel (BASELINE.json); lattices are thin in y, z, t so the oracle stays fast."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _enable_rowtile():
    """The row-tile kernel is opt-in (GFB200_ROWTILE=1, read once per process by the library): this module must run in a
    process that has not yet launched a fused pass, so it re-executes itself in a subprocess when needed."""
    yield


def _run_in_subprocess(test_name):
    import subprocess
    import sys

    env = dict(os.environ, GFB200_ROWTILE="1", GFB200_ROWTILE_CHILD="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", __file__ + "::" + test_name], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


IN_CHILD = os.environ.get("GFB200_ROWTILE_CHILD") == "1"


def test_rowtile_enabled_in_child_process():
    """Runs the two parity tests below with GFB200_ROWTILE=1 in a fresh process (the switch is read once per process)."""
    if IN_CHILD:
        pytest.skip("already in the child")
    _run_in_subprocess("test_rowtile_force_kick_trajectory")
    _run_in_subprocess("test_rowtile_flow_and_stout")


@pytest.mark.parametrize("dims", [(32, 4, 4, 4), (64, 4, 2, 4), (32, 2, 6, 2)])
def test_rowtile_force_kick_trajectory(backend, oracle, dims):
    import gfb200

    if not IN_CHILD:
        pytest.skip("exercised through test_rowtile_enabled_in_child_process")

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
        md = gfb200.md_driver(U, action, steps=8, trajectory_length=0.4, integrator=integ, fused=True)
        res = gfb200.md_trajectory_(U, P, md)
        Uo, Po = Uh.copy(), Ph.copy()
        H0, H1 = oracle.md_trajectory(Uo, Po, dims, beta, 8, 0.4, integ.code)
        assert abs(res.initial_hamiltonian - H0) <= 1e-12 * abs(H0)
        assert abs(res.delta_hamiltonian - (H1 - H0)) < 1e-9
        assert np.abs(U.to_host() - Uo).max() < 1e-11
        assert np.abs(P.to_host() - Po).max() < 1e-10


def test_rowtile_flow_and_stout(backend, oracle):
    import gfb200

    if not IN_CHILD:
        pytest.skip("exercised through test_rowtile_enabled_in_child_process")

    dims = (32, 4, 2, 4)
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
