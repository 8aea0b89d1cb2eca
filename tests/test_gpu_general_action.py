"""GPU parity of the general-action path (csrc/general.cu): plaquette + rectangle actions (force, Hamiltonian, trajectories,
Gradientflow_general), the Sexton-Weingarten nested integrator over an MDActionSet, and the topological charge.

The oracle side is the GENERIC restatement (oracle/gf_oracle.cpp, "General-action path"): loops are differentiated by rotating
every loop of (set + adjoint set) to each of its +mu steps, as calc_dSdUmu! / make_staple do in the reference
(src/action/GaugeActions.jl:95-123); the CUDA kernels use 18 hand-derived rectangle staples.  Tolerances: north_star's
(force 1e-12 relative, Delta H 1e-9, links 1e-11, flow 1e-12)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# beta = 6/g^2-type couplings: tree-level Symanzik (c1 = -1/12), Iwasaki (c1 = -0.331), DBW2 (c1 = -1.4088): c0 = 1 - 8 c1
ACTIONS = {"symanzik": (4.5 / 2 * (1 + 8 / 12), 4.5 / 2 * (-1 / 12)), "iwasaki": (2.6 / 2 * (1 + 8 * 0.331), 2.6 / 2 * (-0.331)),
           "dbw2": (0.9 / 2 * (1 + 8 * 1.4088), 0.9 / 2 * (-1.4088)), "rect_only": (0.0, 0.7)}


def _action(gfb200, U, cp, cr):
    a = gfb200.GaugeAction(U)
    if cp != 0.0:
        p = gfb200.make_loops_fromname("plaquette")
        a.push(cp, p + p.adjoint())
    if cr != 0.0:
        r = gfb200.make_loops_fromname("rectangular")
        a.push(cr, r + r.adjoint())
    return a


@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (8, 4, 6, 4), (6, 10, 4, 8)])
@pytest.mark.parametrize("name", sorted(ACTIONS))
def test_general_force_action_and_hamiltonian(backend, oracle, dims, name):
    import gfb200

    cp, cr = ACTIONS[name]
    Uh = oracle.hot_start_philox(dims, 21)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    action = _action(gfb200, U, cp, cr)
    F = gfb200.gauge_momenta(U)
    gfb200.md_force_(F, action, U)
    want = oracle.force_general(Uh, dims, cp, cr)
    assert np.abs(F.to_host() - want).max() < 1e-12 * np.abs(want).max()
    sp, sr = oracle.loop_sums(Uh, dims)
    got = gfb200.evaluate_GaugeAction(action, U).real
    assert abs(got - 2 * (cp * sp + cr * sr)) <= 1e-12 * (abs(cp * sp) + abs(cr * sr))
    Ph = oracle.gaussian_momenta(dims, 0x5678, 3)
    P = gfb200.gauge_momenta(U).upload(Ph)
    md = gfb200.md_driver(U, action, steps=1, trajectory_length=0.1)
    h, hw = gfb200.md_hamiltonian(U, P, md), oracle.hamiltonian_general(Uh, Ph, dims, cp, cr)
    assert abs(h - hw) <= 1e-12 * abs(hw)


@pytest.mark.parametrize("integ", ["QPQ", "PQP"])
@pytest.mark.parametrize("fused", [True, False])
def test_general_trajectory_matches_oracle(backend, oracle, integ, fused):
    import gfb200

    dims = (4, 6, 4, 8)
    cp, cr = ACTIONS["symanzik"]
    Uh = oracle.hot_start_philox(dims, 5)
    for _ in range(3):
        oracle.flow_step(Uh, dims, 0.02)
    Ph = oracle.gaussian_momenta(dims, 0x5678, 1)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    P = gfb200.gauge_momenta(U).upload(Ph)
    action = _action(gfb200, U, cp, cr)
    I = getattr(gfb200, integ)
    md = gfb200.md_driver(U, action, steps=5, trajectory_length=0.25, integrator=I, fused=fused)
    res = gfb200.md_trajectory_(U, P, md)
    Uo, Po = Uh.copy(), Ph.copy()
    H0, H1 = oracle.md_trajectory_general(Uo, Po, dims, cp, cr, 5, 0.25, I.code)
    assert abs(res.initial_hamiltonian - H0) <= 1e-12 * abs(H0)
    assert abs(res.delta_hamiltonian - (H1 - H0)) < 1e-9
    assert np.abs(U.to_host() - Uo).max() < 1e-11
    assert np.abs(P.to_host() - Po).max() < 1e-10


def test_general_flow_matches_oracle_and_reduces_to_wilson_flow(backend, oracle):
    import gfb200

    dims = (4, 6, 4, 8)
    Uh = oracle.hot_start_philox(dims, 9)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    # Symanzik-type flow ("Zeuthen/Symanzik flow" kernel): link values (5/3, -1/12)
    g = gfb200.Gradientflow_general(U, ["plaquette", "rectangular"], [5 / 3, -1 / 12], Nflow=3, eps=0.01)
    gfb200.flow_(U, g)
    Uo = Uh.copy()
    for _ in range(3):
        oracle.flow_step_general(Uo, dims, 0.01, 5 / 3, -1 / 12)
    assert np.abs(U.to_host() - Uo).max() < 1e-12
    # link values (1, 0) are the Wilson flow
    U.upload(Uh)
    gfb200.flow_(U, gfb200.Gradientflow_general(U, ["plaquette"], [1.0], Nflow=2, eps=0.01))
    Uw = Uh.copy()
    for _ in range(2):
        oracle.flow_step(Uw, dims, 0.01)
    assert np.abs(U.to_host() - Uw).max() < 1e-12


@pytest.mark.parametrize("ordering", ["QPQ", "PQP"])
def test_sexton_weingarten_matches_oracle(backend, oracle, ordering):
    """SextonWeingarten over an MDActionSet (src/molecular_dynamics.jl:57-236, 618-700): the plaquette term on the fast
    (inner, n_fast = 3) level, the rectangle term on the slow level; Delta H of the full action against the oracle."""
    import gfb200

    dims = (4, 4, 6, 4)
    cp, cr = ACTIONS["iwasaki"]
    Uh = oracle.hot_start_philox(dims, 2)
    for _ in range(3):
        oracle.flow_step(Uh, dims, 0.02)
    Ph = oracle.gaussian_momenta(dims, 0x5678, 7)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    P = gfb200.gauge_momenta(U).upload(Ph)
    actions = gfb200.MDActionSet(plaq=_action(gfb200, U, cp, 0.0), rect=_action(gfb200, U, 0.0, cr))
    O = getattr(gfb200, ordering)
    sw = gfb200.SextonWeingarten(fast="plaq", slow=("rect",), n_fast=3, ordering=O)
    md = gfb200.md_driver(U, actions, steps=4, trajectory_length=0.2, integrator=sw)
    res = gfb200.md_trajectory_(U, P, md)
    Uo, Po = Uh.copy(), Ph.copy()
    H0, H1 = oracle.sexton_weingarten_trajectory(Uo, Po, dims, {"plaq": (cp, 0.0), "rect": (0.0, cr)}, ("plaq",), ("rect",), 3, 4, 0.2, O.code)
    assert abs(res.initial_hamiltonian - H0) <= 1e-12 * abs(H0)
    assert abs(res.delta_hamiltonian - (H1 - H0)) < 1e-9
    assert np.abs(U.to_host() - Uo).max() < 1e-11
    assert np.abs(P.to_host() - Po).max() < 1e-10
    # the selected-force kick alone (update_momenta! with a group) and the error contract
    F = gfb200.gauge_momenta(U)
    gfb200.md_force_(F, actions, U, group=gfb200.MDForceGroup("rect"))
    want = oracle.force_general(Uo, dims, 0.0, cr)
    assert np.abs(F.to_host() - want).max() < 1e-12 * np.abs(want).max()
    with pytest.raises(ValueError):
        gfb200.SextonWeingarten(fast="plaq", slow="plaq", n_fast=2)
    with pytest.raises(ValueError):
        gfb200.md_driver(U, actions, steps=2, integrator=gfb200.SextonWeingarten(fast="plaq", slow="nope", n_fast=2))
    with pytest.raises(ValueError):
        gfb200.MDForceGroup("a", "a")


@pytest.mark.parametrize("method,code", [("plaquette", 0), ("clover", 1), ("improved", 2)])
def test_topological_charge_matches_oracle(backend, oracle, method, code):
    """topological_charge / topological_charge_density (src/AbstractGaugefields.jl:1184-1490) on a smoothed configuration;
    the reference's own contract sum(density) == charge (test/latticematrices_compat.jl:480-485) as well."""
    import gfb200

    dims = (4, 6, 4, 8)
    Uh = oracle.hot_start_philox(dims, 17)
    for _ in range(5):
        oracle.flow_step(Uh, dims, 0.02)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    want = oracle.topological_charge_density(Uh, dims, code)
    got = gfb200.topological_charge_density(U, method=method)
    assert got.shape == (8, 4, 6, 4)
    assert np.abs(got - want).max() < 1e-13 * max(1.0, np.abs(want).max())
    q = gfb200.topological_charge(U, method=":" + method)
    assert abs(q - want.sum()) < 1e-12 * max(1.0, abs(want).sum())
    assert abs(got.sum() - q) < 1e-12 * max(1.0, abs(q))
    with pytest.raises(ValueError):
        gfb200.topological_charge(U, method="wilson")


def test_cold_configuration_has_no_force_and_no_charge(backend, oracle):
    import gfb200

    dims = (4, 4, 4, 4)
    U = gfb200.gauge_configuration(dims, backend=backend)  # cold
    action = _action(gfb200, U, *ACTIONS["symanzik"])
    F = gfb200.gauge_momenta(U)
    gfb200.md_force_(F, action, U)
    assert np.abs(F.to_host()).max() == 0.0
    v = gfb200.evaluate_GaugeAction(action, U).real
    cp, cr = ACTIONS["symanzik"]
    assert abs(v - 2 * 256 * (cp * 18 + cr * 36)) < 1e-9
    for m in ("plaquette", "clover", "improved"):
        assert gfb200.topological_charge(U, method=m) == 0.0
