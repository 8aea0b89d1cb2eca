"""Multi-GPU parity (t-slab decomposition + NCCL halo exchange).

Needs >= 2 GPUs on the box (`gpurun --gpus 2`); skipped otherwise.  Covers both launch forms:
one process driving two GPUs (gfb_init) and one process per GPU under torchrun (gfb_init_rank)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    import torch

    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def backend2():
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    import gfb200

    b = gfb200.B200Backend(ngpu=2)
    yield b


DIMS = (4, 6, 4, 8)


# (8, 4, 2, 12): the 8x4x2 tile divides the lattice, so whole-slab passes and the interior slices of the split passes run in
# the t-marching kernel (halo slots read through TMA and through the segment prologue), the two boundary slices in k_force_fused
@pytest.mark.parametrize("dims", [DIMS, (8, 4, 2, 12)])
def test_two_slabs_match_oracle(backend2, oracle, dims, monkeypatch):
    import gfb200

    # slabs this thin would stay in k_force_fused (the library keeps interiors under 8 slices there): force the tile kernel
    monkeypatch.setenv("GFB200_TMARCH", "2")
    Uh = oracle.hot_start_philox(dims, 1234)
    U = gfb200.gauge_configuration(dims, backend=backend2).upload(Uh)
    assert gfb200.gauge_process_grid(U) == (1, 1, 1, 2)
    assert np.array_equal(U.to_host(), Uh)
    want = oracle.plaquette_sum(Uh, dims)
    assert abs(gfb200.calculate_Plaquette(U) - want) <= 1e-12 * abs(want) + 1e-12
    loops = gfb200.make_loops_fromname("plaquette")
    action = gfb200.GaugeAction(U).push(5.7 / 2, loops + loops.adjoint())
    F = gfb200.gauge_momenta(U)
    gfb200.md_force_(F, action, U)
    Fw = oracle.force(Uh, dims, 5.7)
    assert np.abs(F.to_host() - Fw).max() / np.abs(Fw).max() < 1e-12
    e = gfb200.energy_density(U)
    assert abs(e - oracle.energy_density_clover(Uh, dims)) < 1e-11 * max(1.0, abs(e))
    # Polyakov loop: the running product is handed from slab to slab in t order
    assert abs(gfb200.measure_polyakov_loop(U, normalize=False) - oracle.polyakov(Uh, dims)) < 1e-13
    for integ in (gfb200.QPQ, gfb200.PQP):
        for fused in (False, True):
            U.upload(Uh)
            Ph = oracle.gaussian_momenta(dims, 0x5678, 1)
            P = gfb200.gauge_momenta(U).upload(Ph)
            md = gfb200.md_driver(U, action, steps=10, trajectory_length=0.5, integrator=integ, fused=fused)
            res = gfb200.md_trajectory_(U, P, md)
            Uo = Uh.copy()
            H0, H1 = oracle.md_trajectory(Uo, Ph, dims, 5.7, 10, 0.5, integ.code)
            assert abs(res.delta_hamiltonian - (H1 - H0)) < 1e-9
            assert np.abs(U.to_host() - Uo).max() < 1e-11
    U.upload(Uh)
    gfb200.flow_(U, gfb200.gradient_flow(U, steps=3, step_size=0.01))
    Uo = Uh.copy()
    for _ in range(3):
        oracle.flow_step(Uo, dims, 0.01)
    assert np.abs(U.to_host() - Uo).max() < 1e-12
    U.upload(Uh)
    out = gfb200.smear(U, gfb200.stout_smearing(U, rho=0.1, layers=2))
    want = oracle.stout_forward(oracle.stout_forward(Uh, dims, 0.1), dims, 0.1)
    assert np.abs(out.to_host() - want).max() < 1e-12
    # stout backward across the slab faces (halo of U and of Lambda)
    sm = gfb200.stout_smearing(U, rho=0.1, layers=2)
    Uout, multi = gfb200.calc_smearedU(U, sm)
    bare = gfb200.back_prop(gfb200.calc_dSdU(action, Uout), sm, multi, U)
    U1 = oracle.stout_forward(Uh, dims, 0.1)
    d0 = oracle.stout_backward(oracle.stout_backward(oracle.wilson_dSdU(want, dims, 5.7), U1, dims, 0.1), Uh, dims, 0.1)
    assert np.abs(bare.to_host() - d0).max() < 1e-12 * np.abs(d0).max()
    # unfused primitives with the overlapped exchange: link update then force on the fresh halo
    U.upload(Uh)
    Ph = oracle.gaussian_momenta(dims, 3, 0)
    P = gfb200.gauge_momenta(U).upload(Ph)
    md = gfb200.md_driver(U, action, steps=4)
    gfb200.update_gaugefields_(U, P, 0.1, md)
    gfb200.md_force_(F, action, U)
    Fw = oracle.force(oracle.update_links(Uh, Ph, dims, 0.1), dims, 5.7)
    assert np.abs(F.to_host() - Fw).max() / np.abs(Fw).max() < 1e-12


@pytest.mark.parametrize("dims", [(4, 6, 4, 8), (8, 4, 2, 4)])
def test_general_action_across_slabs(backend2, oracle, dims):
    """Rectangle terms reach two slices across a slab face: the general-action kernels run on a wide copy of each slab
    (two halo slices either side, csrc/api.cu build_wide).  (8, 4, 2, 4) has two slices per slab: both halos are the whole
    neighbour."""
    import gfb200

    cp, cr = 4.5 / 2 * (1 + 8 / 12), 4.5 / 2 * (-1 / 12)
    Uh = oracle.hot_start_philox(dims, 77)
    for _ in range(3):
        oracle.flow_step(Uh, dims, 0.02)
    U = gfb200.gauge_configuration(dims, backend=backend2).upload(Uh)
    p, r = gfb200.make_loops_fromname("plaquette"), gfb200.make_loops_fromname("rectangular")
    action = gfb200.GaugeAction(U).push(cp, p + p.adjoint()).push(cr, r + r.adjoint())
    F = gfb200.gauge_momenta(U)
    gfb200.md_force_(F, action, U)
    want = oracle.force_general(Uh, dims, cp, cr)
    assert np.abs(F.to_host() - want).max() < 1e-12 * np.abs(want).max()
    sp, sr = oracle.loop_sums(Uh, dims)
    assert abs(gfb200.evaluate_GaugeAction(action, U).real - 2 * (cp * sp + cr * sr)) < 1e-12 * (abs(cp * sp) + abs(cr * sr))
    Ph = oracle.gaussian_momenta(dims, 0x5678, 5)
    for integ in (gfb200.QPQ, gfb200.PQP):
        for fused in (True, False):
            U.upload(Uh)
            P = gfb200.gauge_momenta(U).upload(Ph)
            md = gfb200.md_driver(U, action, steps=4, trajectory_length=0.2, integrator=integ, fused=fused)
            res = gfb200.md_trajectory_(U, P, md)
            Uo, Po = Uh.copy(), Ph.copy()
            H0, H1 = oracle.md_trajectory_general(Uo, Po, dims, cp, cr, 4, 0.2, integ.code)
            assert abs(res.delta_hamiltonian - (H1 - H0)) < 1e-9
            assert np.abs(U.to_host() - Uo).max() < 1e-11
            assert np.abs(P.to_host() - Po).max() < 1e-10
    # a Wilson pass right after a rectangle pass must see a fresh one-slice halo (the wide path does not fill it)
    wil = gfb200.GaugeAction(U).push(2.9, p + p.adjoint())
    gfb200.md_force_(F, wil, U)
    Fw = oracle.force(Uo, dims, 5.8)
    assert np.abs(F.to_host() - Fw).max() < 1e-12 * np.abs(Fw).max()
    U.upload(Uh)
    gfb200.flow_(U, gfb200.Gradientflow_general(U, ["plaquette", "rectangular"], [5 / 3, -1 / 12], Nflow=2, eps=0.01))
    Uf = Uh.copy()
    for _ in range(2):
        oracle.flow_step_general(Uf, dims, 0.01, 5 / 3, -1 / 12)
    assert np.abs(U.to_host() - Uf).max() < 1e-12
    for code, method in enumerate(("plaquette", "clover", "improved")):
        wq = oracle.topological_charge_density(Uf, dims, code)
        got = gfb200.topological_charge_density(U, method=method)
        assert np.abs(got - wq).max() < 1e-13 * max(1.0, np.abs(wq).max())
        assert abs(gfb200.topological_charge(U, method=method) - wq.sum()) < 1e-12 * max(1.0, np.abs(wq).sum())


def test_primitive_table_across_slabs(backend2, oracle):
    """shifted / adjoint mul! and tr with t-shifts that cross the slab faces (field-level halo exchange)."""
    import gfb200 as g

    dims = DIMS
    Uh = oracle.hot_start_philox(dims, 21)
    U = g.gauge_configuration(dims, backend=backend2).upload(Uh)
    L = [g.link_field(U, mu) for mu in range(4)]
    t1, t2, V = L[0].similar(), L[0].similar(), L[0].similar()
    plaq = 0.0
    for mu in range(4):
        g.clear_U_(V)
        for nu in range(4):
            if nu == mu:
                continue
            g.mul_(t1, L[nu], g.shift_U(L[mu], nu + 1))
            g.mul_(V, t1, g.shift_U(L[nu], mu + 1).H, 1.0, 1.0)
            # lower staple too, so backward t-shifts of views and temporaries are exercised
            g.mul_(t1, g.shift_U(L[nu], -(nu + 1)).H, g.shift_U(L[mu], -(nu + 1)))
            sh = [0, 0, 0, 0]
            sh[mu] += 1
            sh[nu] -= 1
            g.mul_(V, t1, g.shift_U(L[nu], sh), 1.0, 1.0)
        g.mul_(t2, L[mu], V.H)
        plaq += g.tr(t2)
    want = oracle.plaquette_sum(Uh, dims)
    assert abs(plaq.real * 0.25 - want) < 1e-12 * abs(want) + 1e-12  # upper + lower staples count every plaquette four times
    B = t1.similar()
    g.substitute_U_(B, g.shift_U(t2, (0, 0, 0, 1)))
    assert np.array_equal(B.to_host(), np.roll(t2.to_host(), -1, axis=0))
    with pytest.raises(ValueError):
        g.substitute_U_(B, g.shift_U(t2, (0, 0, 0, 2)))  # wider than the halo


def test_random_fields_are_decomposition_independent(backend2, backend):
    """hot start and Gaussian momenta are keyed by the GLOBAL site: bit-identical on 1 and 2 slabs
    (test/MPIJACCtest/random_fields_site_rng.jl:148-172)."""
    import gfb200

    dims = DIMS
    a = gfb200.gauge_configuration(dims, backend=backend, start="hot", seed=99).to_host()
    b = gfb200.gauge_configuration(dims, backend=backend2, start="hot", seed=99).to_host()
    assert np.array_equal(a, b)
    U1 = gfb200.gauge_configuration(dims, backend=backend)
    U2 = gfb200.gauge_configuration(dims, backend=backend2)
    pa = gfb200.gaussian_momenta(U1, seed=5, sweep=7).to_host()
    pb = gfb200.gaussian_momenta(U2, seed=5, sweep=7).to_host()
    assert np.array_equal(pa, pb)


def test_heatbath_is_decomposition_independent(backend2, backend, oracle):
    """Streams are keyed by the global site and the sweep refreshes the slab halos between colours: 1 and 2 slabs agree bitwise
    (the reference's contract for its site-RNG kernels, src/heatbath/heatbathmodule.jl:1624-1632)."""
    import gfb200

    dims = (4, 6, 4, 8)
    Uh = oracle.hot_start_philox(dims, 12)
    outs = []
    for b in (backend, backend2):
        U = gfb200.gauge_configuration(dims, backend=b).upload(Uh)
        h = gfb200.Heatbath(U, 5.9, seed=77)
        gfb200.heatbath_(U, h)
        gfb200.overrelaxation_(U, h)
        gfb200.heatbath_(U, h)
        outs.append(U.to_host().copy())
    assert np.array_equal(outs[0], outs[1])


def test_slab_constraints(backend2):
    import gfb200

    with pytest.raises(ValueError):
        gfb200.gauge_configuration((4, 4, 4, 5), backend=backend2)  # NT not divisible by 2
    with pytest.raises(ValueError):
        gfb200.gauge_configuration((4, 4, 4, 2), backend=backend2)  # one slice per slab


@pytest.mark.timeout(600)
def test_one_process_per_gpu_under_torchrun():
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "scripts", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=500, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dist_check ok" in r.stdout
