"""Generates tests/golden/wilson_4x4x4x4.npz and wilson_8x4x2x4.npz -- known-answer vectors for the hot path on the reference's
own test lattice (4^4, SU(3), Wilson beta = 5.7: test/HMC_test.jl scale, BASELINE.json configs[0]) and on a lattice of the same
volume that the 8x4x2 tile of the t-marching kernel divides (4^4 runs in k_force_fused, 8x4x2x4 in k_tmarch_fused).

The reference cannot run in the build image (Julia + un-vendored LatticeMatrices.jl), so these vectors come from the CPU
oracle (oracle/gf_oracle.cpp), which is itself pinned to the reference's golden values in tests/test_oracle_pins.py; two of
those reference values are stored here as well.  The fixtures let the GPU tests check the CUDA path against committed
numbers even where the oracle library is not rebuilt, and freeze today's answers against silent drift of either side.

    python tests/golden/make_golden.py        # rewrites the .npz (deterministic)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import gf_oracle as oracle  # noqa: E402

BETA = 5.7


def make(DIMS):
    out = {}
    U = oracle.hot_start_philox(DIMS, 1234)
    P = oracle.gaussian_momenta(DIMS, 0x5678, 0)
    out["U0"] = U.copy()
    out["P0"] = P.copy()
    out["plaquette_sum"] = oracle.plaquette_sum(U, DIMS)
    out["force"] = oracle.force(U, DIMS, BETA)
    out["kinetic"] = oracle.momentum_norm2(P, DIMS)
    out["energy_clover"] = oracle.energy_density_clover(U, DIMS)
    for name, integ in (("qpq", 0), ("pqp", 1)):
        Ut, Pt = U.copy(), P.copy()
        H0, H1 = oracle.md_trajectory(Ut, Pt, DIMS, BETA, 20, 1.0, integ)
        out["H0_" + name], out["dH_" + name] = H0, H1 - H0
        if integ == 0:
            out["U_after_qpq"] = Ut
        else:
            out["U_after_pqp_mu0_t0"] = Ut[0, 0].copy()  # one time-slice of one direction keeps the fixture small
    # flow: E(t) series (clover, plaquette form) over 10 RK3 steps of eps = 0.01
    Uf = U.copy()
    e_c, e_p = [], []
    for _ in range(10):
        oracle.flow_step(Uf, DIMS, 0.01)
        e_c.append(oracle.energy_density_clover(Uf, DIMS))
        e_p.append(2.0 * (18.0 - oracle.plaquette_sum(Uf, DIMS) / np.prod(DIMS)))
    out["flow_E_clover"], out["flow_E_plaquette"] = np.array(e_c), np.array(e_p)
    out["U_flowed_mu3_t3"] = Uf[3, 3].copy()
    # stout: two layers rho = 0.1 forward, and the smeared-action force
    U1 = oracle.stout_forward(U, DIMS, 0.1)
    U2 = oracle.stout_forward(U1, DIMS, 0.1)
    out["U_stout2"] = U2
    d0 = oracle.stout_backward(oracle.stout_backward(oracle.wilson_dSdU(U2, DIMS, BETA), U1, DIMS, 0.1), U, DIMS, 0.1)
    out["stout_force"] = oracle.kick_from_dSdU(oracle.new_p(DIMS), U, d0, DIMS, -1.0 / 3.0)
    # the reference's own golden values (test/init.jl:276-283, test/gradientflow_test.jl:129-139)
    if DIMS == (4, 4, 4, 4):
        Us = oracle.hot_start_stable123(DIMS)
        out["ref_hot_plaquette"] = np.float64(0.008449494077606137)
        out["oracle_hot_plaquette"] = oracle.plaquette(Us, DIMS)
        for _ in range(100):
            oracle.flow_step(Us, DIMS, 0.01)
        out["ref_flow_plaquette"] = np.float64(0.8786515255315753)
        out["oracle_flow_plaquette"] = oracle.plaquette(Us, DIMS)
    name = "wilson_%s.npz" % "x".join(map(str, DIMS))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items()})


if __name__ == "__main__":
    make((4, 4, 4, 4))
    make((8, 4, 2, 4))
