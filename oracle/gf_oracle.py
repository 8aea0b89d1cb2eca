"""ctypes binding of the CPU oracle (oracle/gf_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(gaugefields.jl_b200/) never imports this module.

Arrays use the reference's host layout (src/API.jl:516-529):
  U : complex128, shape (4, NT, NZ, NY, NX, 3, 3) C-order with the LAST TWO axes
      being (column j, row i)  == Julia ComplexF64[3,3,NX,NY,NZ,NT] per direction
  P : float64,   shape (4, NT, NZ, NY, NX, 8)    == Julia Float64[8,1,NX,NY,NZ,NT]
`dims` is always (NX, NY, NZ, NT).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libgforacle.so")
    src = os.path.join(_HERE, "gf_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libgforacle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libgforacle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.orc_plaquette_sum.restype = ctypes.c_double
        _LIB.orc_momentum_norm2.restype = ctypes.c_double
        _LIB.orc_hamiltonian.restype = ctypes.c_double
        _LIB.orc_energy_density_clover.restype = ctypes.c_double
    return _LIB


def threads(n=0):
    """OpenMP threads of the oracle's loops; n > 0 sets the count first (torchrun exports OMP_NUM_THREADS=1)."""
    return int(lib().orc_threads(int(n)))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _dims(dims):
    return (ctypes.c_int * 4)(*[int(d) for d in dims])


def u_shape(dims):
    nx, ny, nz, nt = dims
    return (4, nt, nz, ny, nx, 3, 3)


def p_shape(dims):
    nx, ny, nz, nt = dims
    return (4, nt, nz, ny, nx, 8)


def new_u(dims):
    return np.zeros(u_shape(dims), dtype=np.complex128)


def new_p(dims):
    return np.zeros(p_shape(dims), dtype=np.float64)


def mats(U):
    """View with math indexing [..., i, j] (row, column)."""
    return np.swapaxes(U, -1, -2)


def set_cold(dims):
    U = new_u(dims)
    lib().orc_set_cold(_dp(U), _dims(dims))
    return U


def hot_start_stable123(dims):
    U = new_u(dims)
    lib().orc_hot_start_stable123(_dp(U), _dims(dims))
    return U


def hot_start_philox(dims, seed):
    U = new_u(dims)
    lib().orc_hot_start_philox(_dp(U), _dims(dims), ctypes.c_uint64(seed))
    return U


def gaussian_momenta(dims, seed, sweep, sigma=1.0):
    P = new_p(dims)
    lib().orc_gaussian_momenta(_dp(P), _dims(dims), ctypes.c_uint64(seed), ctypes.c_uint64(sweep), ctypes.c_double(sigma))
    return P


def reunitarize(U, dims):
    lib().orc_reunitarize(_dp(U), _dims(dims))
    return U


def plaquette_sum(U, dims):
    return lib().orc_plaquette_sum(_dp(U), _dims(dims))


def plaquette(U, dims):
    return plaquette_sum(U, dims) / (6.0 * np.prod(dims) * 3.0)


def momentum_norm2(P, dims):
    return lib().orc_momentum_norm2(_dp(P), _dims(dims))


def hamiltonian(U, P, dims, beta):
    return lib().orc_hamiltonian(_dp(U), _dp(P), _dims(dims), ctypes.c_double(beta))


def polyakov(U, dims):
    out = np.zeros(2)
    lib().orc_polyakov(_dp(U), _dims(dims), _dp(out))
    return complex(out[0], out[1])


def energy_density_clover(U, dims):
    return lib().orc_energy_density_clover(_dp(U), _dims(dims))


def force(U, dims, beta):
    F = new_p(dims)
    lib().orc_force(_dp(F), _dp(U), _dims(dims), ctypes.c_double(beta))
    return F


def flow_force(U, dims):
    F = new_p(dims)
    lib().orc_flow_force(_dp(F), _dp(U), _dims(dims))
    return F


def update_links(U, P, dims, eps, route=0):
    out = np.empty_like(U)
    lib().orc_update_links_route(_dp(out), _dp(U), _dp(P), _dims(dims), ctypes.c_double(eps), ctypes.c_int(route))
    return out


def update_momenta(P, U, dims, eps, beta):
    lib().orc_update_momenta(_dp(P), _dp(U), _dims(dims), ctypes.c_double(eps), ctypes.c_double(beta))
    return P


def md_step(U, P, dims, beta, eps, integrator=0):
    lib().orc_md_step(_dp(U), _dp(P), _dims(dims), ctypes.c_double(beta), ctypes.c_double(eps), ctypes.c_int(integrator))


def md_trajectory(U, P, dims, beta, steps, tau=1.0, integrator=0):
    H = np.zeros(2)
    lib().orc_md_trajectory(_dp(U), _dp(P), _dims(dims), ctypes.c_double(beta), ctypes.c_int(steps), ctypes.c_double(tau), ctypes.c_int(integrator), _dp(H))
    return H[0], H[1]


def flow_step(U, dims, eps):
    lib().orc_flow_step(_dp(U), _dims(dims), ctypes.c_double(eps))


def exp_ta(c8, t=1.0, route=0):
    c8 = np.ascontiguousarray(c8, dtype=np.float64)
    out = np.zeros((3, 3), dtype=np.complex128)
    lib().orc_exp_ta(_dp(c8), ctypes.c_double(t), ctypes.c_int(route), _dp(out))
    return out.T.copy()  # stored column-major -> math [i, j]


def ta_coeffs(m):
    buf = np.ascontiguousarray(np.asarray(m, dtype=np.complex128).T)
    c = np.zeros(8)
    lib().orc_ta_coeffs(_dp(buf), _dp(c))
    return c


def philox(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().orc_philox(c, k, o)
    return [int(v) for v in o]


def staple_sum(U, dims):
    V = np.empty_like(U)
    lib().orc_staple_sum(_dp(V), _dp(U), _dims(dims))
    return V


def stout_forward(U, dims, rho, want_q=False):
    out = np.empty_like(U)
    Q = new_p(dims) if want_q else None
    lib().orc_stout_forward(_dp(out), _dp(U), _dims(dims), ctypes.c_double(rho), _dp(Q) if want_q else None)
    return (out, Q) if want_q else out


def stout_backward(d_out, U, dims, rho, route=0):
    """back-prop through one stout layer: d_out = dS/dU' -> returns dS/dU (reference dSdU convention)."""
    d_in = np.empty_like(U)
    lib().orc_stout_backward(_dp(d_in), _dp(np.ascontiguousarray(d_out)), _dp(U), _dims(dims), ctypes.c_double(rho), ctypes.c_int(route))
    return d_in


def wilson_dSdU(U, dims, beta):
    """calc_dSdUmu! of the Wilson action pushed as beta/2*(plaq+plaq'): (beta/2) * sum of six staples."""
    D = np.empty_like(U)
    lib().orc_wilson_dSdU(_dp(D), _dp(U), _dims(dims), ctypes.c_double(beta))
    return D


def kick_from_dSdU(P, U, D, dims, factor):
    """P_mu += factor * TAcoeffs(U_mu * D_mu) (md_force! tail)."""
    lib().orc_kick_from_dSdU(_dp(P), _dp(U), _dp(np.ascontiguousarray(D)), _dims(dims), ctypes.c_double(factor))
    return P


def exp_pullback(C, Q, route=0):
    """L with tr(L dQ) = tr(C d exp(Q)); C, Q 3x3 in math indexing.  Returns (L, applied)."""
    c = np.ascontiguousarray(np.asarray(C, dtype=np.complex128).T)
    q = np.ascontiguousarray(np.asarray(Q, dtype=np.complex128).T)
    out = np.zeros((3, 3), dtype=np.complex128)
    lib().orc_exp_pullback.restype = ctypes.c_int
    ok = lib().orc_exp_pullback(_dp(c), _dp(q), ctypes.c_int(route), _dp(out))
    return out.T.copy(), bool(ok)


# ---- general-action path (plaquette + rectangle terms), loop sums, topological charge ------------------------------------
def force_general(U, dims, c_plaq, c_rect, scale=-1.0 / 3.0):
    """md_force! of the action c_plaq (plaq + plaq') + c_rect (rect + rect') by generic loop differentiation; scale = 1 gives
    F_update! of Gradientflow_general."""
    F = new_p(dims)
    lib().orc_force_general(_dp(F), _dp(U), _dims(dims), ctypes.c_double(c_plaq), ctypes.c_double(c_rect), ctypes.c_double(scale))
    return F


def loop_sums(U, dims):
    out = np.zeros(2)
    lib().orc_loop_sums(_dp(U), _dims(dims), _dp(out))
    return out[0], out[1]


def hamiltonian_general(U, P, dims, c_plaq, c_rect):
    sp, sr = loop_sums(U, dims)
    return -(2.0 / 3.0) * (c_plaq * sp + c_rect * sr) + 0.5 * momentum_norm2(P, dims)


def update_momenta_general(P, U, dims, eps, c_plaq, c_rect):
    P += eps * force_general(U, dims, c_plaq, c_rect)
    return P


def md_trajectory_general(U, P, dims, c_plaq, c_rect, steps, tau=1.0, integrator=0):
    """md_trajectory! with md_step! QPQ (0) / PQP (1) for the general action; U, P updated in place; returns (H0, H1)."""
    H0 = hamiltonian_general(U, P, dims, c_plaq, c_rect)
    eps = tau / steps
    for _ in range(steps):
        if integrator == 0:
            U[...] = update_links(U, P, dims, eps / 2)
            update_momenta_general(P, U, dims, eps, c_plaq, c_rect)
            U[...] = update_links(U, P, dims, eps / 2)
        else:
            update_momenta_general(P, U, dims, eps / 2, c_plaq, c_rect)
            U[...] = update_links(U, P, dims, eps)
            update_momenta_general(P, U, dims, eps / 2, c_plaq, c_rect)
    return H0, hamiltonian_general(U, P, dims, c_plaq, c_rect)


def flow_step_general(U, dims, eps, c_plaq, c_rect):
    """One RK3 step of flow!(U, ::Gradientflow_general) (src/smearing/gradientflow.jl:240-316), in place."""
    F0 = force_general(U, dims, c_plaq, c_rect, 1.0)
    W1 = update_links(U, -eps / 4 * F0, dims, 1.0)
    F1 = force_general(W1, dims, c_plaq, c_rect, 1.0)
    W2 = update_links(W1, -(8 * eps / 9) * F1 + (17 * eps / 36) * F0, dims, 1.0)
    F2 = force_general(W2, dims, c_plaq, c_rect, 1.0)
    U[...] = update_links(W2, -(3 * eps / 4) * F2 + (8 * eps / 9) * F1 - (17 * eps / 36) * F0, dims, 1.0)
    return U


def topological_charge_density(U, dims, method):
    """method 0 plaquette, 1 clover, 2 improved; array indexed [t, z, y, x]."""
    nx, ny, nz, nt = dims
    out = np.zeros((nt, nz, ny, nx))
    lib().orc_topological_charge_density(_dp(U), _dims(dims), ctypes.c_int(method), _dp(out))
    return out


def sexton_weingarten_trajectory(U, P, dims, terms, fast, slow, n_fast, steps, tau, ordering=0):
    """md_trajectory! with the SextonWeingarten integrator (src/molecular_dynamics.jl:618-700): `terms` maps a name to
    (c_plaq, c_rect); fast / slow are tuples of names.  In place; returns (H0, H1) of the full action."""
    def coeff(names):
        return sum(terms[n][0] for n in names), sum(terms[n][1] for n in names)

    cf, cs, ca = coeff(fast), coeff(slow), coeff(tuple(terms))

    def fast_qpq(duration):
        e = duration / n_fast
        U[...] = update_links(U, P, dims, e / 2)
        for k in range(n_fast):
            update_momenta_general(P, U, dims, e, *cf)
            U[...] = update_links(U, P, dims, e / 2 if k == n_fast - 1 else e)

    H0 = hamiltonian_general(U, P, dims, *ca)
    eps = tau / steps
    for _ in range(steps):
        if ordering == 0:
            fast_qpq(eps / 2)
            update_momenta_general(P, U, dims, eps, *cs)
            fast_qpq(eps / 2)
        else:
            update_momenta_general(P, U, dims, eps / 2, *cs)
            fast_qpq(eps)
            update_momenta_general(P, U, dims, eps / 2, *cs)
    return H0, hamiltonian_general(U, P, dims, *ca)


def heatbath_sweep(U, dims, beta, seed, sweep, overrelax=False):
    """One heatbath (or overrelaxation) sweep of the Wilson action in place; raises when a site update fails."""
    n = lib().orc_heatbath_sweep(_dp(U), _dims(dims), ctypes.c_double(beta), ctypes.c_uint64(seed), ctypes.c_uint64(sweep), ctypes.c_int(1 if overrelax else 0))
    if n:
        raise RuntimeError("heatbath failed at %d site(s)" % n)
    return U
