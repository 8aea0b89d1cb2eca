"""GPU tests of the primitive table (gfb_mul, gfb_axpy, gfb_tr, gfb_ta_project, gfb_exp, shift / adjoint views): the
reference's GENERIC algorithms, written with these primitives exactly as in the reference, must reproduce the fused kernels
and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DIMS = (4, 6, 4, 8)


def setup(backend, oracle, seed=17):
    import gfb200

    Uh = oracle.hot_start_philox(DIMS, seed)
    U = gfb200.gauge_configuration(DIMS, backend=backend).upload(Uh)
    return U, Uh, [gfb200.link_field(U, mu) for mu in range(4)]


def test_generic_plaquette_and_staple(backend, oracle):
    """calculate_Plaquette written as in src/AbstractGaugefields.jl:2684-2699 / construct_staple! :2873-2923."""
    import gfb200 as g

    U, Uh, L = setup(backend, oracle)
    t1, t2, V = L[0].similar(), L[0].similar(), L[0].similar()
    plaq = 0.0
    for mu in range(4):
        g.clear_U_(V)
        for nu in range(4):
            if nu == mu:
                continue
            # upper staple U_nu(x) U_mu(x+nu) U_nu(x+mu)^dag
            g.mul_(t1, L[nu], g.shift_U(L[mu], nu + 1))
            g.mul_(V, t1, g.shift_U(L[nu], mu + 1).H, 1.0, 1.0)
        g.mul_(t2, L[mu], V.H)
        plaq += g.tr(t2)
    want = oracle.plaquette_sum(Uh, DIMS)
    assert abs(plaq.real * 0.5 - want) < 1e-12 * abs(want) + 1e-12
    assert abs(plaq.real * 0.5 - g.calculate_Plaquette(U)) < 1e-11


def test_generic_force_matches_fused(backend, oracle):
    """md_force! composed from primitives (molecular_dynamics.jl:251-267): six staples with shifted / adjoint views,
    U*dSdU, Traceless_antihermitian_add! -- against the fused kernel and the oracle."""
    import gfb200 as g

    beta = 5.7
    U, Uh, L = setup(backend, oracle, 5)
    t1, t2, D = L[0].similar(), L[0].similar(), L[0].similar()
    F = g.gauge_momenta(U)
    for mu in range(4):
        g.clear_U_(D)
        for nu in range(4):
            if nu == mu:
                continue
            # derivative of tr(loop) w.r.t. U_mu: U_nu(x+mu) U_mu(x+nu)^dag U_nu(x)^dag  (= upper staple^dag)
            g.mul_(t1, g.shift_U(L[nu], mu + 1), g.shift_U(L[mu], nu + 1).H)
            g.mul_(D, t1, L[nu].H, beta / 2, 1.0)
            # lower: U_nu(x+mu-nu)^dag U_mu(x-nu)^dag U_nu(x-nu)
            sh = [0, 0, 0, 0]
            sh[mu] += 1
            sh[nu] -= 1
            g.mul_(t1, g.shift_U(L[nu], sh).H, g.shift_U(L[mu], -(nu + 1)).H)
            g.mul_(D, t1, g.shift_U(L[nu], -(nu + 1)), beta / 2, 1.0)
        g.mul_(t2, L[mu], D)
        g.Traceless_antihermitian_add_(F, mu, -1.0 / 3.0, t2)
    want = oracle.force(Uh, DIMS, beta)
    assert np.abs(F.to_host() - want).max() < 1e-12 * np.abs(want).max()
    # D of the last direction equals calc_dSdUmu (wilson_dSdU)
    assert np.abs(D.to_host() - oracle.wilson_dSdU(Uh, DIMS, beta)[3]).max() < 1e-12


def test_elementwise_ops(backend, oracle):
    import gfb200 as g

    U, Uh, L = setup(backend, oracle, 9)
    rng = np.random.default_rng(1)
    A = g.MatrixField(backend, DIMS)
    ah = rng.normal(size=A.host_shape()) + 1j * rng.normal(size=A.host_shape())
    A.upload(ah)
    assert np.array_equal(A.to_host(), ah)
    am = np.swapaxes(ah, -1, -2)  # math indexing
    # tr, tr(A,B)
    assert abs(g.tr(A) - np.trace(am, axis1=-2, axis2=-1).sum()) < 1e-10
    um = oracle.mats(Uh)
    assert abs(g.tr(A, L[2]) - np.trace(am @ um[2], axis1=-2, axis2=-1).sum()) < 1e-10
    # add_U! with adjoint, substitute with shift
    B = A.similar()
    g.substitute_U_(B, g.shift_U(A, (1, 0, -1, 2)))
    want = np.roll(ah, shift=(-2, 1, 0, -1), axis=(0, 1, 2, 3))
    assert np.array_equal(B.to_host(), want)
    g.add_U_(B, 0.5 - 2j, A.H)
    want_m = np.swapaxes(want, -1, -2) + (0.5 - 2j) * am.conj().swapaxes(-1, -2)
    assert np.abs(np.swapaxes(B.to_host(), -1, -2) - want_m).max() < 1e-13
    # Traceless_antihermitian!, exptU! (matrix form and from momenta)
    Q, E = A.similar(), A.similar()
    g.Traceless_antihermitian_(Q, A)
    q = (am - am.conj().swapaxes(-1, -2)) / 2
    q = q - np.trace(q, axis1=-2, axis2=-1)[..., None, None] / 3 * np.eye(3)
    assert np.abs(np.swapaxes(Q.to_host(), -1, -2) - q).max() < 1e-14
    small = A.similar()
    g.clear_U_(small)
    g.add_U_(small, 0.1, A)
    g.exptU_(E, 0.7, small)
    from scipy.linalg import expm

    e = np.swapaxes(E.to_host(), -1, -2)
    idx = (1, 2, 3, 0)
    assert np.abs(e[idx] - expm(0.7 * 0.1 * q[idx])).max() < 1e-13
    Ph = oracle.gaussian_momenta(DIMS, 3, 0)
    P = g.gauge_momenta(U).upload(Ph)
    g.exptU_(E, 0.3, P, mu=1)
    g.mul_(B, E, L[1])
    wantU = oracle.update_links(Uh, Ph, DIMS, 0.3)
    assert np.abs(B.to_host() - wantU[1]).max() < 1e-14
    g.unit_U_(E)
    assert np.array_equal(E.to_host()[0, 0, 0, 0], np.eye(3))
    with pytest.raises(ValueError):
        g.mul_(B, g.shift_U(B, 1), A)  # destination aliases a shifted operand
