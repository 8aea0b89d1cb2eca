"""CPU tests: the oracle against the reference's own golden values and invariants.

These pin the checker before it is trusted to judge the CUDA path.
"""
import numpy as np
import pytest

DIMS = (4, 4, 4, 4)


def test_philox_known_answers(oracle):
    """Random123 known-answer vectors for philox4x32-10."""
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_golden_hot_start_plaquette(oracle):
    """Legacy reproducible hot start, 4^4 SU(3): 0.008449494077606137 @1e-8 (test/init.jl:276-283)."""
    U = oracle.hot_start_stable123(DIMS)
    val = 0.008449494077606137
    assert abs(oracle.plaquette(U, DIMS) - val) / abs(val) < 1e-8
    assert abs(oracle.plaquette(U, DIMS) - val) < 2e-17 * 1e3
    # Appendix B quirk: every direction re-seeds StableRNG(123) -> identical fields
    assert np.array_equal(U[0], U[3])


def test_golden_flow_plaquette(oracle):
    """4^4 SU(3) Wilson flow 100 x eps=0.01 from that start: 0.8786515255315753 (test/gradientflow_test.jl:129-139)."""
    U = oracle.hot_start_stable123(DIMS)
    for _ in range(100):
        oracle.flow_step(U, DIMS, 0.01)
    val = 0.8786515255315753
    assert abs(oracle.plaquette(U, DIMS) - val) / val < 1e-1  # the reference's tolerance
    assert abs(oracle.plaquette(U, DIMS) - val) < 1e-12  # what the restatement actually achieves


def test_cold_start(oracle):
    """cold start: plaquette exactly 1, p*p 0, Delta H 0 (test/init.jl:190-211, test/md_driver.jl:371-395)."""
    U = oracle.set_cold(DIMS)
    assert oracle.plaquette(U, DIMS) == 1.0
    P = oracle.new_p(DIMS)
    H0, H1 = oracle.md_trajectory(U, P, DIMS, 5.7, 5, 1.0, 0)
    assert H1 - H0 == 0.0 and oracle.momentum_norm2(P, DIMS) == 0.0
    assert oracle.plaquette(U, DIMS) == 1.0


def su2_instanton_link(mu, site, L, sign=+1):
    """_su2_instanton_link (src/AbstractGaugefields.jl:1126-1170), 1-based site."""
    from scipy.linalg import expm

    center = [l / 2 + 0.5 for l in L]
    radius = L[0] // 2
    s1 = np.array([[0, 1], [1, 0]], dtype=complex)
    s2 = np.array([[0, -1j], [1j, 0]], dtype=complex)
    s3 = np.array([[1, 0], [0, -1]], dtype=complex)
    e = np.eye(2, dtype=complex)
    ss = [1j * s1, 1j * s2, 1j * s3, e]
    sd = [-1j * s1, -1j * s2, -1j * s3, e]
    nv = np.array([site[k] - 1 - center[k] for k in range(4)], dtype=complex)
    n2 = float(np.real(np.vdot(nv, nv)))
    tau = np.zeros((2, 2), dtype=complex)
    for nu in range(4):
        smunu = sd[mu] @ ss[nu] - sd[nu] @ ss[mu] if sign == +1 else ss[mu] @ sd[nu] - ss[nu] @ sd[mu]
        tau += smunu * nv[nu]
    return expm(1j * tau * 0.5 * (1 / n2) * (1j * radius**2 / (n2 + radius**2)))


def build_instanton_su3(oracle, L):
    """SU(2) instanton links embedded in the (1,2) block of SU(3) (Oneinstanton_SUN_embedded,
    src/AbstractGaugefields.jl:1495-1587); tr_3 = tr_2 + 1."""
    nx, ny, nz, nt = L
    U = oracle.new_u(L)
    m = oracle.mats(U)
    for t in range(nt):
        for z in range(nz):
            for y in range(ny):
                for x in range(nx):
                    for mu in range(4):
                        link = np.eye(3, dtype=complex)
                        link[:2, :2] = su2_instanton_link(mu, (x + 1, y + 1, z + 1, t + 1), L)
                        m[mu, t, z, y, x] = link
    return U


def test_one_instanton_plaquette(oracle):
    """SU(2) one-instanton plaquette 0.9796864531099871 @1e-8 (test/init.jl:351-371).

    The reference test calls Oneinstanton(NC, NX, NY, NZ, NT, Nwing) against the signature
    Oneinstanton(NC, NDW, NN...), so the lattice it really builds is NX x NY x NZ x NT = 4 x 4 x 4 x 1
    (NDW = 4); on that lattice the golden value is reproduced to 16 digits.  Four distinct link
    directions make this the pin of the plaquette geometry (the hot-start golden value has U_1 = ... = U_4)."""
    L = (4, 4, 4, 1)
    U = build_instanton_su3(oracle, L)
    p3 = oracle.plaquette(U, L)
    p2 = (3.0 * p3 - 1.0) / 2.0
    val = 0.9796864531099871
    assert abs(p2 - val) / val < 1e-8
    assert abs(p2 - val) < 1e-14


@pytest.mark.parametrize("integrator", [0, 1])
def test_md_reversibility(oracle, integrator):
    """forward + backward trajectory restores U, P to < 2e-12 (test/md_driver.jl:417-482)."""
    U0 = oracle.hot_start_philox(DIMS, 1)
    P0 = oracle.gaussian_momenta(DIMS, 2, 0)
    U, P = U0.copy(), P0.copy()
    oracle.md_trajectory(U, P, DIMS, 5.7, 8, 0.4, integrator)
    oracle.md_trajectory(U, P, DIMS, 5.7, 8, -0.4, integrator)
    assert np.abs(U - U0).max() < 2e-12 and np.abs(P - P0).max() < 2e-12


def test_force_is_minus_gradient_of_potential(oracle):
    """finite differences: dV/dt along U -> exp(t X) U equals -sum_a X_a F_a (consistency of md_force!,
    md_potential and exptU! conventions, src/molecular_dynamics.jl:247-267)."""
    dims = DIMS
    beta = 5.7
    U = oracle.hot_start_philox(dims, 3)
    F = oracle.force(U, dims, beta)
    rng = np.random.default_rng(0)
    X = rng.normal(size=oracle.p_shape(dims))
    h = 1e-5
    Vp = -(beta / 3.0) * oracle.plaquette_sum(oracle.update_links(U, X, dims, +h), dims)
    Vm = -(beta / 3.0) * oracle.plaquette_sum(oracle.update_links(U, X, dims, -h), dims)
    fd = (Vp - Vm) / (2 * h)
    an = -np.sum(X * F)
    assert abs(fd - an) / abs(an) < 1e-8


def test_force_additivity(oracle):
    """split-action force equals the total force (test/md_driver.jl:85-118): F(beta1)+F(beta2) == F(beta1+beta2)."""
    U = oracle.hot_start_philox(DIMS, 4)
    assert np.abs(oracle.force(U, DIMS, 2.0) + oracle.force(U, DIMS, 3.7) - oracle.force(U, DIMS, 5.7)).max() < 2e-12


def test_delta_h_scales_as_step_squared(oracle):
    U0 = oracle.hot_start_philox(DIMS, 5)
    for _ in range(5):
        oracle.flow_step(U0, DIMS, 0.02)
    P0 = oracle.gaussian_momenta(DIMS, 6, 0)
    dh = []
    for steps in (10, 20, 40):
        U, P = U0.copy(), P0.copy()
        H0, H1 = oracle.md_trajectory(U, P, DIMS, 5.7, steps, 1.0, 0)
        dh.append(H1 - H0)
    assert 3.0 < dh[0] / dh[1] < 5.0 and 3.0 < dh[1] / dh[2] < 5.0


def test_exp_routes_agree(oracle):
    """eigen-decomposition route (the legacy algorithm, TA_gaugefields_4D_serial.jl:850-1074) vs Taylor: 1e-14."""
    rng = np.random.default_rng(1)
    for scale in (1e-8, 1e-3, 0.1, 1.0, 5.0):
        for _ in range(20):
            c = rng.normal(size=8) * scale
            a, b = oracle.exp_ta(c, 1.0, 0), oracle.exp_ta(c, 1.0, 1)
            assert np.abs(a - b).max() < 1e-14 * max(1.0, scale)
            assert np.abs(a @ a.conj().T - np.eye(3)).max() < 1e-14 * max(1.0, scale)
    assert np.array_equal(oracle.exp_ta(np.zeros(8), 1.0, 0), np.eye(3))


def test_ta_projection_roundtrip(oracle):
    """TA(M) = sum_a c_a i lambda_a/2 with the Gell-Mann ordering of TA_gaugefields_4D_serial.jl:181-269."""
    lam = np.zeros((8, 3, 3), dtype=complex)
    lam[0][0, 1] = lam[0][1, 0] = 1
    lam[1][0, 1], lam[1][1, 0] = -1j, 1j
    lam[2][0, 0], lam[2][1, 1] = 1, -1
    lam[3][0, 2] = lam[3][2, 0] = 1
    lam[4][0, 2], lam[4][2, 0] = -1j, 1j
    lam[5][1, 2] = lam[5][2, 1] = 1
    lam[6][1, 2], lam[6][2, 1] = -1j, 1j
    lam[7] = np.diag([1, 1, -2]) / np.sqrt(3)
    rng = np.random.default_rng(2)
    c = rng.normal(size=8)
    m = sum(c[a] * 1j * lam[a] / 2 for a in range(8))
    assert np.abs(oracle.ta_coeffs(m) - c).max() < 1e-15
    g = rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3))
    ta = 0.5 * (g - g.conj().T)
    ta -= np.trace(ta) / 3 * np.eye(3)
    rec = sum(oracle.ta_coeffs(g)[a] * 1j * lam[a] / 2 for a in range(8))
    assert np.abs(rec - ta).max() < 1e-15
    # exptU! convention: exp(t * sum c_a i lambda_a/2)
    from scipy.linalg import expm

    assert np.abs(oracle.exp_ta(c, 0.3, 0) - expm(0.3 * m)).max() < 1e-14


def test_gaussian_stream_structure(oracle):
    """(value, spare) pairs per site stream, sweep/direction/seed separation, decomposition independence
    (test/MPIJACCtest/random_fields_site_rng.jl:22-42, 148-172, 225-274)."""
    dims = (4, 2, 2, 2)
    a = oracle.gaussian_momenta(dims, 300, 4, 1.5)
    assert np.array_equal(a, oracle.gaussian_momenta(dims, 300, 4, 1.5))
    assert not np.array_equal(a, oracle.gaussian_momenta(dims, 300, 5, 1.5))
    assert not np.array_equal(a, oracle.gaussian_momenta(dims, 301, 4, 1.5))
    assert not np.array_equal(a[0], a[3])
    # a larger lattice in t shares the streams of its first sites (global-site keyed, no rank/local index)
    b = oracle.gaussian_momenta((4, 2, 2, 4), 300, 4, 1.5)
    assert np.array_equal(b[:, :2], a)
    big = oracle.gaussian_momenta((8, 8, 8, 8), 1, 0, 1.0)
    assert abs(big.mean()) < 0.01 and abs(big.std() - 1.0) < 0.01


def test_hot_start_unitarity(oracle):
    U = oracle.hot_start_philox((4, 6, 2, 4), 0x6A09E667F3BCC909)
    m = oracle.mats(U)
    assert np.abs(m @ m.conj().swapaxes(-1, -2) - np.eye(3)).max() < 2e-12
    assert np.abs(np.linalg.det(m) - 1).max() < 2e-12
    assert not np.array_equal(U[0], U[1])


def test_stout_forward_properties(oracle):
    """rho = 0 is the identity map; smearing raises the plaquette; output stays in SU(3)."""
    U = oracle.hot_start_philox(DIMS, 8)
    assert np.abs(oracle.stout_forward(U, DIMS, 0.0) - U).max() < 1e-15
    V = oracle.stout_forward(U, DIMS, 0.1)
    assert oracle.plaquette(V, DIMS) > oracle.plaquette(U, DIMS)
    m = oracle.mats(V)
    assert np.abs(m @ m.conj().swapaxes(-1, -2) - np.eye(3)).max() < 1e-13


# ------------------------------------------------------------------------------------------------
# stout backward (src/smearing/stout_fast.jl:317-407, 712-785, 888-946)
# ------------------------------------------------------------------------------------------------
def _rand_ta(rng, scale):
    m = rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3))
    q = (m - m.conj().T) / 2
    q -= np.trace(q) / 3 * np.eye(3)
    return q * scale


def test_exp_pullback_series_vs_closed_form_and_fd(oracle):
    """CdexpQdQ!: the reference's closed form (AbstractGaugefields.jl:3284-3343) against the independent series
    route (< 1e-11, the bar of test/latticematrices_compat.jl:440-465) and against finite differences of expm."""
    from scipy.linalg import expm

    rng = np.random.default_rng(3)
    for scale in (1e-4, 0.05, 0.3, 1.0):
        for _ in range(6):
            Q = _rand_ta(rng, scale)
            C = rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3))
            L0, _ = oracle.exp_pullback(C, Q, 0)
            L1, ok = oracle.exp_pullback(C, Q, 1)
            assert ok
            # the closed form is derived with the traceless Cayley-Hamilton relation, so it may differ from the
            # unconstrained derivative by a multiple of the identity, which no traceless dQ can see (and which
            # calc_dSdΩ!'s projection removes): compare the traceless parts
            tl = lambda m: m - np.trace(m) / 3 * np.eye(3)
            assert np.abs(tl(L0) - tl(L1)).max() < 1e-11 * max(1.0, np.abs(L0).max())
            # tr(L dQ) == tr(C d exp(Q)) for a random direction (central difference)
            dQ = _rand_ta(rng, 1.0)
            h = 1e-5
            fd = np.trace(C @ (expm(Q + h * dQ) - expm(Q - h * dQ))) / (2 * h)
            assert abs(np.trace(L0 @ dQ) - fd) < 1e-8 * max(1.0, abs(fd))
    # the reference skips (leaves the output untouched) below |tr Q^2| = 1e-18
    _, ok = oracle.exp_pullback(np.eye(3), np.zeros((3, 3)), 1)
    assert not ok
    L0, _ = oracle.exp_pullback(np.eye(3) * 2.0, np.zeros((3, 3)), 0)
    assert np.allclose(L0, np.eye(3) * 2.0)


def test_stout_backward_against_finite_differences(oracle):
    """dS/dU through two stout layers: the chain-rule result, fed to the md_force! tail, must equal the
    directional derivative of S(U) = -(beta/3) sum Re tr P[stout(stout(U))] (pattern: stout_fast.jl:787-886)."""
    dims = (4, 4, 2, 2)
    beta, rho = 5.7, 0.1
    U = oracle.hot_start_philox(dims, 21)
    for _ in range(2):
        oracle.flow_step(U, dims, 0.02)

    def action(Uin):
        V = oracle.stout_forward(oracle.stout_forward(Uin, dims, rho), dims, rho)
        return -(beta / 3.0) * oracle.plaquette_sum(V, dims)

    U1 = oracle.stout_forward(U, dims, rho)
    U2 = oracle.stout_forward(U1, dims, rho)
    d2 = oracle.wilson_dSdU(U2, dims, beta)
    for route in (0, 1):
        d1 = oracle.stout_backward(d2, U1, dims, rho, route)
        d0 = oracle.stout_backward(d1, U, dims, rho, route)
        F = oracle.kick_from_dSdU(oracle.new_p(dims), U, d0, dims, -1.0 / 3.0)  # force = -(1/NC) TA(U dSdU)
        # d/ds S(exp(s X) U)|_0 with X = sum_a x_a i lambda_a/2 equals -sum_a x_a F_a (dP/dt = F = -dS/dX)
        rng = np.random.default_rng(9)
        X = rng.normal(size=oracle.p_shape(dims))
        h = 1e-5
        Sp = action(oracle.update_links(U, X, dims, +h))
        Sm = action(oracle.update_links(U, X, dims, -h))
        fd = (Sp - Sm) / (2 * h)
        an = -float(np.sum(X * F))
        assert abs(fd - an) < 2e-7 * max(1.0, abs(an)), (route, fd, an)
    # with rho = 0 the layer is the identity and back-prop returns its input
    d = oracle.stout_backward(d2, U, dims, 0.0, 0)
    assert np.abs(d - d2).max() < 1e-14
    # unsmeared consistency: kick from wilson_dSdU == md_force!
    F0 = oracle.kick_from_dSdU(oracle.new_p(dims), U, oracle.wilson_dSdU(U, dims, beta), dims, -1.0 / 3.0)
    assert np.abs(F0 - oracle.force(U, dims, beta)).max() < 1e-13


# ---- general-action oracle (plaquette + rectangle terms, topological charge): no reference golden exists for these on the
# ---- hot path, so the restatement is pinned by identities the reference's definitions imply
def test_general_force_reduces_to_wilson_and_matches_the_action_derivative(oracle):
    dims = (4, 4, 4, 6)
    U = oracle.hot_start_philox(dims, 3)
    assert [oracle.lib().orc_rotations_through(0, m) for m in range(4)] == [6] * 4      # 6 plaquette staples per link
    assert [oracle.lib().orc_rotations_through(1, m) for m in range(4)] == [18] * 4     # 18 rectangle staples per link
    assert np.abs(oracle.force_general(U, dims, 2.9, 0.0) - oracle.force(U, dims, 5.8)).max() < 1e-14
    sp, sr = oracle.loop_sums(U, dims)
    assert abs(sp - oracle.plaquette_sum(U, dims)) < 1e-10
    cold = oracle.set_cold(dims)
    assert oracle.loop_sums(cold, dims) == (18.0 * 384, 36.0 * 384)
    # dS/dtau along U(tau) = exp(tau P) U equals -P.F (energy conservation to first order), S = -(2/NC)(cp Sp + cr Sr)
    P = oracle.gaussian_momenta(dims, 5, 0)
    cp, cr, h = 1.7, -0.35, 1e-5

    def S(V):
        a, b = oracle.loop_sums(V, dims)
        return -(2.0 / 3.0) * (cp * a + cr * b)

    dS = (S(oracle.update_links(U, P, dims, h)) - S(oracle.update_links(U, P, dims, -h))) / (2 * h)
    F = oracle.force_general(U, dims, cp, cr)
    assert abs(dS + (P * F).sum()) < 1e-7 * abs(dS)


def test_topological_charge_oracle_invariants(oracle):
    dims = (4, 4, 4, 4)
    cold = oracle.set_cold(dims)
    for m in range(3):
        assert np.abs(oracle.topological_charge_density(cold, dims, m)).max() == 0.0
    U = oracle.hot_start_philox(dims, 8)
    for _ in range(4):
        oracle.flow_step(U, dims, 0.02)
    qc, qr = oracle.topological_charge_density(U, dims, 1), oracle.topological_charge_density(U, dims, 2)
    assert qc.shape == (4, 4, 4, 4) and np.isfinite(qr).all()
    # reflecting the x axis (U_x(x) -> U_x(-x-1)^dagger, other links mirrored) flips the sign of every definition
    V = np.empty_like(U)
    V[0] = np.conj(np.swapaxes(U[0][:, :, :, ::-1], -1, -2))      # V_x(x) = U_x(-x-1)^dagger
    for mu in (1, 2, 3):
        V[mu] = np.roll(U[mu][:, :, :, ::-1], 1, axis=3)          # V_nu(x) = U_nu(-x)
    # (the one-corner plaquette definition pairs F_x,nu and F_rho,sigma of different corners after the reflection: not odd)
    for m in (1, 2):
        q, qm = oracle.topological_charge_density(U, dims, m).sum(), oracle.topological_charge_density(V, dims, m).sum()
        assert abs(q + qm) < 1e-12 * max(1.0, abs(q))
