"""GPU tests of the heatbath / overrelaxation updater (csrc/heatbath.cu) against the oracle's restatement of the reference's site
algorithm (src/heatbath/portable/kernels.jl; oracle/gf_oracle.cpp "Heatbath and overrelaxation"), on the same streams.

What can and cannot be pinned: the ALGORITHM (subgroup sequence, Kennedy-Pendleton acceptance, embedding, reunitarisation, sweep
order) is the reference's; the random BITS are this repository's (the reference's come from the un-vendored LatticeMatrices.jl),
so the comparison is CUDA <-> oracle plus the physics the reference's own tests check (plaquette after thermalisation,
test/heatbathtest.jl:150-200; overrelaxation leaves the action unchanged)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _unitarity_defect(Uh):
    m = Uh.reshape(-1, 3, 3)
    return np.abs(np.einsum("nij,nkj->nik", m, m.conj()) - np.eye(3)).max()


@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (8, 4, 6, 4)])
def test_heatbath_and_overrelaxation_match_oracle(backend, oracle, dims):
    import gfb200

    beta = 5.7
    Uh = oracle.hot_start_philox(dims, 4)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    h = gfb200.Heatbath(U, beta, seed=0x1234, sweep=3)
    gfb200.heatbath_(U, h)
    gfb200.heatbath_(U, h)
    oracle.heatbath_sweep(Uh, dims, beta, 0x1234, 3)
    oracle.heatbath_sweep(Uh, dims, beta, 0x1234, 4)
    got = U.to_host()
    # an accept/reject decision of the Kennedy-Pendleton loop is discrete: agreement is to rounding unless a candidate sits
    # within rounding of the acceptance boundary (probability ~1e-15 per decision)
    assert np.abs(got - Uh).max() < 1e-11
    assert _unitarity_defect(got) < 1e-14
    gfb200.overrelaxation_(U, h)
    oracle.heatbath_sweep(Uh, dims, beta, 0x1234, 3, overrelax=True)
    assert np.abs(U.to_host() - Uh).max() < 1e-11
    assert h.sweep == 5 and h.overrelaxation_sweep == 4


def test_overrelaxation_is_microcanonical(backend, oracle):
    import gfb200

    dims = (8, 8, 4, 4)
    Uh = oracle.hot_start_philox(dims, 9)
    for _ in range(3):
        oracle.flow_step(Uh, dims, 0.02)
    U = gfb200.gauge_configuration(dims, backend=backend).upload(Uh)
    s0 = gfb200.calculate_Plaquette(U)
    h = gfb200.Heatbath(U, 6.0, seed=5)
    for _ in range(3):
        gfb200.overrelaxation_(U, h)
    s1 = gfb200.calculate_Plaquette(U)
    assert abs(s1 - s0) < 1e-11 * abs(s0)              # the action is unchanged ...
    assert np.abs(U.to_host() - Uh).max() > 0.1         # ... by a large move
    assert _unitarity_defect(U.to_host()) < 1e-14


def test_heatbath_thermalises_to_the_wilson_plaquette(backend):
    """8^4, beta = 5.7 from a cold start: <P> = 0.549 in the thermodynamic limit (the value the reference's HMC and heatbath
    tests converge to, test/heatbathtest.jl, test/HMC_test.jl:110-113); 40 + 40 sweeps with one overrelaxation each."""
    import gfb200

    dims = (8, 8, 8, 8)
    U = gfb200.gauge_configuration(dims, backend=backend)
    h = gfb200.Heatbath(U, 5.7, seed=0xfeed)
    vals = []
    for sweep in range(80):
        gfb200.heatbath_(U, h)
        gfb200.overrelaxation_(U, h)
        if sweep >= 40:
            vals.append(gfb200.measure_plaquette(U))
    mean = float(np.mean(vals))
    assert 0.540 < mean < 0.560, mean


def test_heatbath_argument_checks(backend):
    import gfb200

    U = gfb200.gauge_configuration((4, 4, 4, 6), backend=backend)
    with pytest.raises(ValueError):
        gfb200.Heatbath(U, -1.0)
    V = gfb200.gauge_configuration((6, 5, 4, 4), backend=backend)
    with pytest.raises(ValueError):
        gfb200.heatbath_(V, gfb200.Heatbath(V, 5.7))  # odd extent: no checkerboard
