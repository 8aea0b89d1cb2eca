"""ILDG at the boundary: the LIME container on the host (CPU tests, against the reference's own fixture
files when /root/reference is present) and the payload <-> device path on the GPU (byte swap + transpose kernels)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gaugefields.jl_b200"))
REF_FIXTURES = "/root/reference/test/data"


def payload_of(Uh, precision):
    """numpy restatement of _save_binarydata (src/output/ildg_format.jl:697-746): [t][z][y][x][mu][row][col], big-endian.
    The gathered host array stores element (row, col) of a site at [..., col, row] (Julia column-major, src/API.jl:516-529)."""
    a = np.transpose(Uh, (1, 2, 3, 4, 0, 6, 5))
    return np.ascontiguousarray(a).astype(">c16" if precision == 64 else ">c8").tobytes()


def test_lime_container_round_trip(tmp_path):
    from gfb200 import ildg

    lattice = (4, 2, 6, 2)
    rng = np.random.default_rng(1)
    payload = rng.integers(0, 256, size=4 * 2 * 6 * 2 * 4 * 9 * 16, dtype=np.uint8).tobytes()
    fn = str(tmp_path / "c.ildg")
    ildg.write_ildg(fn, lattice, 64, payload)
    got = ildg.read_ildg(fn)
    assert got[0] == lattice and got[1] == 64 and got[2] == payload and got[3] == 3
    recs = ildg.read_records(fn)
    assert [r[0] for r in recs] == ["ildg-format", "ildg-binary-data"]
    assert os.path.getsize(fn) % 8 == 0


def test_bridge_text_is_the_ildg_order_in_decimal(tmp_path):
    """save_textdata / load_BridgeText! (src/output/bridge_format.jl:201-297) use the ILDG element order, one number per line."""
    from gfb200 import ildg

    lattice = (2, 2, 2, 2)
    rng = np.random.default_rng(3)
    Uh = rng.normal(size=(4, 2, 2, 2, 2, 3, 3)) + 1j * rng.normal(size=(4, 2, 2, 2, 2, 3, 3))
    payload = payload_of(Uh, 64)
    fn = str(tmp_path / "conf.txt")
    ildg.write_bridge_text(fn, payload)
    lines = open(fn).read().split()
    assert len(lines) == 4 * 16 * 9 * 2
    # first link of the file: site (1,1,1,1), mu = 1, element (a=1,b=1) real, imaginary; then (a=1,b=2)
    assert float(lines[0]) == Uh[0, 0, 0, 0, 0, 0, 0].real and float(lines[1]) == Uh[0, 0, 0, 0, 0, 0, 0].imag
    assert float(lines[2]) == Uh[0, 0, 0, 0, 0, 1, 0].real  # U[a=1,b=2]: row 1, col 2 -> gathered array index [col, row]
    assert ildg.read_bridge_text(fn, lattice) == payload
    with pytest.raises(ValueError):
        ildg.read_bridge_text(fn, (2, 2, 2, 4))


@pytest.mark.parametrize("name", ["conf_00000100_4444nc2.ildg", "conf_00000100_4444_test.ildg"])
def test_reads_the_reference_fixture_files(name):
    """The reference ships two 4^4 ILDG files (test/data); they are SU(2), so only the container and the size logic apply."""
    from gfb200 import ildg

    path = os.path.join(REF_FIXTURES, name)
    if not os.path.exists(path):
        pytest.skip("reference fixtures not present on this machine")
    assert [r[0] for r in ildg.read_records(path)] == ["ildg-binary-data"]  # bare payload record, no ildg-format
    with pytest.raises(ValueError):
        ildg.read_ildg(path)
    lattice, precision, payload, nc = ildg.read_ildg(path, lattice=(4, 4, 4, 4), precision=64)
    assert lattice == (4, 4, 4, 4) and precision in (32, 64) and nc == 2
    assert len(payload) == 256 * 4 * nc * nc * 2 * (precision // 8)
    # links of a thermalised SU(2) configuration are unitary: a direct check that the payload order is [..][mu][row][col] complex
    u = np.frombuffer(payload, dtype=">c16" if precision == 64 else ">c8").reshape(4, 4, 4, 4, 4, nc, nc)
    eye = np.einsum("...ab,...cb->...ac", u, u.conj())
    assert np.abs(eye - np.eye(nc)).max() < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (8, 4, 2, 6)])
def test_ildg_payload_to_device_and_back(backend, oracle, dims, tmp_path):
    import gfb200
    from gfb200 import ildg

    Uh = oracle.hot_start_philox(dims, 77)
    U = gfb200.gauge_configuration(dims, backend=backend)
    p64 = payload_of(Uh, 64)
    U.upload_ildg(p64, 64)
    assert np.array_equal(U.to_host(), Uh)            # byte swap + transpose are exact
    assert U.to_ildg(64) == p64
    # through a file, single precision (what production ensembles use)
    fn = str(tmp_path / "conf.ildg")
    ildg.write_ildg(fn, dims, 32, U.to_ildg(32))
    lattice, precision, payload, nc = ildg.read_ildg(fn)
    assert lattice == dims and precision == 32 and nc == 3
    assert payload == payload_of(Uh, 32)
    V = gfb200.gauge_configuration(dims, backend=backend).upload_ildg(payload, 32)
    assert np.abs(V.to_host() - Uh).max() < 1e-7
    assert abs(gfb200.measure_plaquette(V) - oracle.plaquette(Uh, dims)) < 1e-6
    with pytest.raises(ValueError):
        U.upload_ildg(p64[:-16], 64)
    with pytest.raises(ValueError):
        U.upload_ildg(p64, 16)
