"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/gfb200.h declares,
the ctypes table mirrors the header, and the host-side validation matches the reference's error behaviour.
No compute call is made (there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gfb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gfb_[a-zA-Z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import gfb200
    from gfb200 import _lib

    _lib._preload_bundled_nccl()  # same NCCL copy as a later `import torch` in this process (see _lib.load)
    lib = ctypes.CDLL(gfb200.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 40
    for name in syms:
        assert hasattr(lib, name), "libgfb200.so does not export %s" % name


def test_ctypes_table_matches_header():
    from gfb200 import _lib

    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the context cannot be created: the product path fails loudly."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import gfb200

    with pytest.raises(gfb200.GfbError) as e:
        gfb200.B200Backend(ngpu=1)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    """Nothing under gaugefields.jl_b200/ may reference the oracle."""
    pkg = os.path.join(ROOT, "gaugefields.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "gf_oracle" not in text and "oracle/" not in text and "libgforacle" not in text, f


def test_host_validation_mirrors_reference():
    """Argument checks that do not need a device (src/API.jl:196-201, src/molecular_dynamics.jl:447-465)."""
    import gfb200

    class FakeU:
        lattice = (4, 4, 4, 4)

    action = gfb200.GaugeAction(FakeU())
    loops = gfb200.make_loops_fromname("plaquette")
    with pytest.raises(ValueError):
        action.wilson_beta()  # empty action
    action.push(5.7 / 2, loops + loops.adjoint())
    assert abs(action.wilson_beta() - 5.7) < 1e-15
    bad = gfb200.GaugeAction(FakeU()).push(1.0, loops)  # missing the conjugate loops
    with pytest.raises(NotImplementedError):
        bad.wilson_beta()
    with pytest.raises(NotImplementedError):
        gfb200.make_loops_fromname("rectangular")
    with pytest.raises(ValueError):
        gfb200.gradient_flow(FakeU(), steps=0)
    with pytest.raises(ValueError):
        gfb200.gradient_flow(FakeU(), steps=1, step_size=-0.1)
    with pytest.raises(ValueError):
        gfb200.gauge_configuration((4, 4, 4), backend=object())
    with pytest.raises(ValueError):
        gfb200.gauge_configuration((4, 4, 4, 4), backend=object(), colors=2)
    with pytest.raises(ValueError):
        gfb200.gauge_configuration((4, 4, 4, 4), backend=object(), start="warm")
    with pytest.raises(ValueError):
        gfb200.gaussian_momenta_(None, sweep=-1)
