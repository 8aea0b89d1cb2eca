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


# ---- argument TYPES, not just names: header <-> ctypes table <-> every ccall of the Julia glue ------------------------------
def header_signatures():
    """name -> (return class, [argument classes]) parsed from include/gfb200.h.  Classes: 'int', 'double', 'u64', 'size',
    'i64', 'ptr' (handle / void* / char* / double* / int*), 'pptr' (T**), 'str' (const char* return)."""
    text = open(os.path.join(ROOT, "include", "gfb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)

    def cls(t):
        t = " ".join(t.replace("const", " ").split())
        stars = t.count("*")
        base = t.replace("*", "").strip()
        if stars >= 2:
            return "pptr"
        if stars == 1:
            return "ptr"
        return {"int": "int", "double": "double", "uint64_t": "u64", "size_t": "u64", "long long": "i64", "void": "void"}[base]

    out = {}
    for ret, name, args in re.findall(r"^\s*((?:const\s+)?[A-Za-z_][A-Za-z0-9_ ]*?\s*\*?)\s*(gfb_[a-zA-Z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.M):
        alist = []
        for a in [x.strip() for x in args.replace("\n", " ").split(",")]:
            if a in ("void", ""):
                continue
            m = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)$", a)  # strip the parameter name
            alist.append(cls(m.group(1)))
        r = "str" if ("char" in ret and "*" in ret) else cls(ret)
        out[name] = (r, alist)
    return out


def test_ctypes_argument_types_match_header():
    from gfb200 import _lib

    def cls(t):
        if t is None:
            return "void"
        if t in (ctypes.c_int,):
            return "int"
        if t is ctypes.c_double:
            return "double"
        if t is ctypes.c_uint64:
            return "u64"
        if t is ctypes.c_size_t:  # the same ctypes object as c_uint64 on LP64: one class for both
            return "u64"
        if t is ctypes.c_longlong:
            return "i64"
        if t is ctypes.c_char_p:
            return "str"
        if t is ctypes.c_void_p:
            return "ptr"
        if hasattr(t, "_type_"):  # POINTER(x)
            return "pptr" if t._type_ is ctypes.c_void_p else "ptr"
        raise AssertionError("unmapped ctypes type %r" % (t,))

    hdr = header_signatures()
    assert sorted(hdr) == header_symbols()
    for name, (res, args) in _lib.SIGNATURES.items():
        hres, hargs = hdr[name]
        got = ["ptr" if cls(a) == "str" else cls(a) for a in args]  # a char* ARGUMENT is a pointer like any other
        assert got == hargs, "%s: ctypes %s vs header %s" % (name, got, hargs)
        if name in ("gfb_last_error",):
            assert hres == "str"
        else:
            assert cls(res) == hres, name


def julia_ccalls():
    """[(symbol, return type, [argument types], number of actual arguments)] of every ccall in B200Backend.jl."""
    text = open(os.path.join(ROOT, "gaugefields.jl_b200", "julia", "B200Backend.jl")).read()
    text = re.sub(r"#.*", "", text)
    out = []
    for m in re.finditer(r"ccall\(\(:(gfb_[a-z0-9_A-Z]+), LIBGFB200\),\s*([A-Za-z{}]+),\s*\(", text):
        i = m.end()
        depth, j = 1, i
        while depth:  # the type tuple
            depth += {"(": 1, ")": -1}.get(text[j], 0)
            j += 1
        types = [t.strip() for t in re.split(r",(?![^{]*})", text[i:j - 1]) if t.strip()]
        depth, k = 1, j  # the rest of the ccall: actual arguments up to its closing parenthesis
        while depth:
            depth += {"(": 1, ")": -1, "[": 1, "]": -1}.get(text[k], 0)
            k += 1
        rest = text[j:k - 1].strip().lstrip(",")
        nargs, d, cur = 0, 0, ""
        for ch in rest:
            if ch in "([{":
                d += 1
            elif ch in ")]}":
                d -= 1
            if ch == "," and d == 0:
                nargs += 1 if cur.strip() else 0
                cur = ""
            else:
                cur += ch
        nargs += 1 if cur.strip() else 0
        out.append((m.group(1), m.group(2), types, nargs))
    return out


def test_julia_ccalls_match_header():
    JL = {"Cint": "int", "Cdouble": "double", "UInt64": "u64", "Csize_t": "u64", "Clonglong": "i64", "Cstring": "str",
          "Ptr{Cvoid}": "ptr", "Ptr{Cdouble}": "ptr", "Ref{Cdouble}": "ptr", "Ptr{ComplexF64}": "ptr", "Ptr{Cint}": "ptr",
          "Ptr{UInt8}": "ptr", "Ref{Ptr{Cvoid}}": "pptr"}
    hdr = header_signatures()
    calls = julia_ccalls()
    assert len(calls) >= 55
    for name, ret, types, nargs in calls:
        assert name in hdr, "B200Backend.jl calls %s, which include/gfb200.h does not declare" % name
        hres, hargs = hdr[name]
        assert [JL[t] for t in types] == hargs, "%s: ccall types %s vs header %s" % (name, types, hargs)
        assert JL[ret] == hres, name
        assert nargs == len(types), "%s: %d arguments for %d declared types" % (name, nargs, len(types))
    # every helper the file calls is defined in it (round 1 shipped a call to an undefined _is_plaquette_pair)
    text = open(os.path.join(ROOT, "gaugefields.jl_b200", "julia", "B200Backend.jl")).read()
    code = re.sub(r"#.*", "", text)
    used = set(re.findall(r"\b(_[a-z0-9_]+)\(", code))
    defined = set(re.findall(r"^(?:function\s+)?(_[a-z0-9_]+)\(", code, flags=re.M)) | set(re.findall(r"^const\s+(_[A-Za-z0-9_]+)", code, flags=re.M))
    assert used <= defined, sorted(used - defined)
    # the hot-path entry points of SURVEY.md section 8b are all bound
    bound = {c[0] for c in calls}
    for need in ("gfb_md_trajectory_general", "gfb_update_momenta_general", "gfb_update_links", "gfb_hamiltonian_general", "gfb_flow", "gfb_flow_general",
                 "gfb_stout_forward", "gfb_stout_backward", "gfb_plaquette_sum", "gfb_gaussian_momenta", "gfb_set_hot", "gfb_gauge_copy", "gfb_mul",
                 "gfb_field_copy", "gfb_axpy", "gfb_tr", "gfb_ta_project", "gfb_ta_coeffs_add", "gfb_exp", "gfb_exp_mom", "gfb_field_alloc",
                 "gfb_field_view", "gfb_field_upload", "gfb_field_download", "gfb_reunitarize", "gfb_topological_charge", "gfb_mom_axpy_dir"):
        assert need in bound, need


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
        gfb200.make_loops_fromname("chair")
    rect = gfb200.make_loops_fromname("rectangular")
    sym = gfb200.GaugeAction(FakeU()).push(1.5, loops + loops.adjoint()).push(-0.1, rect + rect.adjoint())
    assert sym.coefficients() == (1.5, -0.1)
    with pytest.raises(NotImplementedError):
        sym.wilson_beta()  # Wilson-only call sites refuse an action with a rectangle term
    with pytest.raises(ValueError):
        gfb200.MDActionSet()
    with pytest.raises(ValueError):
        gfb200.MDForceGroup()
    with pytest.raises(ValueError):
        gfb200.SextonWeingarten(fast="a", slow="b", n_fast=0)
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
