"""ctypes binding of libgfb200.so -- the stand-in for Julia's `ccall`.

There is deliberately no fallback: if the CUDA library is missing or no GPU is usable, every
entry point raises.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# GFB200_LIB selects a tuning variant built by build.py (default: libgfb200.so)
LIB_PATH = os.environ.get("GFB200_LIB") or os.path.join(_PKG, "libgfb200.so")

c_int, c_double, c_void_p, c_char_p = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_char_p
c_u64, c_size_t, c_ll = ctypes.c_uint64, ctypes.c_size_t, ctypes.c_longlong
P = ctypes.POINTER

# name -> (restype, argtypes); mirrors include/gfb200.h one to one
SIGNATURES = {
    "gfb_version": (c_int, []),
    "gfb_init": (c_int, [c_int, P(c_int), P(c_void_p)]),
    "gfb_nccl_unique_id": (c_int, [c_char_p]),
    "gfb_init_rank": (c_int, [c_int, c_int, c_char_p, c_int, P(c_void_p)]),
    "gfb_finalize": (c_int, [c_void_p]),
    "gfb_last_error": (c_char_p, [c_void_p]),
    "gfb_sync": (c_int, [c_void_p]),
    "gfb_num_slabs": (c_int, [c_void_p, P(c_int), P(c_int)]),
    "gfb_timer_tic": (c_int, [c_void_p]),
    "gfb_timer_toc": (c_int, [c_void_p, P(c_double)]),
    "gfb_kernel_launches": (c_int, [c_void_p, P(c_ll)]),
    "gfb_host_alloc": (c_int, [P(c_void_p), c_size_t]),
    "gfb_host_free": (c_int, [c_void_p]),
    "gfb_gauge_alloc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, P(c_void_p)]),
    "gfb_gauge_free": (c_int, [c_void_p]),
    "gfb_mom_alloc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, P(c_void_p)]),
    "gfb_mom_free": (c_int, [c_void_p]),
    "gfb_gauge_upload": (c_int, [c_void_p, c_int, c_void_p]),
    "gfb_gauge_download": (c_int, [c_void_p, c_int, c_void_p]),
    "gfb_gauge_upload_ildg": (c_int, [c_void_p, c_void_p, c_int]),
    "gfb_gauge_download_ildg": (c_int, [c_void_p, c_void_p, c_int]),
    "gfb_mom_upload": (c_int, [c_void_p, c_int, c_void_p]),
    "gfb_mom_download": (c_int, [c_void_p, c_int, c_void_p]),
    "gfb_gauge_copy": (c_int, [c_void_p, c_void_p]),
    "gfb_mom_copy": (c_int, [c_void_p, c_void_p]),
    "gfb_mom_zero": (c_int, [c_void_p]),
    "gfb_mom_axpy": (c_int, [c_void_p, c_double, c_void_p]),
    "gfb_mom_axpy_dir": (c_int, [c_void_p, c_int, c_double, c_void_p]),
    "gfb_set_cold": (c_int, [c_void_p]),
    "gfb_set_hot": (c_int, [c_void_p, c_u64, c_int]),
    "gfb_gaussian_momenta": (c_int, [c_void_p, c_u64, c_u64, c_double, c_int]),
    "gfb_reunitarize": (c_int, [c_void_p]),
    "gfb_plaquette_sum": (c_int, [c_void_p, P(c_double)]),
    "gfb_wilson_action": (c_int, [c_void_p, c_double, P(c_double)]),
    "gfb_kinetic": (c_int, [c_void_p, P(c_double)]),
    "gfb_hamiltonian": (c_int, [c_void_p, c_void_p, c_double, P(c_double)]),
    "gfb_energy_density": (c_int, [c_void_p, c_int, P(c_double)]),
    "gfb_polyakov": (c_int, [c_void_p, P(c_double)]),
    "gfb_force": (c_int, [c_void_p, c_void_p, c_double]),
    "gfb_update_momenta": (c_int, [c_void_p, c_void_p, c_double, c_double]),
    "gfb_update_links": (c_int, [c_void_p, c_void_p, c_double]),
    "gfb_md_trajectory": (c_int, [c_void_p, c_void_p, c_double, c_int, c_double, c_int, c_int, P(c_double)]),
    "gfb_flow": (c_int, [c_void_p, c_double, c_int]),
    "gfb_flow_force": (c_int, [c_void_p, c_void_p]),
    "gfb_exp_aF_U": (c_int, [c_void_p, c_double, c_void_p, c_void_p]),
    "gfb_stout_forward": (c_int, [c_void_p, c_void_p, c_double, c_void_p]),
    "gfb_stout_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_double]),
    "gfb_kick_from_dSdU": (c_int, [c_void_p, c_void_p, c_void_p, c_double]),
    "gfb_wilson_dSdU": (c_int, [c_void_p, c_void_p, c_double]),
    # general-action path and topological charge
    "gfb_loop_sums": (c_int, [c_void_p, P(c_double)]),
    "gfb_force_general": (c_int, [c_void_p, c_void_p, c_double, c_double]),
    "gfb_update_momenta_general": (c_int, [c_void_p, c_void_p, c_double, c_double, c_double]),
    "gfb_hamiltonian_general": (c_int, [c_void_p, c_void_p, c_double, c_double, P(c_double)]),
    "gfb_md_trajectory_general": (c_int, [c_void_p, c_void_p, c_double, c_double, c_int, c_double, c_int, c_int, P(c_double)]),
    "gfb_flow_general": (c_int, [c_void_p, c_double, c_int, c_double, c_double]),
    "gfb_topological_charge": (c_int, [c_void_p, c_int, P(c_double)]),
    "gfb_topological_charge_density": (c_int, [c_void_p, c_int, c_void_p]),
    "gfb_heatbath": (c_int, [c_void_p, c_double, c_u64, c_u64, c_int]),
    "gfb_overrelaxation": (c_int, [c_void_p, c_double, c_u64, c_u64, c_int]),
    # primitive table
    "gfb_field_alloc": (c_int, [c_void_p, c_int, c_int, c_int, c_int, P(c_void_p)]),
    "gfb_field_view": (c_int, [c_void_p, c_int, P(c_void_p)]),
    "gfb_field_free": (c_int, [c_void_p]),
    "gfb_field_upload": (c_int, [c_void_p, c_void_p]),
    "gfb_field_download": (c_int, [c_void_p, c_void_p]),
    "gfb_field_clear": (c_int, [c_void_p]),
    "gfb_field_unit": (c_int, [c_void_p]),
    "gfb_field_copy": (c_int, [c_void_p, c_void_p, P(c_int), c_int]),
    "gfb_mul": (c_int, [c_void_p, c_void_p, P(c_int), c_int, c_void_p, P(c_int), c_int, c_double, c_double, c_double, c_double]),
    "gfb_axpy": (c_int, [c_void_p, c_double, c_double, c_void_p, c_int]),
    "gfb_tr": (c_int, [c_void_p, P(c_double)]),
    "gfb_tr2": (c_int, [c_void_p, c_void_p, P(c_double)]),
    "gfb_ta_project": (c_int, [c_void_p, c_void_p]),
    "gfb_ta_coeffs_add": (c_int, [c_void_p, c_int, c_double, c_void_p]),
    "gfb_exp": (c_int, [c_void_p, c_double, c_void_p]),
    "gfb_exp_mom": (c_int, [c_void_p, c_double, c_void_p, c_int]),
}

_lib = None


class GfbError(RuntimeError):
    """A non-zero status from libgfb200 (the Julia glue throws ErrorException here)."""


def _preload_bundled_nccl():
    """libgfb200.so needs `libnccl.so.2`; so does PyTorch, which ships a NEWER one (site-packages/nvidia/nccl/lib) than the
    system library the dynamic loader would pick for us.  Whichever copy is mapped first serves both, and `import torch`
    fails on the older one (undefined symbol), so map the bundled copy first when it exists.  Harmless if torch is absent."""
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                ctypes.CDLL(cand, mode=ctypes.RTLD_GLOBAL)
                return
    except Exception:  # noqa: BLE001 -- fall back to the loader's choice
        pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GfbError(
            "libgfb200.so is not built (%s). Run `python gaugefields.jl_b200/build.py`; "
            "there is no CPU fallback." % LIB_PATH
        )
    _preload_bundled_nccl()
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, ctx=None):
    if status == 0:
        return
    msg = load().gfb_last_error(ctx)
    text = msg.decode() if msg else "unknown error"
    if status == 1:
        raise ValueError(text)  # ArgumentError in the reference (molecular_dynamics.jl:447-465)
    raise GfbError("libgfb200 status %d: %s" % (status, text))
