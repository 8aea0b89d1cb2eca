"""Host-side mirror of the Gaugefields.jl high-level API for the B200 backend.

Julia is not available in the build image, so this module plays the role of the Julia glue
(`gaugefields.jl_b200/julia/B200Backend.jl`): the same entry points, argument meaning and error
behaviour as the reference (`src/API.jl`, `src/molecular_dynamics.jl`,
`src/smearing/gradientflow.jl`, `src/smearing/stout_fast.jl`), bound to libgfb200.so through
ctypes exactly where Julia would `ccall`.  Julia's mutating `name!` is spelled `name_` here.

Host arrays use the reference's gathered layout (src/API.jl:516-529):
    links   : numpy complex128, shape (4, NT, NZ, NY, NX, 3, 3), last two axes (column, row)
              == one Julia `ComplexF64[3,3,NX,NY,NZ,NT]` per direction
    momenta : numpy float64,   shape (4, NT, NZ, NY, NX, 8) == `Float64[8,1,NX,NY,NZ,NT]`
"""
import ctypes
import math
import os

import numpy as np

from . import _lib
from ._lib import GfbError, check

__all__ = [
    "B200Backend", "GaugeConfiguration", "Momenta", "GaugeAction", "MDDriver", "QPQ", "PQP", "Gradientflow",
    "StoutSmearing", "gauge_configuration", "gauge_momenta", "gaussian_momenta", "gaussian_momenta_",
    "copy_configuration", "copy_configuration_", "measure_plaquette", "calculate_Plaquette",
    "measure_polyakov_loop", "make_loops_fromname", "md_driver", "md_trajectory_", "md_hamiltonian",
    "md_step_", "update_gaugefields_", "update_momenta_", "md_force_", "gradient_flow", "flow_",
    "energy_density", "stout_smearing", "smear", "Philox4x32", "GfbError", "gauge_lattice_size",
    "gauge_num_colors", "gauge_process_grid", "download_configuration", "upload_configuration_",
    "calc_smearedU", "back_prop", "calc_dSdU", "stout_force_", "evaluate_GaugeAction", "StoutWorkspace", "stout_hamiltonian", "reunitarize_", "normalize_U_", "MDActionSet", "MDForceGroup", "SextonWeingarten",
    "Gradientflow_general", "topological_charge", "topological_charge_density", "Heatbath", "heatbath_", "overrelaxation_",
    "MatrixField", "link_field", "shift_U", "clear_U_", "unit_U_", "substitute_U_", "mul_", "add_U_", "tr",
    "Traceless_antihermitian_", "Traceless_antihermitian_add_", "exptU_",
]


class Philox4x32:
    """SiteRNGAlgorithm default (src/API.jl:189)."""
    code = 0


class QPQ:
    """md_step! ordering Q(e/2) P(e) Q(e/2) (src/molecular_dynamics.jl:611-616)."""
    code = 0


class PQP:
    """md_step! ordering P(e/2) Q(e) P(e/2) (src/molecular_dynamics.jl:604-609)."""
    code = 1


class B200Backend:
    """Backend selector, the analogue of LatticeMatricesBackend() (src/API.jl:12-27).

    One instance owns one libgfb200 context.  `ngpu` local GPUs are driven from this process; when
    torch.distributed is initialised with world_size > 1 (one process per GPU, as bench.py is
    launched by torchrun) the ranks form one context over NCCL and each rank owns one t-slab.
    """

    _default = None

    def __init__(self, ngpu=1, devices=None, distributed=None):
        lib = _lib.load()
        self._ctx = ctypes.c_void_p()
        self.rank, self.world = 0, 1
        if distributed is None:
            distributed = _dist_world() > 1
        if distributed:
            import torch
            import torch.distributed as dist

            self.rank, self.world = dist.get_rank(), dist.get_world_size()
            uid = ctypes.create_string_buffer(128)
            if self.rank == 0:
                check(lib.gfb_nccl_unique_id(uid))
            # share the NCCL id through the existing process group (gloo or nccl), like the reference's
            # rank-0 seed broadcast (src/AbstractGaugefields.jl:135-150)
            t = torch.tensor(list(uid.raw), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            dist.broadcast(t, src=0)
            uid = ctypes.create_string_buffer(bytes(t.cpu().tolist()), 128)
            device = int(os.environ.get("LOCAL_RANK", self.rank)) if devices is None else int(devices[0])
            check(lib.gfb_init_rank(self.rank, self.world, uid, device, ctypes.byref(self._ctx)))
        else:
            dev = None
            if devices is not None:
                dev = (ctypes.c_int * int(ngpu))(*[int(d) for d in devices])
            check(lib.gfb_init(int(ngpu), dev, ctypes.byref(self._ctx)))
        local, total = ctypes.c_int(), ctypes.c_int()
        check(lib.gfb_num_slabs(self._ctx, ctypes.byref(local), ctypes.byref(total)), self._ctx)
        self.local_slabs, self.total_slabs = local.value, total.value

    @classmethod
    def default(cls):
        if cls._default is None:
            cls._default = cls()
        return cls._default

    # -- plumbing -------------------------------------------------------------------------------
    @property
    def lib(self):
        return _lib.load()

    def call(self, name, *args):
        check(getattr(self.lib, name)(*args), self._ctx)

    def sync(self):
        self.call("gfb_sync", self._ctx)

    def tic(self):
        self.call("gfb_timer_tic", self._ctx)

    def toc(self):
        ms = ctypes.c_double()
        self.call("gfb_timer_toc", self._ctx, ctypes.byref(ms))
        return ms.value

    def kernel_launches(self):
        n = ctypes.c_longlong()
        self.call("gfb_kernel_launches", self._ctx, ctypes.byref(n))
        return n.value

    def t_range(self, nt):
        """Global t-range owned by this process (all of it unless one-process-per-GPU)."""
        if self.world == 1:
            return 0, nt
        tloc = nt // self.world
        return self.rank * tloc, (self.rank + 1) * tloc

    def finalize(self):
        if self._ctx:
            self.lib.gfb_finalize(self._ctx)
            self._ctx = ctypes.c_void_p()


def _dist_world():
    try:
        import torch.distributed as dist
    except Exception:
        return 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


def pinned_empty(shape, dtype):
    """numpy array over cudaMallocHost memory (fast upload/download)."""
    lib = _lib.load()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = ctypes.c_void_p()
    check(lib.gfb_host_alloc(ctypes.byref(ptr), n))
    buf = (ctypes.c_char * n).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr


# ------------------------------------------------------------------------------------------------
# field types
# ------------------------------------------------------------------------------------------------
class _LinkView:
    """U[mu]: one direction of a configuration (a Gaugefields_4D in the reference)."""

    def __init__(self, parent, mu):
        self.parent, self.mu = parent, mu
        self.NC, self.NDW = 3, 1
        self.NX, self.NY, self.NZ, self.NT = parent.lattice
        self.NV = parent.NV

    def to_host(self):
        return self.parent.to_host()[self.mu]


class GaugeConfiguration:
    """The `Vector` of four link fields returned by gauge_configuration (src/API.jl:178-251)."""

    def __init__(self, backend, lattice):
        lattice = tuple(int(v) for v in lattice)
        if len(lattice) != 4:
            raise ValueError("the B200 backend supports 4 dimensions; got %d" % len(lattice))
        if not all(v > 0 for v in lattice):
            raise ValueError("all lattice extents must be positive; got %s" % (lattice,))
        self.backend, self.lattice = backend, lattice
        self.NC, self.NDW = 3, 1
        self.NV = int(np.prod(lattice))
        self._h = ctypes.c_void_p()
        nx, ny, nz, nt = lattice
        backend.call("gfb_gauge_alloc", backend._ctx, nx, ny, nz, nt, ctypes.byref(self._h))

    def __len__(self):
        return 4

    def __getitem__(self, mu):
        if not 0 <= mu < 4:
            raise IndexError(mu)
        return _LinkView(self, mu)

    def __del__(self):
        try:
            if self._h and self.backend._ctx:
                self.backend.lib.gfb_gauge_free(self._h)
        except Exception:
            pass

    def host_shape(self):
        nx, ny, nz, nt = self.lattice
        return (4, nt, nz, ny, nx, 3, 3)

    def local_shape(self):
        nx, ny, nz, nt = self.lattice
        t0, t1 = self.backend.t_range(nt)
        return (4, t1 - t0, nz, ny, nx, 3, 3)

    def _ptr(self, host, mu, local):
        # the C ABI takes the GLOBAL array of one direction and touches only this context's t-range; a
        # rank-local chunk is passed as the pointer the global array would have had
        base = host[mu].ctypes.data
        if local:
            nx, ny, nz, nt = self.lattice
            t0, _ = self.backend.t_range(nt)
            base -= t0 * nx * ny * nz * 9 * 16
        return ctypes.c_void_p(base)

    def upload(self, host, local=False):
        """Load a (4,NT,NZ,NY,NX,3,3) complex128 array (gathered layout); each process copies its own t-range.
        With local=True `host` holds only this process's t-range (shape local_shape())."""
        host = np.ascontiguousarray(host, dtype=np.complex128)
        want = self.local_shape() if local else self.host_shape()
        if host.shape != want:
            raise ValueError("expected shape %s, got %s" % (want, host.shape))
        for mu in range(4):
            self.backend.call("gfb_gauge_upload", self._h, mu, self._ptr(host, mu, local))
        self.backend.sync()
        return self

    def to_host(self, out=None, local=False):
        """Gathered host copy (gather_matrix, src/API.jl:533).  One-process-per-GPU: only this rank's t-range is filled."""
        if out is None:
            out = np.zeros(self.local_shape() if local else self.host_shape(), dtype=np.complex128)
        for mu in range(4):
            self.backend.call("gfb_gauge_download", self._h, mu, self._ptr(out, mu, local))
        return out


    def upload_ildg(self, payload, precision=64):
        """Load the ILDG binary payload of the whole lattice (big-endian [t][z][y][x][mu][row][col], src/output/ildg_format.jl:67-83);
        byte swap, precision conversion and transpose run on the GPU.  `payload`: bytes-like of the global lattice."""
        nx, ny, nz, nt = self.lattice
        buf = np.frombuffer(payload, dtype=np.uint8)
        want = nx * ny * nz * nt * 4 * 9 * 2 * (precision // 8)
        if precision not in (32, 64):
            raise ValueError("ILDG precision must be 32 or 64")
        if buf.size != want:
            raise ValueError("ILDG payload has %d bytes; expected %d" % (buf.size, want))
        buf = np.ascontiguousarray(buf)
        self.backend.call("gfb_gauge_upload_ildg", self._h, ctypes.c_void_p(buf.ctypes.data), precision)
        self.backend.sync()
        return self

    def to_ildg(self, precision=64):
        """ILDG binary payload (bytes) of the whole lattice (_save_binarydata, src/output/ildg_format.jl:697-746).
        One-process-per-GPU: only this rank's time-slices are filled."""
        nx, ny, nz, nt = self.lattice
        if precision not in (32, 64):
            raise ValueError("ILDG precision must be 32 or 64")
        out = np.zeros(nx * ny * nz * nt * 4 * 9 * 2 * (precision // 8), dtype=np.uint8)
        self.backend.call("gfb_gauge_download_ildg", self._h, ctypes.c_void_p(out.ctypes.data), precision)
        return out.tobytes()


class Momenta:
    """Conjugate momenta: four 8-coefficient fields (initialize_TA_Gaugefields, src/TA_Gaugefields.jl:140-195)."""

    def __init__(self, backend, lattice):
        self.backend, self.lattice = backend, tuple(int(v) for v in lattice)
        self._h = ctypes.c_void_p()
        nx, ny, nz, nt = self.lattice
        backend.call("gfb_mom_alloc", backend._ctx, nx, ny, nz, nt, ctypes.byref(self._h))

    def __len__(self):
        return 4

    def __del__(self):
        try:
            if self._h and self.backend._ctx:
                self.backend.lib.gfb_mom_free(self._h)
        except Exception:
            pass

    def host_shape(self):
        nx, ny, nz, nt = self.lattice
        return (4, nt, nz, ny, nx, 8)

    def local_shape(self):
        nx, ny, nz, nt = self.lattice
        t0, t1 = self.backend.t_range(nt)
        return (4, t1 - t0, nz, ny, nx, 8)

    def _ptr(self, host, mu, local):
        base = host[mu].ctypes.data
        if local:
            nx, ny, nz, nt = self.lattice
            t0, _ = self.backend.t_range(nt)
            base -= t0 * nx * ny * nz * 8 * 8
        return ctypes.c_void_p(base)

    def upload(self, host, local=False):
        host = np.ascontiguousarray(host, dtype=np.float64)
        want = self.local_shape() if local else self.host_shape()
        if host.shape != want:
            raise ValueError("expected shape %s, got %s" % (want, host.shape))
        for mu in range(4):
            self.backend.call("gfb_mom_upload", self._h, mu, self._ptr(host, mu, local))
        self.backend.sync()
        return self

    def to_host(self, out=None, local=False):
        if out is None:
            out = np.zeros(self.local_shape() if local else self.host_shape(), dtype=np.float64)
        for mu in range(4):
            self.backend.call("gfb_mom_download", self._h, mu, self._ptr(out, mu, local))
        return out

    def dot(self, other=None):
        """p * p (src/TA_Gaugefields.jl:127-137)."""
        if other is not None and other is not self:
            raise NotImplementedError("only p*p is provided")
        v = ctypes.c_double()
        self.backend.call("gfb_kinetic", self._h, ctypes.byref(v))
        return v.value

    def __mul__(self, other):
        return self.dot(other)

    def clear_(self):
        self.backend.call("gfb_mom_zero", self._h)
        return self

    def add_(self, t, other, mu=None):
        """add_U!(P, t, F) on all four directions, or add_U!(P[mu], t, F[mu]) (TA_gaugefields_4D_serial.jl:150-173): how an
        external force provider (md_force! of a fermion action, say) adds into the momenta of this backend."""
        if mu is None:
            self.backend.call("gfb_mom_axpy", self._h, float(t), other._h)
        else:
            self.backend.call("gfb_mom_axpy_dir", self._h, int(mu), float(t), other._h)
        return self


# ------------------------------------------------------------------------------------------------
# configuration API (src/API.jl)
# ------------------------------------------------------------------------------------------------
def gauge_configuration(lattice, backend=None, colors=3, halo=1, start="cold", seed=None, process_grid=None,
                        boundary="periodic", eltype=np.complex128, rng=Philox4x32, verbose=0):
    """gauge_configuration(lattice; kwargs...) (src/API.jl:178-251) on the B200 backend."""
    lattice = tuple(lattice)
    if len(lattice) != 4:
        raise ValueError("the B200 backend supports 4 dimensions; got %d" % len(lattice))
    if colors != 3:
        raise ValueError("the B200 backend supports colors=3; got %s" % colors)
    start = str(start).lstrip(":")
    if start not in ("cold", "hot"):
        raise ValueError("start must be :cold or :hot; got %s" % start)
    if int(halo) < 0:
        raise ValueError("halo must be nonnegative; got %s" % halo)
    if boundary != "periodic":
        raise ValueError("the B200 backend supports periodic boundaries only")
    if np.dtype(eltype) != np.complex128:
        raise ValueError("the B200 backend computes in ComplexF64")
    backend = backend or B200Backend.default()
    U = GaugeConfiguration(backend, lattice)
    if start == "cold":
        backend.call("gfb_set_cold", U._h)
    else:
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little") if backend.world == 1 else 0
        backend.call("gfb_set_hot", U._h, int(seed) & (2**64 - 1), rng.code)
    return U


def gauge_lattice_size(U):
    return U.lattice


def gauge_num_colors(U):
    return 3


def gauge_process_grid(U):
    """From Julia's point of view the field is undecomposed; the t-slabs are internal."""
    return (1, 1, 1, U.backend.total_slabs)


def copy_configuration_(destination, source):
    """copy_configuration!(destination, source) (src/API.jl:307-322)."""
    if destination.lattice != source.lattice:
        raise ValueError("destination and source lattice sizes differ")
    destination.backend.call("gfb_gauge_copy", destination._h, source._h)
    return destination


def copy_configuration(source):
    dst = GaugeConfiguration(source.backend, source.lattice)
    return copy_configuration_(dst, source)


def upload_configuration_(U, host):
    return U.upload(host)


def download_configuration(U):
    return U.to_host()


def reunitarize_(U):
    """normalize_U!(U) (src/4D/nowing/gaugefields_4D_nowing.jl:2387-2458): Gram-Schmidt rows 0, 1 and row 2 = conj(row0 x row1).
    Call it after uploading a configuration that is unitary only to single precision (32-bit ILDG files): the fused passes use
    their two-row SU(3) products only on configurations known to be unitary to 1e-12 and full 3x3 products otherwise."""
    U.backend.call("gfb_reunitarize", U._h)
    return U


normalize_U_ = reunitarize_


def gauge_momenta(U):
    """gauge_momenta(U) = initialize_TA_Gaugefields(U) (src/API.jl:331)."""
    return Momenta(U.backend, U.lattice)


def gaussian_momenta_(momenta, sigma=1.0, seed=None, sweep=0, rng=Philox4x32):
    """gaussian_momenta!(momenta; sigma, seed, sweep, rng) (src/API.jl:341-368)."""
    if sweep < 0:
        raise ValueError("sweep must be nonnegative; got %s" % sweep)
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little") if momenta.backend.world == 1 else 0
    momenta.backend.call("gfb_gaussian_momenta", momenta._h, int(seed) & (2**64 - 1), int(sweep), float(sigma), rng.code)
    return momenta


def gaussian_momenta(U, sigma=1.0, seed=None, sweep=0, rng=Philox4x32):
    return gaussian_momenta_(gauge_momenta(U), sigma=sigma, seed=seed, sweep=sweep, rng=rng)


def calculate_Plaquette(U, temp1=None, temp2=None):
    """Un-normalised sum over x, mu<nu of Re tr P (src/AbstractGaugefields.jl:2684-2699)."""
    v = ctypes.c_double()
    U.backend.call("gfb_plaquette_sum", U._h, ctypes.byref(v))
    return v.value


def measure_plaquette(U, normalize=True):
    """measure_plaquette(U; normalize=true) (src/API.jl:395-404)."""
    value = calculate_Plaquette(U)
    if not normalize:
        return value
    return value / (math.comb(4, 2) * U.NV * 3)


def measure_polyakov_loop(U, normalize=True):
    """measure_polyakov_loop (src/API.jl:412-417)."""
    out = (ctypes.c_double * 2)()
    U.backend.call("gfb_polyakov", U._h, out)
    value = complex(out[0], out[1])
    return value / 3 if normalize else value


def energy_density(U, kind="clover"):
    """E from the clover definition of samples/measurements/energydensity.jl:4-78, or 2*sum_{mu<nu} Re tr(1-P)/V."""
    v = ctypes.c_double()
    U.backend.call("gfb_energy_density", U._h, {"clover": 0, "plaquette": 1}[kind], ctypes.byref(v))
    return v.value


# ------------------------------------------------------------------------------------------------
# actions (src/action/GaugeActions.jl)
# ------------------------------------------------------------------------------------------------
class _Loops:
    def __init__(self, names):
        self.names = list(names)  # entries: ("plaquette", dagger?)

    def adjoint(self):
        return _Loops([(n, not d) for n, d in self.names])

    def __add__(self, other):
        return _Loops(self.names + other.names)


def make_loops_fromname(name, Dim=4):
    """make_loops_fromname("plaquette" | "rectangular", Dim=4) (Wilsonloop.jl; docs/src/hmc.md:146-150,
    docs/src/wilsonloops_actions.md:27): the six plaquettes / the twelve 1x2 and 2x1 rectangles of the mu < nu planes."""
    if Dim != 4:
        raise ValueError("the B200 backend supports Dim=4")
    if name not in ("plaquette", "rectangular"):
        raise NotImplementedError("the B200 backend builds the plaquette and rectangular loop sets; got %r" % (name,))
    return _Loops([(name, False)])


class GaugeAction:
    """GaugeAction(U) + push!(action, coefficient, loops) (src/action/GaugeActions.jl:20-62).

    Terms are (real coefficient, loops + loops') with loops = "plaquette" or "rectangular"; the plaquette-only action
    (Wilson, beta = 2 * coefficient) runs the t-marching kernels, an action with a rectangle term the general-action
    kernels (csrc/general.cu).
    """

    def __init__(self, U):
        self.lattice = U.lattice
        self.terms = []

    def push(self, coefficient, loops):
        self.terms.append((complex(coefficient), loops))
        return self

    def coefficients(self):
        """(c_plaq, c_rect): the summed coefficients of the plaquette and rectangle terms."""
        if not self.terms:
            raise ValueError("the action has no terms")
        c = {"plaquette": 0.0, "rectangular": 0.0}
        for coeff, loops in self.terms:
            kinds = sorted(loops.names)
            if len(kinds) != 2 or kinds[0][0] != kinds[1][0] or [d for _, d in kinds] != [False, True]:
                raise NotImplementedError("push!(action, coefficient, [loops; loops']) with loops = plaquette or rectangular")
            if coeff.imag != 0.0:
                raise NotImplementedError("complex loop coefficients are not fused on the B200 backend")
            c[kinds[0][0]] += coeff.real
        return c["plaquette"], c["rectangular"]

    def wilson_beta(self):
        cp, cr = self.coefficients()
        if cr != 0.0:
            raise NotImplementedError("this call takes the Wilson action only (plaquette + plaquette' terms)")
        return 2.0 * cp


def evaluate_GaugeAction(action, U):
    """evaluate_GaugeAction (GaugeActions.jl:132-142): sum over terms of coefficient * sum_x tr(loops + loops')
    = 2 (c_plaq sum Re tr P + c_rect sum Re tr R); beta * sum Re tr P for the Wilson action."""
    cp, cr = action.coefficients()
    if cr == 0.0:
        v = ctypes.c_double()
        U.backend.call("gfb_wilson_action", U._h, 2.0 * cp, ctypes.byref(v))
        return complex(v.value, 0.0)
    v = (ctypes.c_double * 2)()
    U.backend.call("gfb_loop_sums", U._h, v)
    return complex(2.0 * (cp * v[0] + cr * v[1]), 0.0)


# ------------------------------------------------------------------------------------------------
# molecular dynamics (src/molecular_dynamics.jl)
# ------------------------------------------------------------------------------------------------
class MDActionSet:
    """MDActionSet(; name = action, ...) (src/molecular_dynamics.jl:57-69): named action providers whose forces add."""

    def __init__(self, **terms):
        if not terms:
            raise ValueError("an MDActionSet must contain at least one action provider")
        self.terms = dict(terms)

    def coefficients(self, names=None):
        cp = cr = 0.0
        for n in (self.terms if names is None else names):
            if n not in self.terms:
                raise ValueError("the action set has no member %r" % (n,))
            a, b = self.terms[n].coefficients()
            cp, cr = cp + a, cr + b
        return cp, cr


class MDForceGroup:
    """MDForceGroup(names...) (src/molecular_dynamics.jl:78-92)."""

    def __init__(self, *names):
        if len(names) == 1 and isinstance(names[0], (tuple, list)):
            names = tuple(names[0])
        if not names:
            raise ValueError("an MDForceGroup must contain at least one action name")
        if len(set(names)) != len(names):
            raise ValueError("an MDForceGroup must not contain duplicate action names: %s" % (names,))
        self.names = tuple(names)


def _force_group(g):
    return g if isinstance(g, MDForceGroup) else MDForceGroup(g)


class SextonWeingarten:
    """SextonWeingarten(; fast, slow, n_fast, ordering=QPQ()) nested leapfrog (src/molecular_dynamics.jl:355-410, 618-700)."""

    def __init__(self, fast, slow, n_fast, ordering=QPQ):
        if int(n_fast) <= 0:
            raise ValueError("n_fast must be positive; got %s" % (n_fast,))
        self.fast, self.slow, self.n_fast = _force_group(fast), _force_group(slow), int(n_fast)
        self.ordering = ordering if isinstance(ordering, type) else type(ordering)
        if self.ordering not in (QPQ, PQP):
            raise ValueError("ordering must be QPQ or PQP")
        if set(self.fast.names) & set(self.slow.names):
            raise ValueError("an action must not be in both the fast and the slow group")


class MDDriver:
    """Preallocated deterministic MD driver (src/molecular_dynamics.jl:413-423)."""

    def __init__(self, action, integrator, trajectory_length, steps, force, fused):
        self.action, self.integrator = action, integrator
        self.trajectory_length, self.steps = float(trajectory_length), int(steps)
        self.force, self.fused = force, bool(fused)
        self.c_plaq, self.c_rect = action.coefficients()
        self.beta = 2.0 * self.c_plaq if self.c_rect == 0.0 else None


def md_driver(U, action, steps=None, trajectory_length=1.0, integrator=QPQ, fused=True):
    """md_driver(U, action; steps, trajectory_length=1.0, integrator=QPQ()) (src/molecular_dynamics.jl:440-483).
    `action` is a GaugeAction or an MDActionSet; `integrator` QPQ, PQP or a SextonWeingarten instance.

    `fused=True` (default) runs each kick together with the following link update in one kernel
    and merges adjacent half link updates; `fused=False` issues the reference's op sequence.
    """
    if steps is None:
        raise TypeError("md_driver requires the keyword `steps`")
    if steps <= 0:
        raise ValueError("steps must be positive; got %s" % steps)
    if not math.isfinite(trajectory_length):
        raise ValueError("trajectory_length must be finite; got %s" % trajectory_length)
    if trajectory_length == 0:
        raise ValueError("trajectory_length must not be zero")
    if isinstance(integrator, SextonWeingarten):
        if not isinstance(action, MDActionSet):
            raise ValueError("SextonWeingarten selects named members of an MDActionSet")
        for n in integrator.fast.names + integrator.slow.names:
            if n not in action.terms:
                raise ValueError("the action set has no member %r" % (n,))
        return MDDriver(action, integrator, trajectory_length, steps, gauge_momenta(U), fused)
    integ = integrator if isinstance(integrator, type) else type(integrator)
    if integ not in (QPQ, PQP):
        raise ValueError("md_step! is not implemented for %r" % (integrator,))
    return MDDriver(action, integ, trajectory_length, steps, gauge_momenta(U), fused)


def md_step_size(driver):
    return driver.trajectory_length / driver.steps


def md_hamiltonian(U, p, driver):
    """md_hamiltonian (src/molecular_dynamics.jl:494-505): -(1/NC) Re evaluate_GaugeAction + p*p/2
    (= -(beta/3) sum Re tr P + p*p/2 for the Wilson action)."""
    v = ctypes.c_double()
    U.backend.call("gfb_hamiltonian_general", U._h, p._h, driver.c_plaq, driver.c_rect, ctypes.byref(v))
    return v.value


def update_gaugefields_(U, P, step_size, driver=None):
    """update_gaugefields!(U, P, step_size, driver) (src/molecular_dynamics.jl:513-531)."""
    if not math.isfinite(step_size):
        raise ValueError("the gauge-field step size must be finite; got %s" % step_size)
    U.backend.call("gfb_update_links", U._h, P._h, float(step_size))
    return U


def update_momenta_(P, U, step_size, driver, group=None):
    """update_momenta!(P, U, step_size, driver[, group]) (src/molecular_dynamics.jl:539-583): with an MDForceGroup only the
    named members of the driver's MDActionSet kick (the elementary kick of SextonWeingarten)."""
    if not math.isfinite(step_size):
        raise ValueError("the momentum step size must be finite; got %s" % step_size)
    if group is not None:
        if not isinstance(driver.action, MDActionSet):
            raise ValueError("a force group selects members of an MDActionSet")
        cp, cr = driver.action.coefficients(_force_group(group).names)
    else:
        cp, cr = driver.c_plaq, driver.c_rect
    U.backend.call("gfb_update_momenta_general", P._h, U._h, float(step_size), cp, cr)
    return P


def md_force_(force, action, U, workspace=None, group=None):
    """md_force!(force, action, U, workspace[, group]) (src/molecular_dynamics.jl:251-267, 205-236)."""
    if group is not None:
        cp, cr = action.coefficients(_force_group(group).names)
    else:
        cp, cr = action.coefficients()
    U.backend.call("gfb_force_general", force._h, U._h, cp, cr)
    return None


def _fast_qpq(U, P, duration, nsteps, driver, fast):
    """_md_fast_qpq! (src/molecular_dynamics.jl:618-635)."""
    eps = duration / nsteps
    update_gaugefields_(U, P, eps / 2, driver)
    for k in range(nsteps):
        update_momenta_(P, U, eps, driver, fast)
        update_gaugefields_(U, P, eps / 2 if k == nsteps - 1 else eps, driver)


def md_step_(integrator, U, P, step_size, driver):
    """md_step! for PQP / QPQ / SextonWeingarten / a callable (src/molecular_dynamics.jl:585-700)."""
    if isinstance(integrator, SextonWeingarten):
        sw = integrator
        if sw.ordering is QPQ:
            _fast_qpq(U, P, step_size / 2, sw.n_fast, driver, sw.fast)
            update_momenta_(P, U, step_size, driver, sw.slow)
            _fast_qpq(U, P, step_size / 2, sw.n_fast, driver, sw.fast)
        else:
            update_momenta_(P, U, step_size / 2, driver, sw.slow)
            _fast_qpq(U, P, step_size, sw.n_fast, driver, sw.fast)
            update_momenta_(P, U, step_size / 2, driver, sw.slow)
        return None
    if callable(integrator) and not isinstance(integrator, type):
        return integrator(U, P, step_size, driver)
    integ = integrator if isinstance(integrator, type) else type(integrator)
    if integ is PQP:
        update_momenta_(P, U, step_size / 2, driver)
        update_gaugefields_(U, P, step_size, driver)
        update_momenta_(P, U, step_size / 2, driver)
    elif integ is QPQ:
        update_gaugefields_(U, P, step_size / 2, driver)
        update_momenta_(P, U, step_size, driver)
        update_gaugefields_(U, P, step_size / 2, driver)
    else:
        raise ValueError("md_step! is not implemented for %r" % (integrator,))
    return None


class TrajectoryResult(tuple):
    """Named tuple (initial_hamiltonian, final_hamiltonian, delta_hamiltonian)."""

    def __new__(cls, h0, h1):
        return super().__new__(cls, (h0, h1, h1 - h0))

    initial_hamiltonian = property(lambda s: s[0])
    final_hamiltonian = property(lambda s: s[1])
    delta_hamiltonian = property(lambda s: s[2])


def md_trajectory_(U, P, driver, diagnostics=True):
    """md_trajectory!(U, p, driver; diagnostics=true) (src/molecular_dynamics.jl:712-730).  QPQ / PQP trajectories are ONE
    library call (fused kick+drift kernels); a SextonWeingarten or callable integrator runs the reference's host loop over
    md_step!, each elementary kick / drift one kernel."""
    if U.lattice != P.lattice:
        raise ValueError("U and P must have the same number of directions / lattice")
    if driver.integrator in (QPQ, PQP):
        H = (ctypes.c_double * 2)() if diagnostics else None
        U.backend.call(
            "gfb_md_trajectory_general", U._h, P._h, driver.c_plaq, driver.c_rect, driver.steps, driver.trajectory_length,
            driver.integrator.code, 1 if driver.fused else 0, H,
        )
        if not diagnostics:
            return None
        return TrajectoryResult(H[0], H[1])
    h0 = md_hamiltonian(U, P, driver) if diagnostics else None
    eps = md_step_size(driver)
    for _ in range(driver.steps):
        md_step_(driver.integrator, U, P, eps, driver)
    if not diagnostics:
        return None
    return TrajectoryResult(h0, md_hamiltonian(U, P, driver))


# ------------------------------------------------------------------------------------------------
# gradient flow (src/smearing/gradientflow.jl) and stout smearing (src/smearing/stout_fast.jl)
# ------------------------------------------------------------------------------------------------
class Gradientflow:
    """Gradientflow(U; Nflow=1, eps=0.01) (src/smearing/gradientflow.jl:118-161)."""

    def __init__(self, U, Nflow=1, eps=0.01):
        self.Nflow, self.eps = int(Nflow), float(eps)


def gradient_flow(U, steps=1, step_size=0.01):
    """gradient_flow(U; steps, step_size) (src/API.jl:420-424)."""
    if steps <= 0:
        raise ValueError("steps must be positive; got %s" % steps)
    if not step_size > 0:
        raise ValueError("step_size must be positive; got %s" % step_size)
    return Gradientflow(U, Nflow=steps, eps=step_size)


class Gradientflow_general:
    """Gradientflow_general(U, linknames, linkvalues; Nflow=1, eps=0.01) (src/smearing/gradientflow.jl:33-116): RK3 flow of
    the action sum_i linkvalues[i] * (loops_i + loops_i'), loops by name ("plaquette", "rectangular"), real values."""

    def __init__(self, U, linknames, linkvalues, Nflow=1, eps=0.01):
        if len(linknames) != len(linkvalues):
            raise ValueError("linknames and linkvalues must have the same length")
        action = GaugeAction(U)
        for name, value in zip(linknames, linkvalues):
            loops = make_loops_fromname(str(name).lstrip(":"))
            action.push(value, loops + loops.adjoint())
        self.c_plaq, self.c_rect = action.coefficients()
        self.Nflow, self.eps = int(Nflow), float(eps)


def flow_(U, g):
    """flow!(U, g::Gradientflow | Gradientflow_general) (src/smearing/gradientflow.jl:171-316): g.Nflow RK3 steps."""
    if isinstance(g, Gradientflow_general):
        U.backend.call("gfb_flow_general", U._h, g.eps, g.Nflow, g.c_plaq, g.c_rect)
    else:
        U.backend.call("gfb_flow", U._h, g.eps, g.Nflow)
    return U


class Heatbath:
    """Heatbath(U, beta; seed, sweep=0, rng_algorithm=Philox4x32()) (src/heatbath/heatbathmodule.jl:55-99): the Wilson-action
    heatbath / overrelaxation updater.  `sweep` counts the sweeps done and keys their random streams."""

    def __init__(self, U, beta, seed=0, sweep=0, rng_algorithm=Philox4x32):
        if rng_algorithm not in (Philox4x32,) and not isinstance(rng_algorithm, Philox4x32):
            raise ValueError("the B200 backend implements the Philox4x32 site RNG")
        if not (beta > 0 and math.isfinite(beta)):
            raise ValueError("beta must be positive and finite; got %s" % (beta,))
        self.beta, self.seed, self.sweep, self.overrelaxation_sweep = float(beta), int(seed), int(sweep), int(sweep)


def heatbath_(U, h):
    """heatbath!(U, h::Heatbath) (src/heatbath/heatbathmodule.jl:843-845): one sweep over 4 directions x 2 colours."""
    U.backend.call("gfb_heatbath", U._h, h.beta, h.seed, h.sweep, 0)
    h.sweep += 1
    return U


def overrelaxation_(U, h):
    """overrelaxation!(U, h::Heatbath) (src/heatbath/heatbathmodule.jl:847-852): one microcanonical sweep."""
    U.backend.call("gfb_overrelaxation", U._h, h.beta, h.seed, h.overrelaxation_sweep, 0)
    h.overrelaxation_sweep += 1
    return U


_TOPO = {"plaquette": 0, "clover": 1, "improved": 2}


def topological_charge(U, method="plaquette"):
    """topological_charge(U; method=:plaquette | :clover | :improved) (src/AbstractGaugefields.jl:1471-1490)."""
    m = str(method).lstrip(":")
    if m not in _TOPO:
        raise ValueError("supported topological_charge methods are :plaquette, :clover, and :improved")
    v = ctypes.c_double()
    U.backend.call("gfb_topological_charge", U._h, _TOPO[m], ctypes.byref(v))
    return v.value


def topological_charge_density(U, method="plaquette"):
    """topological_charge_density(U; method) (src/AbstractGaugefields.jl:1447-1461): q(x) as an array indexed [t, z, y, x]
    (the reference's density[ix, iy, iz, it] in column-major order)."""
    m = str(method).lstrip(":")
    if m not in _TOPO:
        raise ValueError("supported topological_charge_density methods are :plaquette, :clover, and :improved")
    nx, ny, nz, nt = U.lattice
    out = np.zeros((nt, nz, ny, nx), dtype=np.float64)
    U.backend.call("gfb_topological_charge_density", U._h, _TOPO[m], out.ctypes.data_as(ctypes.c_void_p))
    return out


class StoutSmearing:
    """One or more plaquette-staple stout layers (stout_smearing, src/API.jl:452-466)."""

    def __init__(self, rhos):
        self.rhos = [float(r) for r in rhos]


def stout_smearing(U, loops="plaquette", rho=0.1, layers=1):
    if str(loops).lstrip(":") != "plaquette":
        raise NotImplementedError("the B200 backend implements plaquette-staple stout smearing")
    return StoutSmearing([rho] * int(layers))


def smear(U, smearing, record=False):
    """smear(U, smearing; record=false) (src/API.jl:468-480): returns the smeared configuration
    (and, with record=True, the per-layer inputs needed by back_prop)."""
    history = []
    cur = U
    for rho in smearing.rhos:
        out = GaugeConfiguration(U.backend, U.lattice)
        U.backend.call("gfb_stout_forward", out._h, cur._h, rho, None)
        history.append(cur)
        cur = out
    if record:
        return {"configuration": cur, "history": history}
    return cur


def calc_smearedU(U, smearing):
    """calc_smearedU(Uin, nn) (src/smearing/Abstractsmearing.jl:237-314): returns (Uout, Uout_multi) with Uout_multi the
    outputs of layers 1..n-1 (the tape back_prop needs besides Uin)."""
    outs = []
    cur = U
    for rho in smearing.rhos:
        out = GaugeConfiguration(U.backend, U.lattice)
        U.backend.call("gfb_stout_forward", out._h, cur._h, rho, None)
        outs.append(out)
        cur = out
    return cur, outs[:-1]


def calc_dSdU(action, U):
    """calc_dSdUμ! for all four directions (src/action/GaugeActions.jl:95-123) as one configuration-shaped field."""
    D = GaugeConfiguration(U.backend, U.lattice)
    U.backend.call("gfb_wilson_dSdU", D._h, U._h, action.wilson_beta())
    return D


def back_prop(dSdU, smearing, Uout_multi, Uin):
    """back_prop(δL, nn, Uout_multi, Uin) (src/smearing/Abstractsmearing.jl:352-411): pull dS/dU' at the smeared links back
    through the stout layers to dS/dU at the bare links."""
    inputs = [Uin] + list(Uout_multi)
    cur = dSdU
    for rho, inp in zip(reversed(smearing.rhos), reversed(inputs)):
        prev = GaugeConfiguration(Uin.backend, Uin.lattice)
        Uin.backend.call("gfb_stout_backward", prev._h, cur._h, inp._h, rho)
        cur = prev
    return cur


class StoutWorkspace:
    """Preallocated fields for the stout-smeared force (the reference allocates its layer outputs once in `calc_smearedU`'s
    caller, test/HMCstout_test_nowing.jl:60-75): the layer outputs, dS/dU at the smeared links and the pulled-back derivatives."""

    def __init__(self, U, smearing):
        n = len(smearing.rhos)
        self.outs = [GaugeConfiguration(U.backend, U.lattice) for _ in range(n)]
        self.dS = GaugeConfiguration(U.backend, U.lattice)
        self.prev = [GaugeConfiguration(U.backend, U.lattice) for _ in range(n)]


def stout_force_(P, U, action, smearing, step_size, workspace=None, timer=None):
    """The momentum kick of HMC with a stout-smeared action as the user composes it in
    test/HMCstout_test_nowing.jl:99-118: Uout = calc_smearedU(U); dSdU = calc_dSdUμ(action, Uout);
    dSdUbare = back_prop(dSdU); P_mu += -step_size/NC * TA(U_mu dSdUbare_mu).
    `workspace` (StoutWorkspace) avoids allocating seven configurations per call; `timer(fn) -> ms` (bench.py) runs each
    stage through it and makes the call return {"forward", "dSdU", "back_prop", "kick"} in ms instead of P."""
    ws = workspace if workspace is not None else StoutWorkspace(U, smearing)
    be = U.backend
    beta = action.wilson_beta()
    inputs = [U] + ws.outs[:-1]

    def forward():
        for rho, inp, out in zip(smearing.rhos, inputs, ws.outs):
            be.call("gfb_stout_forward", out._h, inp._h, rho, None)

    def dsdu():
        be.call("gfb_wilson_dSdU", ws.dS._h, ws.outs[-1]._h, beta)

    def backward():
        cur = ws.dS
        for rho, inp, prev in zip(reversed(smearing.rhos), reversed(inputs), reversed(ws.prev)):
            be.call("gfb_stout_backward", prev._h, cur._h, inp._h, rho)
            cur = prev

    def kick():
        be.call("gfb_kick_from_dSdU", P._h, U._h, ws.prev[0]._h, -float(step_size) / 3.0)

    if timer is not None:
        return {"forward": timer(forward), "dSdU": timer(dsdu), "back_prop": timer(backward), "kick": timer(kick)}
    forward(); dsdu(); backward(); kick()
    return P


def stout_hamiltonian(U, P, action, smearing, workspace=None):
    """H = S(smeared U) + p.p/2 for the stout-smeared action (test/HMCstout_test_nowing.jl:77-97)."""
    ws = workspace if workspace is not None else StoutWorkspace(U, smearing)
    inputs = [U] + ws.outs[:-1]
    for rho, inp, out in zip(smearing.rhos, inputs, ws.outs):
        U.backend.call("gfb_stout_forward", out._h, inp._h, rho, None)
    # same convention as md_hamiltonian (molecular_dynamics.jl:494-505, :247-249): S = -(beta/NC) sum Re tr P
    return -(action.wilson_beta() / 3.0) * calculate_Plaquette(ws.outs[-1]) + 0.5 * P.dot()


# ------------------------------------------------------------------------------------------------
# primitive table: the element-wise operations the reference's generic algorithms are written in
# (src/4D/mpi_jacc/gaugefields_4D_MPILattice.jl:474-842; semantics SURVEY.md Appendix A)
# ------------------------------------------------------------------------------------------------
class _Lazy:
    """shift_U(A, s) / A' : lazy views (Shifted_/Adjoint_Gaugefields_4D_MPILattice, :647-689)."""

    def __init__(self, field, shift=(0, 0, 0, 0), dagger=False):
        self.field, self.shift, self.dagger = field, tuple(int(v) for v in shift), bool(dagger)

    def adjoint(self):
        return _Lazy(self.field, self.shift, not self.dagger)

    @property
    def H(self):
        return self.adjoint()


def _unlazy(a):
    if isinstance(a, _Lazy):
        return a.field, a.shift, a.dagger
    return a, (0, 0, 0, 0), False


def _shift_arg(shift):
    return (ctypes.c_int * 4)(*shift) if any(shift) else None


class MatrixField:
    """One Gaugefields_4D: a 3x3 complex matrix per site (a link direction U[mu] or a temporary from similar(U[1]))."""

    def __init__(self, backend, lattice, handle=None, owner=None):
        self.backend, self.lattice = backend, tuple(int(v) for v in lattice)
        self.NC, self.NDW = 3, 1
        self.NX, self.NY, self.NZ, self.NT = self.lattice
        self.NV = int(np.prod(self.lattice))
        self._owner = owner  # keeps the viewed configuration alive
        self._h = handle if handle is not None else ctypes.c_void_p()
        if handle is None:
            backend.call("gfb_field_alloc", backend._ctx, *self.lattice, ctypes.byref(self._h))

    def __del__(self):
        try:
            if self._h and self.backend._ctx:
                self.backend.lib.gfb_field_free(self._h)
        except Exception:
            pass

    def similar(self):
        return MatrixField(self.backend, self.lattice)

    def host_shape(self):
        nx, ny, nz, nt = self.lattice
        return (nt, nz, ny, nx, 3, 3)

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=np.complex128)
        if host.shape != self.host_shape():
            raise ValueError("expected shape %s, got %s" % (self.host_shape(), host.shape))
        self.backend.call("gfb_field_upload", self._h, ctypes.c_void_p(host.ctypes.data))
        return self

    def to_host(self):
        out = np.zeros(self.host_shape(), dtype=np.complex128)
        self.backend.call("gfb_field_download", self._h, ctypes.c_void_p(out.ctypes.data))
        return out

    def adjoint(self):
        return _Lazy(self, dagger=True)

    @property
    def H(self):
        return self.adjoint()


def link_field(U, mu):
    """U[mu] as a MatrixField aliasing the configuration (valid until the next fused MD/flow call swaps its buffers)."""
    h = ctypes.c_void_p()
    U.backend.call("gfb_field_view", U._h, int(mu), ctypes.byref(h))
    return MatrixField(U.backend, U.lattice, handle=h, owner=U)


def shift_U(A, shift):
    """shift_U(A, nu) / shift_U(A, (s1,s2,s3,s4)): B(x) = A(x + s), periodic (:647-675).  Integer nu is 1-based like Julia's,
    negative for the backward direction."""
    f, s0, dag = _unlazy(A)
    if isinstance(shift, int):
        v = [0, 0, 0, 0]
        v[abs(shift) - 1] = 1 if shift > 0 else -1
        shift = v
    return _Lazy(f, tuple(a + b for a, b in zip(s0, shift)), dag)


def clear_U_(A):
    A.backend.call("gfb_field_clear", A._h)
    return A


def unit_U_(A):
    A.backend.call("gfb_field_unit", A._h)
    return A


def substitute_U_(A, B):
    """substitute_U!(A, B) with B plain / shifted / adjoint (:509-575)."""
    f, s, dag = _unlazy(B)
    A.backend.call("gfb_field_copy", A._h, f._h, _shift_arg(s), int(dag))
    return A


def mul_(C, A, B, alpha=1.0, beta=0.0):
    """mul!(C, A, B[, alpha, beta]): C = alpha*A*B + beta*C with lazy shifted / adjoint operands (AbstractGaugefields.jl:2082-2105)."""
    fa, sa, da = _unlazy(A)
    fb, sb, db = _unlazy(B)
    alpha, beta = complex(alpha), complex(beta)
    C.backend.call("gfb_mul", C._h, fa._h, _shift_arg(sa), int(da), fb._h, _shift_arg(sb), int(db), alpha.real, alpha.imag, beta.real, beta.imag)
    return C


def add_U_(C, *args):
    """add_U!(C, A) / add_U!(C, alpha, A), A plain or adjoint (:739-772)."""
    alpha, A = (1.0, args[0]) if len(args) == 1 else args
    f, s, dag = _unlazy(A)
    if any(s):
        raise ValueError("add_U! takes plain or adjoint operands")
    alpha = complex(alpha)
    C.backend.call("gfb_axpy", C._h, alpha.real, alpha.imag, f._h, int(dag))
    return C


def tr(A, B=None):
    """tr(A) = sum_x tr A(x);  tr(A, B) = sum_x tr(A(x) B(x)) (:721-728)."""
    out = (ctypes.c_double * 2)()
    if B is None:
        A.backend.call("gfb_tr", A._h, out)
    else:
        A.backend.call("gfb_tr2", A._h, B._h, out)
    return complex(out[0], out[1])


def Traceless_antihermitian_(Q, M):
    """Traceless_antihermitian!(Q, M), matrix -> matrix (:774-781)."""
    Q.backend.call("gfb_ta_project", Q._h, M._h)
    return Q


def Traceless_antihermitian_add_(P, mu, factor, M):
    """Traceless_antihermitian_add!(P[mu], factor, M): 8 coefficients (TA_gaugefields_4D_MPILattice.jl:263-283)."""
    M.backend.call("gfb_ta_coeffs_add", P._h, int(mu), float(factor), M._h)
    return P


def exptU_(E, t, Q, mu=None):
    """exptU!(E, t, Q): Q a matrix field (E = exp(t TA(Q)), :798-808) or the momenta with a direction (TA_...:196-210)."""
    if isinstance(Q, Momenta):
        E.backend.call("gfb_exp_mom", E._h, float(t), Q._h, int(mu))
    else:
        E.backend.call("gfb_exp", E._h, float(t), Q._h)
    return E
