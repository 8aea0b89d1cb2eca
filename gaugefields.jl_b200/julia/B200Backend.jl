# B200Backend.jl -- the Julia side of the drop-in boundary (source only: the build image has no Julia, so this file is checked
# by inspection and by tests/test_abi_and_host.py, which parses every `ccall` below and compares its symbol, return type and
# argument types with include/gfb200.h).
#
# Include this file from Gaugefields.jl (after src/API.jl, src/molecular_dynamics.jl and src/smearing/*.jl are loaded):
#
#     include("B200Backend.jl")          # inside module Gaugefields
#     U = gauge_configuration((32,32,32,32); backend=B200Backend(), start=:hot, seed=0x1234)
#
# It adds a third backend tag next to LatticeMatricesBackend/LegacyBackend (src/API.jl:6-27), field types that hold opaque
# device handles, and methods of the existing generic functions that forward to libgfb200.so through `ccall`.  User scripts
# written against the v1 API (docs/src/hmc.md:128-190, docs/src/highlevelapi.md) run unchanged: only `backend=` differs.
#
# Two layers:
#   * fused overrides (md_trajectory!, update_momenta!, flow!, stout forward!/layer_pullback!, calculate_Plaquette, ...): one
#     library call per trajectory / flow / layer -- the hot path of this backend;
#   * the primitive table (mul!, add_U!, clear_U!, unit_U!, substitute_U!, shift_U, adjoint, tr, similar,
#     Traceless_antihermitian*, exptU!, getindex/setindex!) defined ON Gaugefields_4D_B200, so that every generic algorithm of
#     Gaugefields.jl written in those primitives (evaluate_gaugelinks!, calc_dSdUμ!, measurements, custom actions) dispatches
#     to the library instead of the error fall-backs of src/AbstractGaugefields.jl:1631-1724, 2673-3064.
#
# The Python module gaugefields.jl_b200/gfb200/ binds exactly the same symbols through ctypes and is what tests/ exercise.

const LIBGFB200 = get(ENV, "GFB200_LIB", joinpath(@__DIR__, "..", "libgfb200.so"))

"""
    B200Backend(; gpus=1, devices=nothing)

Select the hand-written sm_100a CUDA implementation.  `gpus` local B200s are driven from this process; the 4D lattice is
split into contiguous t-slabs internally (the field reports `process_grid = (1,1,1,1)` to Julia).  There is no CPU
fallback: construction fails with an `ErrorException` when no GPU is usable.
"""
struct B200Backend <: AbstractGaugeBackend
    gpus::Int
    devices::Union{Nothing,Vector{Cint}}
    B200Backend(; gpus::Integer=1, devices=nothing) =
        new(Int(gpus), devices === nothing ? nothing : Cint.(collect(devices)))
end

# ---- error convention: status 1 -> ArgumentError (molecular_dynamics.jl:447-465), others -> ErrorException
mutable struct B200Context
    ptr::Ptr{Cvoid}
end
const _B200_CONTEXTS = Dict{Tuple{Int,Any},B200Context}()

function _gfb_check(status::Cint, ctx::Ptr{Cvoid}=C_NULL)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:gfb_last_error, LIBGFB200), Cstring, (Ptr{Cvoid},), ctx))
    status == 1 && throw(ArgumentError(msg))
    error("libgfb200 status $status: $msg")
end

# Finalizers run in any order at exit.  gfb_finalize releases the device memory of every live handle and orphans it, and the
# gfb_*_free functions only delete the host struct of an orphaned handle (csrc/api.cu, gfb_finalize), so "context first" is safe.
function _b200_context(backend::B200Backend)
    get!(_B200_CONTEXTS, (backend.gpus, backend.devices)) do
        out = Ref{Ptr{Cvoid}}(C_NULL)
        devs = backend.devices === nothing ? Ptr{Cint}(C_NULL) : pointer(backend.devices)
        GC.@preserve backend _gfb_check(ccall((:gfb_init, LIBGFB200), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), backend.gpus, devs, out))
        ctx = B200Context(out[])
        finalizer(c -> ccall((:gfb_finalize, LIBGFB200), Cint, (Ptr{Cvoid},), c.ptr), ctx)
        ctx
    end
end

# ---- handles ---------------------------------------------------------------------------------------
# gfb_gauge: all four directions of a configuration (structure-of-arrays, DESIGN.md "Data layout in HBM");
# gfb_field: ONE 3x3 matrix field -- a view of direction mu of a configuration (no copy) or a temporary of its own;
# gfb_mom:   the 4 x 8 algebra coefficients.
mutable struct B200GaugeHandle
    ptr::Ptr{Cvoid}
    ctx::B200Context
end
mutable struct B200FieldHandle
    ptr::Ptr{Cvoid}
    ctx::B200Context
    owner::Any          # the B200GaugeHandle a view aliases (keeps it alive), nothing for a temporary
end
mutable struct B200MomHandle
    ptr::Ptr{Cvoid}
    ctx::B200Context
end

# `U[μ]` and `similar(U[1])` are both Gaugefields_4D_B200: `conf` is the configuration a view belongs to (nothing for a
# temporary), `field` the gfb_field every primitive operates on.
struct Gaugefields_4D_B200{NC} <: Gaugefields_4D{NC}
    conf::Union{Nothing,B200GaugeHandle}
    field::B200FieldHandle
    mu::Int
    NX::Int
    NY::Int
    NZ::Int
    NT::Int
    NDW::Int
    NV::Int
    NC::Int
    verbose_print::Verbose_print
end

struct TA_Gaugefields_4D_B200{NC,NumofBasis} <: TA_Gaugefields_4D{NC}
    handle::B200MomHandle
    mu::Int
    NX::Int
    NY::Int
    NZ::Int
    NT::Int
    NC::Int
    NumofBasis::Int
end

Base.eltype(::Gaugefields_4D_B200) = ComplexF64
const B200Links = AbstractVector{<:Gaugefields_4D_B200}
const B200Momenta = AbstractVector{<:TA_Gaugefields_4D_B200}

_b200_dims(u::Union{Gaugefields_4D_B200,TA_Gaugefields_4D_B200}) = (u.NX, u.NY, u.NZ, u.NT)
_b200_ctx(u::Gaugefields_4D_B200) = u.field.ctx
function _b200_handle(U::B200Links)
    h = first(U).conf
    h === nothing && throw(ArgumentError("this operation needs a gauge configuration (the Vector returned by gauge_configuration / similar(U)), not temporaries"))
    return h
end
_b200_handle(P::B200Momenta) = first(P).handle

function _b200_alloc_gauge(ctx::B200Context, dims::NTuple{4,Int})
    out = Ref{Ptr{Cvoid}}(C_NULL)
    _gfb_check(ccall((:gfb_gauge_alloc, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
        ctx.ptr, dims[1], dims[2], dims[3], dims[4], out), ctx.ptr)
    h = B200GaugeHandle(out[], ctx)
    finalizer(x -> ccall((:gfb_gauge_free, LIBGFB200), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end
function _b200_view(h::B200GaugeHandle, mu::Int)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    _gfb_check(ccall((:gfb_field_view, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}), h.ptr, mu - 1, out), h.ctx.ptr)
    f = B200FieldHandle(out[], h.ctx, h)
    finalizer(x -> ccall((:gfb_field_free, LIBGFB200), Cint, (Ptr{Cvoid},), x.ptr), f)
    return f
end
function _b200_views(h::B200GaugeHandle, dims::NTuple{4,Int}, verbose::Int)
    vp = Verbose_print(verbose)
    return [Gaugefields_4D_B200{3}(h, _b200_view(h, mu), mu, dims[1], dims[2], dims[3], dims[4], 1, prod(dims), 3, vp) for mu = 1:4]
end

# ---- gauge_configuration: the branch beside src/API.jl:202-220 -----------------------------------------
function _gauge_configuration_b200(backend::B200Backend, dimensions, colors, start, seed, rng, verbose)
    length(dimensions) == 4 || throw(ArgumentError("B200Backend supports 4 dimensions; got $(length(dimensions))"))
    colors == 3 || throw(ArgumentError("B200Backend supports colors=3; got $colors"))
    rng isa Philox4x32 || throw(ArgumentError("B200Backend implements the Philox4x32 site RNG"))
    ctx = _b200_context(backend)
    dims = Tuple(Int.(dimensions))
    h = _b200_alloc_gauge(ctx, dims)
    if start == :cold
        _gfb_check(ccall((:gfb_set_cold, LIBGFB200), Cint, (Ptr{Cvoid},), h.ptr), ctx.ptr)
    else
        s = seed === nothing ? rand(UInt64) : UInt64(seed)
        _gfb_check(ccall((:gfb_set_hot, LIBGFB200), Cint, (Ptr{Cvoid}, UInt64, Cint), h.ptr, s, 0), ctx.ptr)
    end
    return _b200_views(h, dims, Int(verbose))
end
# In gauge_configuration (src/API.jl:202) add, before the LatticeMatricesBackend branch:
#     backend isa B200Backend && return _gauge_configuration_b200(backend, dimensions, colors, start, seed, rng, verbose)

# similar(U): a new configuration;  similar(U[μ]): one temporary matrix field (src/AbstractGaugefields.jl:631-645)
function Base.similar(U::B200Links)
    u = first(U)
    return _b200_views(_b200_alloc_gauge(_b200_ctx(u), _b200_dims(u)), _b200_dims(u), 0)
end
function Base.similar(u::Gaugefields_4D_B200{NC}) where {NC}
    ctx = _b200_ctx(u)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    _gfb_check(ccall((:gfb_field_alloc, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
        ctx.ptr, u.NX, u.NY, u.NZ, u.NT, out), ctx.ptr)
    f = B200FieldHandle(out[], ctx, nothing)
    finalizer(x -> ccall((:gfb_field_free, LIBGFB200), Cint, (Ptr{Cvoid},), x.ptr), f)
    return Gaugefields_4D_B200{NC}(nothing, f, 0, u.NX, u.NY, u.NZ, u.NT, u.NDW, u.NV, u.NC, u.verbose_print)
end

# copy_configuration! (src/API.jl:307-322) lands here through substitute_U! on the vectors
function substitute_U!(dst::B200Links, src::B200Links)
    _gfb_check(ccall((:gfb_gauge_copy, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), _b200_handle(dst).ptr, _b200_handle(src).ptr),
        _b200_handle(dst).ctx.ptr)
    return dst
end

# normalize_U! (src/4D/nowing/gaugefields_4D_nowing.jl:2387-2458).  Needed after loading a single-precision configuration: the
# fused passes use their two-row SU(3) products only on configurations that are unitary to 1e-12 (checked once per upload)
# and fall back to full 3x3 products otherwise (INTEGRATION.md, "Contract on the link field").
function normalize_U!(U::B200Links)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_reunitarize, LIBGFB200), Cint, (Ptr{Cvoid},), h.ptr), h.ctx.ptr)
    return U
end

# host <-> device in the gathered layout used by save_configuration/load_configuration (src/API.jl:516-529, 625-629).
# The C side takes `double*`; ComplexF64 is two Float64, so the array is passed as Ptr{ComplexF64} (same address).
function gather_global_array(u::Gaugefields_4D_B200)
    A = Array{ComplexF64}(undef, 3, 3, u.NX, u.NY, u.NZ, u.NT)
    _gfb_check(ccall((:gfb_field_download, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), u.field.ptr, A), _b200_ctx(u).ptr)
    return A
end
function scatter_global_array!(u::Gaugefields_4D_B200, A::Array{ComplexF64,6})
    size(A) == (3, 3, u.NX, u.NY, u.NZ, u.NT) || throw(DimensionMismatch("expected ComplexF64[3,3,NX,NY,NZ,NT]"))
    _gfb_check(ccall((:gfb_field_upload, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{ComplexF64}), u.field.ptr, A), _b200_ctx(u).ptr)
    return u
end
# whole-configuration transfers (one staged transpose per direction on the device)
function gather_global_array(U::B200Links, mu::Integer)
    u = first(U)
    A = Array{ComplexF64}(undef, 3, 3, u.NX, u.NY, u.NZ, u.NT)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_gauge_download, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}), h.ptr, mu - 1, A), h.ctx.ptr)
    return A
end
function scatter_global_array!(U::B200Links, mu::Integer, A::Array{ComplexF64,6})
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_gauge_upload, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Ptr{ComplexF64}), h.ptr, mu - 1, A), h.ctx.ptr)
    return U
end
# ILDG binary payload (the ildg-binary-data record as it is in the file: big-endian [t][z][y][x][mu][row][col], precision 32
# or 64) <-> device; replaces the site-by-site host loops of load_gaugefield! / _save_binarydata
# (src/output/ildg_format.jl:67-83, 697-746).  Byte swap, precision conversion and transpose run on the GPU.
function scatter_ildg_payload!(U::B200Links, payload::Vector{UInt8}, precision::Integer)
    u = first(U)
    length(payload) == u.NX * u.NY * u.NZ * u.NT * 4 * 9 * 2 * (precision ÷ 8) ||
        throw(DimensionMismatch("ILDG payload has $(length(payload)) bytes"))
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_gauge_upload_ildg, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint), h.ptr, payload, precision), h.ctx.ptr)
    precision == 32 && normalize_U!(U)     # unitary to 1e-7 only: put it back on the group (and on the two-row fast path)
    return U
end
function gather_ildg_payload(U::B200Links, precision::Integer)
    u = first(U)
    payload = Vector{UInt8}(undef, u.NX * u.NY * u.NZ * u.NT * 4 * 9 * 2 * (precision ÷ 8))
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_gauge_download_ildg, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint), h.ptr, payload, precision), h.ctx.ptr)
    return payload
end

# slow-path element access (generic code and tests index fields directly; gaugefields_4D_MPILattice.jl:152-197): one field
# round trip per access -- for tests and set-up code, not for loops over sites
Base.getindex(u::Gaugefields_4D_B200, i, j, x, y, z, t) = gather_global_array(u)[i, j, x, y, z, t]
function Base.setindex!(u::Gaugefields_4D_B200, v, i, j, x, y, z, t)
    A = gather_global_array(u)
    A[i, j, x, y, z, t] = v
    scatter_global_array!(u, A)
    return v
end
getvalue(u::Gaugefields_4D_B200, i, j, x, y, z, t) = u[i, j, x, y, z, t]
setvalue!(u::Gaugefields_4D_B200, v, i, j, x, y, z, t) = (u[i, j, x, y, z, t] = v)
set_wing_U!(u::Gaugefields_4D_B200) = nothing      # halos are exchanged inside the library when a shifted read needs them
set_wing_U!(U::B200Links) = nothing

# ---- momenta: the method beside the if-chain of src/TA_Gaugefields.jl:151-195 -------------------------------
function initialize_TA_Gaugefields(U::B200Links)
    u = first(U)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    ctx = _b200_ctx(u)
    _gfb_check(ccall((:gfb_mom_alloc, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
        ctx.ptr, u.NX, u.NY, u.NZ, u.NT, out), ctx.ptr)
    h = B200MomHandle(out[], ctx)
    finalizer(x -> ccall((:gfb_mom_free, LIBGFB200), Cint, (Ptr{Cvoid},), x.ptr), h)
    return [TA_Gaugefields_4D_B200{3,8}(h, mu, u.NX, u.NY, u.NZ, u.NT, 3, 8) for mu = 1:4]
end

# gaussian_momenta! (src/API.jl:341-368) dispatches on the element type; add this method
function gaussian_momenta!(P::B200Momenta; sigma=1.0, seed=nothing, sweep::Integer=0, rng::SiteRNGAlgorithm=Philox4x32())
    sweep >= 0 || throw(ArgumentError("sweep must be nonnegative; got $sweep"))
    h = _b200_handle(P)
    s = seed === nothing ? rand(UInt64) : UInt64(seed)
    _gfb_check(ccall((:gfb_gaussian_momenta, LIBGFB200), Cint, (Ptr{Cvoid}, UInt64, UInt64, Cdouble, Cint), h.ptr, s, UInt64(sweep), Float64(sigma), 0), h.ctx.ptr)
    return P
end

# p * p (src/TA_Gaugefields.jl:127-137)
function Base.:*(x::B200Momenta, y::B200Momenta)
    _b200_handle(x) === _b200_handle(y) || error("B200Backend provides p*p (kinetic energy); use add_U! for combinations")
    out = Ref{Cdouble}(0)
    _gfb_check(ccall((:gfb_kinetic, LIBGFB200), Cint, (Ptr{Cvoid}, Ref{Cdouble}), _b200_handle(x).ptr, out), _b200_handle(x).ctx.ptr)
    return out[]
end
# clear_U!(P), add_U!(P, t, F) on all directions and on one (TA_gaugefields_4D_MPILattice.jl:285-291, :150-173): the operations
# an external action provider (a fermion force, say) uses to add its force into this backend's momenta
function clear_U!(P::B200Momenta)
    h = _b200_handle(P)
    _gfb_check(ccall((:gfb_mom_zero, LIBGFB200), Cint, (Ptr{Cvoid},), h.ptr), h.ctx.ptr)
    return P
end
function add_U!(P::B200Momenta, t::Number, F::B200Momenta)
    h = _b200_handle(P)
    _gfb_check(ccall((:gfb_mom_axpy, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cvoid}), h.ptr, Float64(t), _b200_handle(F).ptr), h.ctx.ptr)
    return P
end
function add_U!(p::TA_Gaugefields_4D_B200, t::Number, f::TA_Gaugefields_4D_B200)
    p.mu == f.mu || throw(ArgumentError("add_U! on momenta of different directions"))
    _gfb_check(ccall((:gfb_mom_axpy_dir, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cvoid}), p.handle.ptr, p.mu - 1, Float64(t), f.handle.ptr), p.handle.ctx.ptr)
    return p
end
function substitute_U!(dst::B200Momenta, src::B200Momenta)
    h = _b200_handle(dst)
    _gfb_check(ccall((:gfb_mom_copy, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), h.ptr, _b200_handle(src).ptr), h.ctx.ptr)
    return dst
end

# ---- observables ------------------------------------------------------------------------------------------
function calculate_Plaquette(U::B200Links, temp1=nothing, temp2=nothing)
    out = Ref{Cdouble}(0)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_plaquette_sum, LIBGFB200), Cint, (Ptr{Cvoid}, Ref{Cdouble}), h.ptr, out), h.ctx.ptr)
    return out[]
end
function calculate_Polyakov_loop(U::B200Links, temp1=nothing, temp2=nothing)
    out = zeros(Cdouble, 2)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_polyakov, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h.ptr, out), h.ctx.ptr)
    return complex(out[1], out[2])
end
# clover / plaquette energy density E(t) (samples/measurements/energydensity.jl:4-78)
function b200_energy_density(U::B200Links; kind::Symbol=:clover)
    out = Ref{Cdouble}(0)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_energy_density, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Ref{Cdouble}), h.ptr, kind == :clover ? 0 : 1, out), h.ctx.ptr)
    return out[]
end
# topological_charge / topological_charge_density (src/AbstractGaugefields.jl:1447-1490)
const _B200_TOPO = Dict(:plaquette => 0, :clover => 1, :improved => 2)
function topological_charge(U::B200Links; method=:plaquette)
    haskey(_B200_TOPO, method) || throw(ArgumentError("supported topological_charge methods are :plaquette, :clover, and :improved"))
    out = Ref{Cdouble}(0)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_topological_charge, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Ref{Cdouble}), h.ptr, _B200_TOPO[method], out), h.ctx.ptr)
    return out[]
end
function topological_charge_density(U::B200Links; method=:plaquette)
    haskey(_B200_TOPO, method) || throw(ArgumentError("supported topological_charge_density methods are :plaquette, :clover, and :improved"))
    u = first(U)
    density = zeros(Float64, u.NX, u.NY, u.NZ, u.NT)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_topological_charge_density, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), h.ptr, _B200_TOPO[method], density), h.ctx.ptr)
    return density
end

# ---- actions: (c_plaq, c_rect) of a GaugeAction ------------------------------------------------------------
# A GaugeAction term is (β, closedloops, staples) (GaugeAction_dataset, src/action/GaugeActions.jl:14-62).  The fused kernels
# cover terms whose loop set is plaquette ∪ plaquette' (12 four-link loops) or rectangular ∪ rectangular' (24 six-link loops),
# i.e. what `push!(action, c, vcat(loops, loops'))` stores for make_loops_fromname("plaquette" | "rectangular").
_b200_nlinks(w) = hasproperty(w, :glinks) ? length(getproperty(w, :glinks)) : length(w)
function _is_loopset(closedloops, nlinks::Int, nloops::Int)
    length(closedloops) == nloops || return false
    return all(w -> _b200_nlinks(w) == nlinks, closedloops)
end
_is_plaquette_pair(closedloops) = _is_loopset(closedloops, 4, 12)
_is_rectangle_pair(closedloops) = _is_loopset(closedloops, 6, 24)
function _action_coefficients(action::GaugeAction)
    cp = 0.0
    cr = 0.0
    for term in action.dataset
        imag(term.β) == 0 || error("B200Backend fuses real loop coefficients only")
        if _is_plaquette_pair(term.closedloops)
            cp += real(term.β)
        elseif _is_rectangle_pair(term.closedloops)
            cr += real(term.β)
        else
            error("B200Backend fuses plaquette+plaquette' and rectangular+rectangular' terms; compose other loop sets from the primitive table (evaluate_gaugelinks! works on this backend)")
        end
    end
    return cp, cr
end
function _action_coefficients(actions::MDActionSet, names=keys(actions.terms))
    cp = 0.0
    cr = 0.0
    for name in names
        a, b = _action_coefficients(getproperty(actions.terms, name))
        cp += a
        cr += b
    end
    return cp, cr
end
_group_names(::MDForceGroup{Names}) where {Names} = Names

# ---- molecular dynamics: fused overrides of src/molecular_dynamics.jl ------------------------------------------
function md_potential(action::GaugeAction, U::B200Links, workspace)
    cp, cr = _action_coefficients(action)
    out = zeros(Cdouble, 2)
    h = _b200_handle(U)
    if cr == 0
        _gfb_check(ccall((:gfb_plaquette_sum, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h.ptr, out), h.ctx.ptr)
    else
        _gfb_check(ccall((:gfb_loop_sums, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h.ptr, out), h.ctx.ptr)
    end
    return -(2 / 3) * (cp * out[1] + cr * out[2])     # -(1/NC) Re evaluate_GaugeAction, loops + adjoints
end

function _b200_force!(force::B200Momenta, U::B200Links, cp::Float64, cr::Float64)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_force_general, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble), _b200_handle(force).ptr, h.ptr, cp, cr), h.ctx.ptr)
    return nothing
end
md_force!(force::B200Momenta, action::GaugeAction, U::B200Links, workspace) = _b200_force!(force, U, _action_coefficients(action)...)
# MDActionSet of gauge actions: the member forces add, so the selected members' coefficients add (one kernel instead of one
# md_force! + add_U! per member, molecular_dynamics.jl:128-236).  Sets with non-gauge members take the generic path, which works
# through md_force!(::GaugeAction) above and add_U!/clear_U! on the momenta.
function md_force!(force::B200Momenta, actions::MDActionSet, U::B200Links, workspace)
    all(a -> a isa GaugeAction, values(actions.terms)) || return invoke(md_force!, Tuple{Any,MDActionSet,Any,Any}, force, actions, U, workspace)
    return _b200_force!(force, U, _action_coefficients(actions)...)
end
function md_force!(force::B200Momenta, actions::MDActionSet, U::B200Links, workspace, group::MDForceGroup)
    names = _group_names(group)
    all(n -> getproperty(actions.terms, n) isa GaugeAction, names) ||
        return invoke(md_force!, Tuple{Any,MDActionSet,Any,Any,MDForceGroup}, force, actions, U, workspace, group)
    return _b200_force!(force, U, _action_coefficients(actions, names)...)
end

function _b200_kick!(P::B200Momenta, U::B200Links, step_size, cp::Float64, cr::Float64)
    isfinite(step_size) || throw(ArgumentError("the momentum step size must be finite; got $step_size"))
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_update_momenta_general, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cdouble),
        _b200_handle(P).ptr, h.ptr, Float64(step_size), cp, cr), h.ctx.ptr)
    return P
end
_b200_gauge_only(action::GaugeAction) = true
_b200_gauge_only(actions::MDActionSet) = all(a -> a isa GaugeAction, values(actions.terms))
_b200_gauge_only(action) = false
function update_momenta!(P::B200Momenta, U::B200Links, step_size, driver::MDDriver)
    _b200_gauge_only(driver.action) || return invoke(update_momenta!, Tuple{Any,Any,Any,MDDriver}, P, U, step_size, driver)
    return _b200_kick!(P, U, step_size, _action_coefficients(driver.action)...)
end
function update_momenta!(P::B200Momenta, U::B200Links, step_size, driver::MDDriver, group::MDForceGroup)
    names = _group_names(group)
    (driver.action isa MDActionSet && all(n -> getproperty(driver.action.terms, n) isa GaugeAction, names)) ||
        return invoke(update_momenta!, Tuple{Any,Any,Any,MDDriver,MDForceGroup}, P, U, step_size, driver, group)
    return _b200_kick!(P, U, step_size, _action_coefficients(driver.action, names)...)
end

function update_gaugefields!(U::B200Links, P::B200Momenta, step_size, driver::MDDriver)
    isfinite(step_size) || throw(ArgumentError("the gauge-field step size must be finite; got $step_size"))
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_update_links, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble), h.ptr, _b200_handle(P).ptr, Float64(step_size)), h.ctx.ptr)
    return U
end

function md_hamiltonian(U::B200Links, p::B200Momenta, driver::MDDriver)
    _b200_gauge_only(driver.action) || return invoke(md_hamiltonian, Tuple{Any,Any,MDDriver}, U, p, driver)
    cp, cr = _action_coefficients(driver.action)
    out = Ref{Cdouble}(0)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_hamiltonian_general, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Ref{Cdouble}),
        h.ptr, _b200_handle(p).ptr, cp, cr, out), h.ctx.ptr)
    return out[]
end

# md_trajectory! (src/molecular_dynamics.jl:712-730) in ONE library call for QPQ / PQP on gauge actions: `steps` fused kick+drift
# launches.  SextonWeingarten and custom integrators run the reference's generic loop over md_step!, whose elementary
# update_momenta! / update_gaugefields! calls are the overrides above (one kernel each).
# GFB200_FUSED=0 replays the reference's op sequence (link, kick, link) for bit-level comparisons of the ordering.
function md_trajectory!(U::B200Links, p::B200Momenta, driver::MDDriver; diagnostics::Bool=true)
    fusable = (driver.integrator isa QPQ || driver.integrator isa PQP) && _b200_gauge_only(driver.action)
    fusable || return invoke(md_trajectory!, Tuple{Any,Any,MDDriver}, U, p, driver; diagnostics=diagnostics)
    integ = driver.integrator isa QPQ ? 0 : 1
    cp, cr = _action_coefficients(driver.action)
    H = zeros(Cdouble, 2)
    h = _b200_handle(U)
    fused = get(ENV, "GFB200_FUSED", "1") == "1" ? 1 : 0
    GC.@preserve H _gfb_check(ccall((:gfb_md_trajectory_general, LIBGFB200), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble, Cint, Cdouble, Cint, Cint, Ptr{Cdouble}),
        h.ptr, _b200_handle(p).ptr, cp, cr, driver.steps, Float64(driver.trajectory_length), integ, fused,
        diagnostics ? pointer(H) : Ptr{Cdouble}(C_NULL)), h.ctx.ptr)
    diagnostics || return nothing
    return (initial_hamiltonian=H[1], final_hamiltonian=H[2], delta_hamiltonian=H[2] - H[1])
end

# ---- gradient flow and stout ----------------------------------------------------------------------------------
function flow!(U::B200Links, g::Gradientflow)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_flow, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, Cint), h.ptr, Float64(g.eps), g.Nflow), h.ctx.ptr)
    return U
end
# Gradientflow_general (src/smearing/gradientflow.jl:33-116, 240-316): its gaugeaction holds (value, loops ∪ loops') terms
function flow!(U::B200Links, g::Gradientflow_general)
    cp, cr = _action_coefficients(g.gaugeaction)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_flow_general, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, Cint, Cdouble, Cdouble), h.ptr, Float64(g.eps), g.Nflow, cp, cr), h.ctx.ptr)
    return U
end
# add_force!(F, U; plaqonly=true) after clear_U!(F) and exp_aF_U!(W, a, F, U): the two halves of a flow stage
# (src/AbstractGaugefields.jl:2717-2762, 2810-2841), for user code that composes its own integrator
function b200_flow_force!(F::B200Momenta, U::B200Links)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_flow_force, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), _b200_handle(F).ptr, h.ptr), h.ctx.ptr)
    return F
end
function exp_aF_U!(W::B200Links, a::Number, F::B200Momenta, U::B200Links, temps=nothing)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_exp_aF_U, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}), _b200_handle(W).ptr, Float64(a), _b200_handle(F).ptr, h.ptr), h.ctx.ptr)
    return W
end

function forward!(s::STOUT_Layer, Uout::B200Links, ρs::Vector{<:Number}, Uin::B200Links)
    length(ρs) == 1 || error("B200Backend implements the plaquette-staple stout layer with one ρ")
    h = _b200_handle(Uin)
    _gfb_check(ccall((:gfb_stout_forward, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
        _b200_handle(Uout).ptr, h.ptr, Float64(real(ρs[1])), C_NULL), h.ctx.ptr)
    substitute_U!(s.Uin, Uin)   # the backward pass recomputes C, Q and exp(Q) from the layer input (no tape in HBM)
    s.ρs[1] = ρs[1]
    return
end

function layer_pullback!(δ_prev::B200Links, δ_current::B200Links, layer::STOUT_Layer, Uprev::B200Links, temps, tempf)
    h = _b200_handle(Uprev)
    _gfb_check(ccall((:gfb_stout_backward, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble),
        _b200_handle(δ_prev).ptr, _b200_handle(δ_current).ptr, h.ptr, Float64(real(layer.ρs[1]))), h.ctx.ptr)
    return
end
# calc_dSdUμ! for all four directions at once and the kick from an explicit derivative field
# (GaugeActions.jl:95-123, molecular_dynamics.jl:255-265; the stout-HMC script test/HMCstout_test_nowing.jl:99-118)
function b200_wilson_dSdU!(dSdU::B200Links, action::GaugeAction, U::B200Links)
    cp, cr = _action_coefficients(action)
    cr == 0 || error("b200_wilson_dSdU! takes a plaquette action")
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_wilson_dSdU, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble), _b200_handle(dSdU).ptr, h.ptr, 2 * cp), h.ctx.ptr)
    return dSdU
end
function b200_kick_from_dSdU!(P::B200Momenta, U::B200Links, dSdU::B200Links, factor::Number)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_kick_from_dSdU, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble),
        _b200_handle(P).ptr, h.ptr, _b200_handle(dSdU).ptr, Float64(factor)), h.ctx.ptr)
    return P
end

# ---- heatbath / overrelaxation: heatbath!(U, h::Heatbath), overrelaxation!(U, h) (src/heatbath/heatbathmodule.jl:843-852) --------
# One sweep = 4 directions x 2 checkerboard colours in the library; `h.sweep` / `h.overrelaxation_sweep` key the streams and advance
# as in the reference (Heatbath fields β, seed, sweep, overrelaxation_sweep: heatbathmodule.jl:55-99).
function heatbath!(U::B200Links, h::Heatbath)
    g = _b200_handle(U)
    _gfb_check(ccall((:gfb_heatbath, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, UInt64, UInt64, Cint), g.ptr, Float64(h.β), UInt64(h.seed), UInt64(h.sweep), 0), g.ctx.ptr)
    h.sweep += 1
    return U
end
function overrelaxation!(U::B200Links, h::Heatbath)
    g = _b200_handle(U)
    _gfb_check(ccall((:gfb_overrelaxation, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, UInt64, UInt64, Cint), g.ptr, Float64(h.β), UInt64(h.seed), UInt64(h.overrelaxation_sweep), 0), g.ctx.ptr)
    h.overrelaxation_sweep += 1
    return U
end

# ---- primitive table on Gaugefields_4D_B200 ---------------------------------------------------------------------
# Lazy shift / adjoint views mirror Shifted_/Adjoint_Gaugefields_4D_MPILattice (gaugefields_4D_MPILattice.jl:647-689): the
# library's gfb_mul / gfb_field_copy take the shift and the dagger flag of each operand, so no shifted copy is ever made.
struct B200Lazy{NC}
    parent::Gaugefields_4D_B200{NC}
    shift::NTuple{4,Cint}
    dagger::Bool
end
const _B200_NOSHIFT = (Cint(0), Cint(0), Cint(0), Cint(0))
_lazy(u::Gaugefields_4D_B200{NC}) where {NC} = B200Lazy{NC}(u, _B200_NOSHIFT, false)
_lazy(l::B200Lazy) = l
Base.adjoint(u::Gaugefields_4D_B200{NC}) where {NC} = B200Lazy{NC}(u, _B200_NOSHIFT, true)
Base.adjoint(l::B200Lazy{NC}) where {NC} = B200Lazy{NC}(l.parent, l.shift, !l.dagger)
function shift_U(u::Union{Gaugefields_4D_B200,B200Lazy}, s::NTuple{4,<:Integer})
    l = _lazy(u)
    return typeof(l)(l.parent, (Cint(l.shift[1] + s[1]), Cint(l.shift[2] + s[2]), Cint(l.shift[3] + s[3]), Cint(l.shift[4] + s[4])), l.dagger)
end
function shift_U(u::Union{Gaugefields_4D_B200,B200Lazy}, ν::Integer)     # ν = ±1..±4 (gaugefields_4D_MPILattice.jl:647-670)
    1 <= abs(ν) <= 4 || throw(ArgumentError("shift direction must be ±1..±4; got $ν"))
    return shift_U(u, ntuple(d -> d == abs(ν) ? sign(ν) : 0, 4))
end
const B200Operand = Union{Gaugefields_4D_B200,B200Lazy}

# mul!(C, A, B[, α, β]): C = α op(A(x+sA)) op(B(x+sB)) + β C  (src/AbstractGaugefields.jl:2082-2105, 2906-2914)
function LinearAlgebra.mul!(C::Gaugefields_4D_B200, A::B200Operand, B::B200Operand, α::Number=1, β::Number=0)
    a, b = _lazy(A), _lazy(B)
    sa, sb = collect(a.shift), collect(b.shift)
    ctx = _b200_ctx(C)
    GC.@preserve sa sb _gfb_check(ccall((:gfb_mul, LIBGFB200), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cint}, Cint, Ptr{Cvoid}, Ptr{Cint}, Cint, Cdouble, Cdouble, Cdouble, Cdouble),
        C.field.ptr, a.parent.field.ptr, sa, a.dagger ? 1 : 0, b.parent.field.ptr, sb, b.dagger ? 1 : 0,
        Float64(real(α)), Float64(imag(α)), Float64(real(β)), Float64(imag(β))), ctx.ptr)
    return C
end
# add_U!(C, α, A | A') / add_U!(C, A)  (gaugefields_4D_MPILattice.jl:739-772)
function add_U!(C::Gaugefields_4D_B200, α::Number, A::B200Operand)
    a = _lazy(A)
    all(iszero, a.shift) || throw(ArgumentError("add_U! takes plain or adjoint operands"))
    _gfb_check(ccall((:gfb_axpy, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cvoid}, Cint),
        C.field.ptr, Float64(real(α)), Float64(imag(α)), a.parent.field.ptr, a.dagger ? 1 : 0), _b200_ctx(C).ptr)
    return C
end
add_U!(C::Gaugefields_4D_B200, A::B200Operand) = add_U!(C, 1.0, A)
function clear_U!(C::Gaugefields_4D_B200)
    _gfb_check(ccall((:gfb_field_clear, LIBGFB200), Cint, (Ptr{Cvoid},), C.field.ptr), _b200_ctx(C).ptr)
    return C
end
function unit_U!(C::Gaugefields_4D_B200)
    _gfb_check(ccall((:gfb_field_unit, LIBGFB200), Cint, (Ptr{Cvoid},), C.field.ptr), _b200_ctx(C).ptr)
    return C
end
# substitute_U!(A, B | shifted B | B')  (gaugefields_4D_MPILattice.jl:509-575)
function substitute_U!(A::Gaugefields_4D_B200, B::B200Operand)
    b = _lazy(B)
    s = collect(b.shift)
    GC.@preserve s _gfb_check(ccall((:gfb_field_copy, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cint}, Cint),
        A.field.ptr, b.parent.field.ptr, s, b.dagger ? 1 : 0), _b200_ctx(A).ptr)
    return A
end
# tr(A), tr(A, B) = Σ_x tr(A(x) B(x))  (gaugefields_4D_MPILattice.jl:721-728)
function LinearAlgebra.tr(A::Gaugefields_4D_B200)
    out = zeros(Cdouble, 2)
    _gfb_check(ccall((:gfb_tr, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), A.field.ptr, out), _b200_ctx(A).ptr)
    return complex(out[1], out[2])
end
function LinearAlgebra.tr(A::Gaugefields_4D_B200, B::Gaugefields_4D_B200)
    out = zeros(Cdouble, 2)
    _gfb_check(ccall((:gfb_tr2, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}), A.field.ptr, B.field.ptr, out), _b200_ctx(A).ptr)
    return complex(out[1], out[2])
end
# Traceless_antihermitian!(Q, M) matrix -> matrix (:774-781); Traceless_antihermitian_add!(P[μ], factor, M) matrix -> 8
# coefficients (TA_gaugefields_4D_MPILattice.jl:263-283)
function Traceless_antihermitian!(Q::Gaugefields_4D_B200, M::Gaugefields_4D_B200)
    _gfb_check(ccall((:gfb_ta_project, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), Q.field.ptr, M.field.ptr), _b200_ctx(Q).ptr)
    return Q
end
function Traceless_antihermitian_add!(P::TA_Gaugefields_4D_B200, factor::Number, M::Gaugefields_4D_B200)
    _gfb_check(ccall((:gfb_ta_coeffs_add, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cvoid}), P.handle.ptr, P.mu - 1, Float64(factor), M.field.ptr), _b200_ctx(M).ptr)
    return P
end
# exptU!(E, t, Q[, temps]) with Q a matrix field (:798-808) or the momenta of one direction (TA_...:196-210)
function exptU!(E::Gaugefields_4D_B200, t::Number, Q::Gaugefields_4D_B200, temps=nothing)
    _gfb_check(ccall((:gfb_exp, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cvoid}), E.field.ptr, Float64(t), Q.field.ptr), _b200_ctx(E).ptr)
    return E
end
function exptU!(E::Gaugefields_4D_B200, t::Number, P::TA_Gaugefields_4D_B200, temps=nothing)
    _gfb_check(ccall((:gfb_exp_mom, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint), E.field.ptr, Float64(t), P.handle.ptr, P.mu - 1), _b200_ctx(E).ptr)
    return E
end
