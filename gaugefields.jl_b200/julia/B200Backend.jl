# B200Backend.jl -- the Julia side of the drop-in boundary (source only: the build image has no Julia).
#
# Include this file from Gaugefields.jl (after src/API.jl and src/molecular_dynamics.jl are loaded):
#
#     include("B200Backend.jl")          # inside module Gaugefields
#     U = gauge_configuration((32,32,32,32); backend=B200Backend(), start=:hot, seed=0x1234)
#
# It adds a third backend tag next to LatticeMatricesBackend/LegacyBackend (src/API.jl:6-27), two field
# types that hold opaque device handles, and methods of the existing generic functions that forward
# to libgfb200.so through `ccall`.  User scripts written against the v1 API (docs/src/hmc.md:128-190,
# docs/src/highlevelapi.md) run unchanged: only the `backend=` keyword differs.
#
# Every ccall below binds one symbol of include/gfb200.h; the Python module gaugefields.jl_b200/gfb200/
# binds exactly the same symbols through ctypes and is what the tests in this repository exercise.

const LIBGFB200 = get(ENV, "GFB200_LIB", joinpath(@__DIR__, "..", "libgfb200.so"))

"""
    B200Backend(; gpus=1, devices=nothing)

Select the hand-written sm_100a CUDA implementation.  `gpus` local B200s are driven from this
process; the 4D lattice is split into contiguous t-slabs internally (the field reports
`process_grid = (1,1,1,1)` to Julia).  There is no CPU fallback: construction fails with an
`ErrorException` when no GPU is usable.
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

function _b200_context(backend::B200Backend)
    get!(_B200_CONTEXTS, (backend.gpus, backend.devices)) do
        out = Ref{Ptr{Cvoid}}(C_NULL)
        devs = backend.devices === nothing ? C_NULL : pointer(backend.devices)
        _gfb_check(ccall((:gfb_init, LIBGFB200), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), backend.gpus, devs, out))
        ctx = B200Context(out[])
        finalizer(c -> ccall((:gfb_finalize, LIBGFB200), Cint, (Ptr{Cvoid},), c.ptr), ctx)
        ctx
    end
end

# ---- field types ---------------------------------------------------------------------------------
# One device object holds all four directions (structure-of-arrays, DESIGN.md "Data layout in HBM");
# the Vector returned by gauge_configuration holds four thin views of it so that `U[μ]`, `length(U)`,
# `similar(U)` and the property reads of src/API.jl:268-280 keep working.
mutable struct B200GaugeHandle
    ptr::Ptr{Cvoid}
    ctx::B200Context
end
mutable struct B200MomHandle
    ptr::Ptr{Cvoid}
    ctx::B200Context
end

struct Gaugefields_4D_B200{NC} <: Gaugefields_4D{NC}
    handle::B200GaugeHandle
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

function _b200_alloc_gauge(ctx::B200Context, dims::NTuple{4,Int})
    out = Ref{Ptr{Cvoid}}(C_NULL)
    _gfb_check(ccall((:gfb_gauge_alloc, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
        ctx.ptr, dims..., out), ctx.ptr)
    h = B200GaugeHandle(out[], ctx)
    finalizer(x -> ccall((:gfb_gauge_free, LIBGFB200), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end

function _b200_views(h::B200GaugeHandle, dims::NTuple{4,Int}, verbose::Int)
    vp = Verbose_print(verbose)
    return [Gaugefields_4D_B200{3}(h, mu, dims..., 1, prod(dims), 3, vp) for mu = 1:4]
end

_b200_dims(u::Union{Gaugefields_4D_B200,TA_Gaugefields_4D_B200}) = (u.NX, u.NY, u.NZ, u.NT)
_b200_handle(U::AbstractVector{<:Gaugefields_4D_B200}) = first(U).handle
_b200_handle(P::AbstractVector{<:TA_Gaugefields_4D_B200}) = first(P).handle

# ---- gauge_configuration: the branch beside src/API.jl:202-220 -----------------------------------------
function _gauge_configuration_b200(backend::B200Backend, dimensions, colors, start, seed, rng, verbose)
    length(dimensions) == 4 || throw(ArgumentError("B200Backend supports 4 dimensions; got $(length(dimensions))"))
    colors == 3 || throw(ArgumentError("B200Backend supports colors=3; got $colors"))
    rng isa Philox4x32 || throw(ArgumentError("B200Backend implements the Philox4x32 site RNG"))
    ctx = _b200_context(backend)
    h = _b200_alloc_gauge(ctx, Tuple(Int.(dimensions)))
    if start == :cold
        _gfb_check(ccall((:gfb_set_cold, LIBGFB200), Cint, (Ptr{Cvoid},), h.ptr), ctx.ptr)
    else
        s = seed === nothing ? rand(UInt64) : UInt64(seed)
        _gfb_check(ccall((:gfb_set_hot, LIBGFB200), Cint, (Ptr{Cvoid}, UInt64, Cint), h.ptr, s, 0), ctx.ptr)
    end
    return _b200_views(h, Tuple(Int.(dimensions)), Int(verbose))
end
# In gauge_configuration (src/API.jl:202) add, before the LatticeMatricesBackend branch:
#     backend isa B200Backend && return _gauge_configuration_b200(backend, dimensions, colors, start, seed, rng, verbose)

function Base.similar(U::AbstractVector{<:Gaugefields_4D_B200})
    u = first(U)
    return _b200_views(_b200_alloc_gauge(u.handle.ctx, _b200_dims(u)), _b200_dims(u), 0)
end

# copy_configuration! (src/API.jl:307-322) lands here through substitute_U!
function substitute_U!(dst::AbstractVector{<:Gaugefields_4D_B200}, src::AbstractVector{<:Gaugefields_4D_B200})
    _gfb_check(ccall((:gfb_gauge_copy, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), _b200_handle(dst).ptr, _b200_handle(src).ptr),
        _b200_handle(dst).ctx.ptr)
    return dst
end

# host <-> device in the gathered layout used by save_configuration/load_configuration (src/API.jl:516-529, 625-629)
function gather_global_array(u::Gaugefields_4D_B200)
    A = Array{ComplexF64}(undef, 3, 3, u.NX, u.NY, u.NZ, u.NT)
    _gfb_check(ccall((:gfb_gauge_download, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), u.handle.ptr, u.mu - 1, A), u.handle.ctx.ptr)
    return A
end
function scatter_global_array!(u::Gaugefields_4D_B200, A::Array{ComplexF64,6})
    size(A) == (3, 3, u.NX, u.NY, u.NZ, u.NT) || throw(DimensionMismatch("expected ComplexF64[3,3,NX,NY,NZ,NT]"))
    _gfb_check(ccall((:gfb_gauge_upload, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), u.handle.ptr, u.mu - 1, A), u.handle.ctx.ptr)
    return u
end
# ILDG binary payload (the ildg-binary-data record as it is in the file: big-endian [t][z][y][x][mu][row][col], precision 32
# or 64) <-> device; replaces the site-by-site host loops of load_gaugefield! / _save_binarydata
# (src/output/ildg_format.jl:67-83, 697-746).  Byte swap, precision conversion and transpose run on the GPU.
function scatter_ildg_payload!(U::AbstractVector{<:Gaugefields_4D_B200}, payload::Vector{UInt8}, precision::Integer)
    u = first(U)
    length(payload) == u.NX * u.NY * u.NZ * u.NT * 4 * 9 * 2 * (precision ÷ 8) ||
        throw(DimensionMismatch("ILDG payload has $(length(payload)) bytes"))
    _gfb_check(ccall((:gfb_gauge_upload_ildg, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint), u.handle.ptr, payload, precision), u.handle.ctx.ptr)
    return U
end
function gather_ildg_payload(U::AbstractVector{<:Gaugefields_4D_B200}, precision::Integer)
    u = first(U)
    payload = Vector{UInt8}(undef, u.NX * u.NY * u.NZ * u.NT * 4 * 9 * 2 * (precision ÷ 8))
    _gfb_check(ccall((:gfb_gauge_download_ildg, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint), u.handle.ptr, payload, precision), u.handle.ctx.ptr)
    return payload
end

# slow-path element access (generic code and tests index fields directly)
Base.getindex(u::Gaugefields_4D_B200, i, j, x, y, z, t) = gather_global_array(u)[i, j, x, y, z, t]

# ---- momenta: the method beside the if-chain of src/TA_Gaugefields.jl:151-195 -------------------------------
function initialize_TA_Gaugefields(U::AbstractVector{<:Gaugefields_4D_B200})
    u = first(U)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    ctx = u.handle.ctx
    _gfb_check(ccall((:gfb_mom_alloc, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
        ctx.ptr, _b200_dims(u)..., out), ctx.ptr)
    h = B200MomHandle(out[], ctx)
    finalizer(x -> ccall((:gfb_mom_free, LIBGFB200), Cint, (Ptr{Cvoid},), x.ptr), h)
    return [TA_Gaugefields_4D_B200{3,8}(h, mu, _b200_dims(u)..., 3, 8) for mu = 1:4]
end

# gaussian_momenta! (src/API.jl:341-368) dispatches on the element type; add this method
function gaussian_momenta!(P::AbstractVector{<:TA_Gaugefields_4D_B200}; sigma=1.0, seed=nothing, sweep::Integer=0, rng::SiteRNGAlgorithm=Philox4x32())
    sweep >= 0 || throw(ArgumentError("sweep must be nonnegative; got $sweep"))
    h = _b200_handle(P)
    s = seed === nothing ? rand(UInt64) : UInt64(seed)
    _gfb_check(ccall((:gfb_gaussian_momenta, LIBGFB200), Cint, (Ptr{Cvoid}, UInt64, UInt64, Cdouble, Cint), h.ptr, s, UInt64(sweep), Float64(sigma), 0), h.ctx.ptr)
    return P
end

# p * p (src/TA_Gaugefields.jl:127-137)
function Base.:*(x::AbstractVector{<:TA_Gaugefields_4D_B200}, y::AbstractVector{<:TA_Gaugefields_4D_B200})
    _b200_handle(x) === _b200_handle(y) || error("B200Backend provides p*p (kinetic energy); use gfb_mom_axpy for combinations")
    out = Ref{Cdouble}(0)
    _gfb_check(ccall((:gfb_kinetic, LIBGFB200), Cint, (Ptr{Cvoid}, Ref{Cdouble}), _b200_handle(x).ptr, out), _b200_handle(x).ctx.ptr)
    return out[]
end

# ---- observables ------------------------------------------------------------------------------------------
function calculate_Plaquette(U::AbstractVector{<:Gaugefields_4D_B200}, temp1=nothing, temp2=nothing)
    out = Ref{Cdouble}(0)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_plaquette_sum, LIBGFB200), Cint, (Ptr{Cvoid}, Ref{Cdouble}), h.ptr, out), h.ctx.ptr)
    return out[]
end
function calculate_Polyakov_loop(U::AbstractVector{<:Gaugefields_4D_B200}, temp1=nothing, temp2=nothing)
    out = zeros(Cdouble, 2)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_polyakov, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h.ptr, out), h.ctx.ptr)
    return complex(out[1], out[2])
end

# ---- molecular dynamics: fused overrides of src/molecular_dynamics.jl ------------------------------------------
# A GaugeAction whose dataset is exactly plaquette ∪ plaquette' takes the fused Wilson kernels with β = 2·coefficient
# (SURVEY.md 8b); anything else must go through the primitive table (not provided by this backend: error).
function _wilson_beta(action::GaugeAction)
    β = 0.0
    for term in action.dataset
        _is_plaquette_pair(term.closedloops) || error("B200Backend fuses only plaquette+plaquette' actions; got another loop set")
        β += 2 * real(term.β)
    end
    return β
end

function md_potential(action::GaugeAction, U::AbstractVector{<:Gaugefields_4D_B200}, workspace)
    out = Ref{Cdouble}(0)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_wilson_action, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, Ref{Cdouble}), h.ptr, _wilson_beta(action), out), h.ctx.ptr)
    return -out[] / 3
end

function md_force!(force::AbstractVector{<:TA_Gaugefields_4D_B200}, action::GaugeAction, U::AbstractVector{<:Gaugefields_4D_B200}, workspace)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_force, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble), _b200_handle(force).ptr, h.ptr, _wilson_beta(action)), h.ctx.ptr)
    return nothing
end

function update_momenta!(P::AbstractVector{<:TA_Gaugefields_4D_B200}, U::AbstractVector{<:Gaugefields_4D_B200}, step_size, driver::MDDriver)
    isfinite(step_size) || throw(ArgumentError("the momentum step size must be finite; got $step_size"))
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_update_momenta, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble),
        _b200_handle(P).ptr, h.ptr, Float64(step_size), _wilson_beta(driver.action)), h.ctx.ptr)
    return P
end

function update_gaugefields!(U::AbstractVector{<:Gaugefields_4D_B200}, P::AbstractVector{<:TA_Gaugefields_4D_B200}, step_size, driver::MDDriver)
    isfinite(step_size) || throw(ArgumentError("the gauge-field step size must be finite; got $step_size"))
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_update_links, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble), h.ptr, _b200_handle(P).ptr, Float64(step_size)), h.ctx.ptr)
    return U
end

function md_hamiltonian(U::AbstractVector{<:Gaugefields_4D_B200}, p::AbstractVector{<:TA_Gaugefields_4D_B200}, driver::MDDriver)
    out = Ref{Cdouble}(0)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_hamiltonian, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ref{Cdouble}),
        h.ptr, _b200_handle(p).ptr, _wilson_beta(driver.action), out), h.ctx.ptr)
    return out[]
end

# md_trajectory! (src/molecular_dynamics.jl:712-730) in ONE library call: `steps` fused kick+drift launches.
# GFB200_FUSED=0 replays the reference's op sequence (link, kick, link) for bit-level comparisons of the ordering.
function md_trajectory!(U::AbstractVector{<:Gaugefields_4D_B200}, p::AbstractVector{<:TA_Gaugefields_4D_B200}, driver::MDDriver; diagnostics::Bool=true)
    integ = driver.integrator isa QPQ ? 0 : driver.integrator isa PQP ? 1 :
        return invoke(md_trajectory!, Tuple{Any,Any,MDDriver}, U, p, driver; diagnostics)  # custom integrators: generic loop over md_step!
    H = zeros(Cdouble, 2)
    h = _b200_handle(U)
    fused = get(ENV, "GFB200_FUSED", "1") == "1" ? 1 : 0
    _gfb_check(ccall((:gfb_md_trajectory, LIBGFB200), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Cdouble, Cint, Cint, Ptr{Cdouble}),
        h.ptr, _b200_handle(p).ptr, _wilson_beta(driver.action), driver.steps, Float64(driver.trajectory_length), integ, fused,
        diagnostics ? pointer(H) : C_NULL), h.ctx.ptr)
    diagnostics || return nothing
    return (initial_hamiltonian=H[1], final_hamiltonian=H[2], delta_hamiltonian=H[2] - H[1])
end

# ---- gradient flow and stout ----------------------------------------------------------------------------------
function flow!(U::AbstractVector{<:Gaugefields_4D_B200}, g::Gradientflow)
    h = _b200_handle(U)
    _gfb_check(ccall((:gfb_flow, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, Cint), h.ptr, Float64(g.eps), g.Nflow), h.ctx.ptr)
    return U
end

function forward!(s::STOUT_Layer, Uout::AbstractVector{<:Gaugefields_4D_B200}, ρs::Vector{<:Number}, Uin::AbstractVector{<:Gaugefields_4D_B200})
    length(ρs) == 1 || error("B200Backend implements the plaquette-staple stout layer with one ρ")
    h = _b200_handle(Uin)
    _gfb_check(ccall((:gfb_stout_forward, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}),
        _b200_handle(Uout).ptr, h.ptr, Float64(real(ρs[1])), C_NULL), h.ctx.ptr)
    substitute_U!(s.Uin, Uin)   # the backward pass recomputes C, Q and exp(Q) from the layer input (no tape in HBM)
    s.ρs[1] = ρs[1]
    return
end

function layer_pullback!(δ_prev::AbstractVector{<:Gaugefields_4D_B200}, δ_current, layer::STOUT_Layer, Uprev, temps, tempf)
    h = _b200_handle(Uprev)
    _gfb_check(ccall((:gfb_stout_backward, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble),
        _b200_handle(δ_prev).ptr, _b200_handle(δ_current).ptr, h.ptr, Float64(real(layer.ρs[1]))), h.ctx.ptr)
    return
end


# ---- primitive table (keeps the generic, un-fused algorithms of Gaugefields.jl working on this backend) --------------
# A single link field / temporary is a `gfb_field`; `U[mu]` of a configuration is a view (gfb_field_view).  Lazy
# shift / adjoint views mirror Shifted_/Adjoint_Gaugefields_4D_MPILattice (gaugefields_4D_MPILattice.jl:647-689).
mutable struct B200FieldHandle
    ptr::Ptr{Cvoid}
    ctx::B200Context
    owner::Any
end
struct B200Lazy
    field::B200FieldHandle
    shift::NTuple{4,Cint}
    dagger::Bool
end
_lazy(f::B200FieldHandle) = B200Lazy(f, (Cint(0), Cint(0), Cint(0), Cint(0)), false)
_lazy(l::B200Lazy) = l
Base.adjoint(f::B200FieldHandle) = B200Lazy(f, (Cint(0), Cint(0), Cint(0), Cint(0)), true)
Base.adjoint(l::B200Lazy) = B200Lazy(l.field, l.shift, !l.dagger)
shift_U(f::Union{B200FieldHandle,B200Lazy}, s::NTuple{4,<:Integer}) =
    (l = _lazy(f); B200Lazy(l.field, Cint.(l.shift .+ s), l.dagger))
_shiftptr(s) = all(iszero, s) ? C_NULL : pointer(collect(s))

function b200_similar(f::B200FieldHandle, dims::NTuple{4,Int})
    out = Ref{Ptr{Cvoid}}(C_NULL)
    _gfb_check(ccall((:gfb_field_alloc, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ref{Ptr{Cvoid}}), f.ctx.ptr, dims..., out), f.ctx.ptr)
    h = B200FieldHandle(out[], f.ctx, nothing)
    finalizer(x -> ccall((:gfb_field_free, LIBGFB200), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end
function b200_link(u::Gaugefields_4D_B200)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    _gfb_check(ccall((:gfb_field_view, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}), u.handle.ptr, u.mu - 1, out), u.handle.ctx.ptr)
    h = B200FieldHandle(out[], u.handle.ctx, u.handle)
    finalizer(x -> ccall((:gfb_field_free, LIBGFB200), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end
# mul!(C, A, B, alpha, beta)  (src/AbstractGaugefields.jl:2082-2105)
function LinearAlgebra.mul!(C::B200FieldHandle, A, B, α::Number=1, β::Number=0)
    a, b = _lazy(A), _lazy(B)
    sa, sb = collect(a.shift), collect(b.shift)
    GC.@preserve sa sb _gfb_check(ccall((:gfb_mul, LIBGFB200), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cint}, Cint, Ptr{Cvoid}, Ptr{Cint}, Cint, Cdouble, Cdouble, Cdouble, Cdouble),
        C.ptr, a.field.ptr, sa, a.dagger, b.field.ptr, sb, b.dagger, real(α), imag(α), real(β), imag(β)), C.ctx.ptr)
    return C
end
add_U!(C::B200FieldHandle, α::Number, A) = (a = _lazy(A);
    _gfb_check(ccall((:gfb_axpy, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cvoid}, Cint), C.ptr, real(α), imag(α), a.field.ptr, a.dagger), C.ctx.ptr); C)
add_U!(C::B200FieldHandle, A) = add_U!(C, 1.0, A)
clear_U!(C::B200FieldHandle) = (_gfb_check(ccall((:gfb_field_clear, LIBGFB200), Cint, (Ptr{Cvoid},), C.ptr), C.ctx.ptr); C)
unit_U!(C::B200FieldHandle) = (_gfb_check(ccall((:gfb_field_unit, LIBGFB200), Cint, (Ptr{Cvoid},), C.ptr), C.ctx.ptr); C)
function substitute_U!(A::B200FieldHandle, B)
    b = _lazy(B); s = collect(b.shift)
    GC.@preserve s _gfb_check(ccall((:gfb_field_copy, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cint}, Cint), A.ptr, b.field.ptr, s, b.dagger), A.ctx.ptr)
    return A
end
function LinearAlgebra.tr(A::B200FieldHandle)
    out = zeros(Cdouble, 2)
    _gfb_check(ccall((:gfb_tr, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), A.ptr, out), A.ctx.ptr)
    return complex(out[1], out[2])
end
function LinearAlgebra.tr(A::B200FieldHandle, B::B200FieldHandle)
    out = zeros(Cdouble, 2)
    _gfb_check(ccall((:gfb_tr2, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}), A.ptr, B.ptr, out), A.ctx.ptr)
    return complex(out[1], out[2])
end
Traceless_antihermitian!(Q::B200FieldHandle, M::B200FieldHandle) =
    (_gfb_check(ccall((:gfb_ta_project, LIBGFB200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), Q.ptr, M.ptr), Q.ctx.ptr); Q)
Traceless_antihermitian_add!(P::TA_Gaugefields_4D_B200, factor, M::B200FieldHandle) =
    (_gfb_check(ccall((:gfb_ta_coeffs_add, LIBGFB200), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cvoid}), P.handle.ptr, P.mu - 1, Float64(factor), M.ptr), M.ctx.ptr); P)
exptU!(E::B200FieldHandle, t, Q::B200FieldHandle, temps=nothing) =
    (_gfb_check(ccall((:gfb_exp, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cvoid}), E.ptr, Float64(t), Q.ptr), E.ctx.ptr); E)
exptU!(E::B200FieldHandle, t, P::TA_Gaugefields_4D_B200, temps=nothing) =
    (_gfb_check(ccall((:gfb_exp_mom, LIBGFB200), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Cint), E.ptr, Float64(t), P.handle.ptr, P.mu - 1), E.ctx.ptr); E)
