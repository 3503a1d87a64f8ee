# ParametricOperatorsB200.jl -- the `ccall` shim that makes ParametricOperators.jl's apply path
# run on libparops.so (sm_100a) for CuArray operands.
#
# NOT EXECUTED IN THE BUILD ENVIRONMENT: there is no Julia toolchain in the container this
# repository was built in (see DESIGN.md).  The same C ABI is exercised end-to-end from Python
# (`parametricoperators.jl_b200/engine.py` is the line-by-line twin of this file), so what is
# untested here is only Julia syntax / package glue, not the library contract.
#
# Usage (inside a project that has ParametricOperators, CUDA and ChainRulesCore):
#     using ParametricOperators, CUDA
#     include("ParametricOperatorsB200.jl"); using .ParametricOperatorsB200
#     ParametricOperatorsB200.load("/path/to/libparops.so")
#     S = spectral_convolution(Float32, [64,64,64,32], modes, 20); θ = init(S) |> gpu
#     y = S(θ) * CuArray(x)          # lowers the tree once, then one po_plan_apply per call
#
# The package's own types, constructors, `init`, `A(θ)`, `'`, `⊗`, `∘`, `distribute` stay in
# Julia untouched; this file only adds methods for `CuArray` operands (the reference dispatches
# on the array type too: src/ParOperator.jl:134-147, src/ParCommon.jl:98-104).
module ParametricOperatorsB200

using ParametricOperators
using ParametricOperators: ParOperator, ParLinearOperator, ParAdjoint, ParParameterized, ParCompose, ParKron,
    ParIdentity, ParDFT, ParRestriction, ParMatrix, ParDiagonal, ParTensor, ParDistributed, ParBroadcasted,
    ParRepartition, ParReduce, ParException, Applicable, Parametric, DDT, RDT, Domain, Range, children, parametricity
using CUDA
using ChainRulesCore
using Libdl
using MPI                                   # ParRepartition / ParBroadcasted carry MPI communicators (src/ParRepartition.jl:4-6)

# ---- library handle ------------------------------------------------------------------------
const LIB = Ref{Ptr{Cvoid}}(C_NULL)
const CTX = Dict{Int,Ptr{Cvoid}}()          # CUDA device ordinal -> po_ctx*
sym(name::Symbol) = Libdl.dlsym(LIB[], name)

function load(path::AbstractString)
    LIB[] = Libdl.dlopen(path, Libdl.RTLD_GLOBAL)
    v = ccall(sym(:po_version), Int32, ())
    v == 100 || throw(ParException("libparops ABI version $v, expected 100"))
    return nothing
end

function check(status::Int32)
    status == 0 && return
    msg = unsafe_string(ccall(sym(:po_last_error), Cstring, ()))
    throw(ParException(msg))                       # src/ParCommon.jl:4-6
end

# Device-side waits are time-bounded; a timed-out wait poisons its outputs with NaNs and raises a sticky error word (parops.h).
# Call after a host synchronisation point (e.g. before `Array(y)` leaves the GPU) -- every later apply fails loudly anyway.
function device_status!()
    st = Ref{Int32}(0)
    check(ccall(sym(:po_ctx_device_status), Int32, (Ptr{Cvoid}, Ptr{Int32}), ctx(), st))
    st[] == 0 || throw(ParException("a device-side wait timed out (status $(st[])): results since the last check are invalid"))
    return nothing
end

function ctx()
    LIB[] == C_NULL && throw(ParException("libparops.so is not loaded (there is no CPU fallback)"))
    dev = CUDA.deviceid(CUDA.device())
    get!(CTX, dev) do
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall(sym(:po_ctx_create), Int32, (Int32, Ptr{Ptr{Cvoid}}), dev, out))
        out[]
    end
end

# ---- wire format (include/parops.h) -----------------------------------------------------------
struct PoNode
    kind::Int32; ddt::Int32; rdt::Int32; adjoint::Int32; param_slot::Int32
    child_begin::Int32; child_count::Int32; aux_begin::Int32; aux_count::Int32
    domain::Int64; range::Int64
end
const NODE_IDENTITY, NODE_DFT, NODE_RESTRICTION, NODE_MATRIX, NODE_DIAGONAL, NODE_TENSOR, NODE_KRON, NODE_COMPOSE,
      NODE_REPARTITION, NODE_BIAS, NODE_GELU = Int32.(0:10)
dtcode(::Type{Float32}) = Int32(0); dtcode(::Type{Float64}) = Int32(1)
dtcode(::Type{ComplexF32}) = Int32(2); dtcode(::Type{ComplexF64}) = Int32(3)

mutable struct Lowered
    nodes::Vector{PoNode}; child_index::Vector{Int32}; aux::Vector{Int64}
    params::Vector{Any}; param_ops::Vector{Any}; slots::IdDict{Any,Int32}
end
Lowered() = Lowered(PoNode[], Int32[], Int64[], Any[], Any[], IdDict{Any,Int32}())

function slot!(lw::Lowered, leaf, θ)
    get!(lw.slots, θ) do
        push!(lw.params, θ); push!(lw.param_ops, leaf); Int32(length(lw.params) - 1)
    end
end

function add!(lw, kind, D, R, adj, slot, dom, ran; aux = Int64[], kids = Int32[])
    ab, cb = length(lw.aux), length(lw.child_index)
    append!(lw.aux, Int64.(aux)); append!(lw.child_index, Int32.(kids))
    push!(lw.nodes, PoNode(kind, dtcode(D), dtcode(R), Int32(adj), Int32(slot), cb, length(kids), ab, length(aux), dom, ran))
    return Int32(length(lw.nodes) - 1)
end

restriction_aux(A::ParRestriction) = vcat(Int64[A.n, length(A.ranges)], reduce(vcat, [[r.start, r.stop] for r in A.ranges]; init = Int64[]))

# leaves ------------------------------------------------------------------------------------------
lower!(lw, A::ParIdentity{T}, adj = false) where {T} = add!(lw, NODE_IDENTITY, T, T, 0, -1, A.n, A.n)
lower!(lw, A::ParDistributed, adj = false) =                      # src/ParIdentity.jl:20-22: only identity applies
    A.op isa ParIdentity ? add!(lw, NODE_IDENTITY, DDT(A), RDT(A), 0, -1, Domain(A), Range(A)) :
    throw(ParException("ParDistributed apply is only defined for ParIdentity"))
lower!(lw, A::ParDFT{D,R}, adj = false) where {D,R} =
    adj ? add!(lw, NODE_DFT, R, D, 1, -1, A.m, A.n; aux = [A.n]) : add!(lw, NODE_DFT, D, R, 0, -1, A.n, A.m; aux = [A.n])
lower!(lw, A::ParRestriction{T}, adj = false) where {T} =
    adj ? add!(lw, NODE_RESTRICTION, T, T, 1, -1, Range(A), A.n; aux = restriction_aux(A)) :
          add!(lw, NODE_RESTRICTION, T, T, 0, -1, A.n, Range(A); aux = restriction_aux(A))
lower!(lw, A::ParAdjoint, adj = false) = lower!(lw, A.op, !adj)
lower!(lw, A::ParBroadcasted, adj = false) = lower!(lw, A.op, adj)          # src/ParBroadcasted.jl:33-35

function lower!(lw, A::ParParameterized, adj = false)
    leaf, θ, isadj = A.op isa ParAdjoint ? (A.op.op, A.params[A.op.op], true) : (A.op, A.params, false)   # src/ParMatrix.jl:45
    isadj = xor(isadj, adj)
    T = DDT(leaf)
    if leaf isa ParMatrix
        d, r = isadj ? (leaf.m, leaf.n) : (leaf.n, leaf.m)
        return add!(lw, NODE_MATRIX, T, T, isadj, slot!(lw, leaf, θ), d, r; aux = [leaf.m, leaf.n])
    elseif leaf isa ParDiagonal
        return add!(lw, NODE_DIAGONAL, T, T, isadj, slot!(lw, leaf, θ), leaf.n, leaf.n; aux = [leaf.n])
    elseif leaf isa ParTensor
        d, r = isadj ? (Range(leaf), Domain(leaf)) : (Domain(leaf), Range(leaf))
        return add!(lw, NODE_TENSOR, T, T, isadj, slot!(lw, leaf, θ), d, r; aux = vcat(length(leaf.weight_shape), collect(leaf.weight_shape)))
    end
    throw(ParException("cannot lower parameterised $(typeof(leaf))"))
end

# combinators -----------------------------------------------------------------------------------------
function lower!(lw, A::ParCompose, adj = false)
    ops = adj ? reverse(map(adjoint, A.ops)) : A.ops                           # src/ParCompose.jl:42
    kids = Int32[lower!(lw, c) for c in ops]
    D, R = adj ? (RDT(A), DDT(A)) : (DDT(A), RDT(A))
    d, r = adj ? (Range(A), Domain(A)) : (Domain(A), Range(A))
    add!(lw, NODE_COMPOSE, D, R, 0, -1, d, r; kids = kids)
end
function lower!(lw, A::ParKron, adj = false)
    K = adj ? adjoint(A) : A                                                   # src/ParKron.jl:82
    kids = Int32[lower!(lw, c) for c in K.ops]
    add!(lw, NODE_KRON, DDT(K), RDT(K), 0, -1, Domain(K), Range(K); aux = K.order, kids = kids)
end
function lower!(lw, A::ParRepartition{T,N}, adj = false) where {T,N}
    R = adj ? adjoint(A) : A
    dims_in, _, _ = MPI.Cart_get(R.comm_in); dims_out, _, _ = MPI.Cart_get(R.comm_out)          # src/ParRepartition.jl:24-25
    aux = vcat(Int64(N), Int64.(collect(R.global_size)), Int64.(dims_in), Int64.(dims_out), Int64(MPI.Comm_rank(R.comm_union)))
    add!(lw, NODE_REPARTITION, T, T, 0, -1, Domain(R), Range(R); aux = aux)
end

# ---- plans -----------------------------------------------------------------------------------------------
# `S(θ) * x` builds a NEW ParParameterized tree on every call (src/ParOperator.jl:206-207), so plans are cached on the lowered
# SIGNATURE (nodes, children, aux, batch, device) -- exactly what po_plan_create sees -- and the parameter pointers are passed per
# call.  Plans are destroyed by a finalizer (po_plan_destroy releases tables, peer flag slots and staging regions).
mutable struct Plan
    handle::Ptr{Cvoid}; ws::CuVector{UInt8}; ws_bytes::Csize_t; tape_bytes::Csize_t
    function Plan(handle, ws, ws_bytes, tape_bytes)
        p = new(handle, ws, ws_bytes, tape_bytes)
        finalizer(p) do q
            q.handle == C_NULL || ccall(sym(:po_plan_destroy), Int32, (Ptr{Cvoid},), q.handle)
            q.handle = C_NULL
        end
        return p
    end
end
const PlanKey = Tuple{Vector{PoNode},Vector{Int32},Vector{Int64},Int32,Int,Int}   # nodes, children, aux, root, batch, device
const PLANS = Dict{PlanKey,Plan}()

function plan_for(A::ParOperator, batch::Int)
    lw = Lowered(); root = lower!(lw, A)
    key = (lw.nodes, lw.child_index, lw.aux, root, batch, CUDA.deviceid(CUDA.device()))
    p = get!(PLANS, key) do
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall(sym(:po_plan_create), Int32,
                    (Ptr{Cvoid}, Ptr{PoNode}, Int32, Ptr{Int32}, Int32, Ptr{Int64}, Int32, Int32, Int64, UInt32, Ptr{Ptr{Cvoid}}),
                    ctx(), lw.nodes, length(lw.nodes), lw.child_index, length(lw.child_index), lw.aux, length(lw.aux), root, batch, 0, out))
        wb, tb = Ref{Csize_t}(0), Ref{Csize_t}(0)
        check(ccall(sym(:po_plan_workspace_bytes), Int32, (Ptr{Cvoid}, Ptr{Csize_t}), out[], wb))
        check(ccall(sym(:po_plan_tape_bytes), Int32, (Ptr{Cvoid}, Ptr{Csize_t}), out[], tb))
        Plan(out[], CUDA.zeros(UInt8, max(1, wb[])), wb[], tb[])
    end
    return p, lw
end

param_ptrs(lw::Lowered) = Ptr{Cvoid}[reinterpret(Ptr{Cvoid}, pointer(θ)) for θ in lw.params]
stream_ptr() = reinterpret(Ptr{Cvoid}, CUDA.stream().handle)

# ---- A * x for CuArray operands (src/ParOperator.jl:221-223) ------------------------------------------------
function apply(A::ParOperator{D,R,L,<:Applicable,T}, x::CuMatrix{D}) where {D,R,L,T}
    size(x, 1) == Domain(A) || throw(ParException("operand has $(size(x,1)) rows but Domain(A) = $(Domain(A))"))
    p, lw = plan_for(A, size(x, 2))
    y = CuArray{R}(undef, Range(A), size(x, 2))
    ptrs = param_ptrs(lw)
    GC.@preserve x y ptrs p lw begin
        check(ccall(sym(:po_plan_apply), Int32,
                    (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Ptr{Cvoid}}, Int32, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                    p.handle, pointer(x), pointer(y), ptrs, length(ptrs), pointer(p.ws), p.ws_bytes, stream_ptr()))
    end
    return y
end
apply(A::ParOperator{D,R,L,<:Applicable,T}, x::CuVector{D}) where {D,R,L,T} = vec(apply(A, reshape(x, length(x), 1)))

# HOST operands through the GPU engine: H2D + apply + D2H inside one C call (po_plan_apply_host; what bench.py's e2e times).
# Not a method of `*` for Array (that stays the reference's CPU path): call it explicitly.
function apply_host(A::ParOperator{D,R,L,<:Applicable,T}, x::Matrix{D}) where {D,R,L,T}
    size(x, 1) == Domain(A) || throw(ParException("operand has $(size(x,1)) rows but Domain(A) = $(Domain(A))"))
    p, lw = plan_for(A, size(x, 2))
    y = Matrix{R}(undef, Range(A), size(x, 2))
    ptrs = param_ptrs(lw)
    GC.@preserve x y ptrs p lw begin
        check(ccall(sym(:po_plan_apply_host), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int32, Ptr{Cvoid}),
                    p.handle, pointer(x), pointer(y), ptrs, length(ptrs), stream_ptr()))
    end
    return y
end

# the same without the wait (po_plan_apply_host_async): x and y must be page-locked (CUDA.pin) and stay alive until the caller
# has synchronised the stream -- the building block of a host pipeline (upload k+1 | apply k | download k-1 on separate streams)
function apply_host_async!(y::Matrix{R}, A::ParOperator{D,R,L,<:Applicable,T}, x::Matrix{D}) where {D,R,L,T}
    size(x, 1) == Domain(A) || throw(ParException("operand has $(size(x,1)) rows but Domain(A) = $(Domain(A))"))
    size(y) == (Range(A), size(x, 2)) || throw(ParException("result buffer has size $(size(y))"))
    p, lw = plan_for(A, size(x, 2))
    ptrs = param_ptrs(lw)
    GC.@preserve x y ptrs p lw begin
        check(ccall(sym(:po_plan_apply_host_async), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int32, Ptr{Cvoid}),
                    p.handle, pointer(x), pointer(y), ptrs, length(ptrs), stream_ptr()))
    end
    return y
end

# single-operator entries (cached one-node plans inside the library)
function dft(x::CuMatrix{T}, n::Integer; adjoint::Bool = false, real_domain::Type = T) where {T}
    # forward: x is the (n, batch) operand; adjoint: x is the spectrum and `real_domain` the DDT of the un-adjointed ParDFT
    Tin = adjoint ? real_domain : T
    m = adjoint ? n : (Tin <: Real ? n ÷ 2 + 1 : n)
    y = CuArray{adjoint ? real_domain : complex(T)}(undef, m, size(x, 2))
    GC.@preserve x y check(ccall(sym(:po_dft), Int32, (Ptr{Cvoid}, Int32, Int64, Int32, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
                                 ctx(), dtcode(Tin), n, adjoint ? 1 : 0, size(x, 2), pointer(x), pointer(y), stream_ptr()))
    y
end
function restrict(x::CuMatrix{T}, n::Integer, ranges::Vector{<:UnitRange}; adjoint::Bool = false) where {T}
    flat = Int64[v for r in ranges for v in (first(r), last(r))]
    m = adjoint ? n : sum(length, ranges)
    y = CuArray{T}(undef, m, size(x, 2))
    GC.@preserve x y flat check(ccall(sym(:po_restrict), Int32,
        (Ptr{Cvoid}, Int32, Int64, Ptr{Int64}, Int32, Int32, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
        ctx(), dtcode(T), n, flat, length(ranges), adjoint ? 1 : 0, size(x, 2), pointer(x), pointer(y), stream_ptr()))
    y
end

# Route the internal-node applies of the reference to the library when the operand lives on the GPU.
(A::ParCompose{D,R,L,<:Applicable,F,N})(x::CuMatrix{D}) where {D,R,L,F,N} = apply(A, x)
(A::ParCompose{D,R,L,<:Applicable,F,N})(x::CuVector{D}) where {D,R,L,F,N} = apply(A, x)
(A::ParKron{D,R,<:Applicable,F,N})(x::CuMatrix{D}) where {D,R,F,N} = apply(A, x)
(A::ParKron{D,R,<:Applicable,F,N})(x::CuVector{D}) where {D,R,F,N} = apply(A, x)
(A::ParRepartition{T,N})(x::CuMatrix{T}) where {T,N} = apply(A, x)

# ---- reverse mode: a ccall is opaque to Zygote, so the apply carries its own rrule (SURVEY 3.5) ---------------
# The cotangent of the OPERATOR argument must have the shape of the parameterised tree, because that is what Zygote pulls back
# through `S(θ)` (= rebuild over children, src/ParOperator.jl:206-207, down to ParParameterized(A, params[A]),
# src/ParParameterized.jl:22): a structural Tangent per node whose `ops` / `op` fields mirror the children and whose leaves carry
#   * `params = θ̄`                    for a parameterised leaf        (A.params is the array,      src/ParMatrix.jl:42)
#   * `params = Dict(leaf => θ̄)`      for a parameterised ADJOINT leaf (A.params is the whole dict, src/ParMatrix.jl:45)
# ParBroadcasted's own rrule then reads `op.op.params` and sum-reduces it to root (src/ParBroadcasted.jl:41-57).
function tangent_tree(A::ParParameterized, ḡ::IdDict)
    leaf = A.op isa ParAdjoint ? A.op.op : A.op
    g = get(ḡ, A.op isa ParAdjoint ? A.params[leaf] : A.params, nothing)
    g === nothing && return ZeroTangent()
    return A.op isa ParAdjoint ? Tangent{typeof(A)}(; params = Dict{Any,Any}(leaf => g)) : Tangent{typeof(A)}(; params = g)
end
tangent_tree(A::Union{ParCompose,ParKron}, ḡ::IdDict) = Tangent{typeof(A)}(; ops = [tangent_tree(c, ḡ) for c in A.ops])
tangent_tree(A::ParBroadcasted, ḡ::IdDict) = Tangent{typeof(A)}(; op = tangent_tree(A.op, ḡ))
tangent_tree(A::ParAdjoint, ḡ::IdDict) = Tangent{typeof(A)}(; op = tangent_tree(A.op, ḡ))
tangent_tree(::ParOperator, ::IdDict) = ZeroTangent()           # non-parametric leaves

function ChainRulesCore.rrule(::typeof(apply), A::ParOperator{D,R,L,<:Applicable,T}, x::CuMatrix{D}) where {D,R,L,T}
    p, lw = plan_for(A, size(x, 2))
    y = CuArray{R}(undef, Range(A), size(x, 2))
    tape = CUDA.zeros(UInt8, max(1, p.tape_bytes))
    ptrs = param_ptrs(lw)
    GC.@preserve x y ptrs p lw tape begin
        check(ccall(sym(:po_plan_apply_taped), Int32,
                    (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Ptr{Cvoid}}, Int32, CuPtr{Cvoid}, Csize_t, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                    p.handle, pointer(x), pointer(y), ptrs, length(ptrs), pointer(tape), p.tape_bytes, pointer(p.ws), p.ws_bytes, stream_ptr()))
    end
    function apply_pullback(ȳ)
        ȳ = CuArray{R}(unthunk(ȳ))
        x̄ = CuArray{D}(undef, Domain(A), size(x, 2))
        ḡs = [CUDA.zeros(eltype(θ), size(θ)) for θ in lw.params]
        gptrs = Ptr{Cvoid}[reinterpret(Ptr{Cvoid}, pointer(g)) for g in ḡs]
        GC.@preserve ȳ x̄ ḡs gptrs ptrs tape p lw begin
            check(ccall(sym(:po_plan_pullback), Int32,
                        (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Int32, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                        p.handle, pointer(tape), pointer(ȳ), pointer(x̄), ptrs, gptrs, length(ptrs), pointer(p.ws), p.ws_bytes, stream_ptr()))
        end
        ḡ = IdDict{Any,Any}(θ => g for (θ, g) in zip(lw.params, ḡs))     # keyed by the parameter ARRAY each slot was lowered from
        return NoTangent(), tangent_tree(A, ḡ), x̄
    end
    return y, apply_pullback
end

# ---- ParBroadcasted parameter traffic on device (src/ParBroadcasted.jl:26-57) ---------------------------------
function comm_init!(nranks::Integer, rank::Integer, id::Vector{UInt8})
    check(ccall(sym(:po_comm_init_rank), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), ctx(), nranks, rank, id))
end
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall(sym(:po_comm_unique_id), Int32, (Ptr{UInt8},), id))
    id                                   # ship to the other ranks with MPI.bcast
end
bcast!(θ::CuArray, root::Integer) =
    check(ccall(sym(:po_bcast), Int32, (Ptr{Cvoid}, CuPtr{Cvoid}, Csize_t, Int32, Ptr{Cvoid}), ctx(), pointer(θ), sizeof(θ), root, stream_ptr()))
reduce_sum!(g::CuArray{T}, root::Integer) where {T} =
    check(ccall(sym(:po_reduce_sum), Int32, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Int32, Int32, Ptr{Cvoid}),
                ctx(), pointer(g), pointer(g), length(g), dtcode(T), root, stream_ptr()))

allreduce_sum!(x::CuArray{T}) where {T} =
    check(ccall(sym(:po_allreduce_sum), Int32, (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int64, Int32, Ptr{Cvoid}),
                ctx(), pointer(x), pointer(x), length(x), dtcode(T), stream_ptr()))
# ParReduce (src/ParReduce.jl:15-40): MPI.Allreduce(SUM) on device instead of through the host; the pullback is the identity
(A::ParReduce{T})(x::CuArray{T}) where {T} = (y = copy(x); allreduce_sum!(y); y)

# Bring-your-own transport: CUDA-aware MPI.jl carries the messages, the library still packs / unpacks with its own kernels
# (po_comm_init_external).  The callbacks run on the calling thread inside po_plan_apply.
function mpi_exchange(user::Ptr{Cvoid}, ns::Int32, speers::Ptr{Int32}, sbufs::Ptr{Ptr{Cvoid}}, sbytes::Ptr{Csize_t},
                      nr::Int32, rpeers::Ptr{Int32}, rbufs::Ptr{Ptr{Cvoid}}, rbytes::Ptr{Csize_t}, stream::Ptr{Cvoid})::Int32
    try
        comm = unsafe_pointer_to_objref(user)::MPI.Comm
        CUDA.synchronize()                                            # the pack kernels have filled the send buffers
        reqs = MPI.Request[]
        for k in 1:nr                                                 # post every receive first (src/ParRepartition.jl:162-166)
            buf = unsafe_wrap(CuArray, reinterpret(CuPtr{UInt8}, unsafe_load(rbufs, k)), Int(unsafe_load(rbytes, k)))
            push!(reqs, MPI.Irecv!(buf, comm; source = Int(unsafe_load(rpeers, k)), tag = 0))
        end
        for k in 1:ns
            buf = unsafe_wrap(CuArray, reinterpret(CuPtr{UInt8}, unsafe_load(sbufs, k)), Int(unsafe_load(sbytes, k)))
            push!(reqs, MPI.Isend(buf, comm; dest = Int(unsafe_load(speers, k)), tag = 0))
        end
        MPI.Waitall(reqs)
        return Int32(0)
    catch
        return Int32(1)
    end
end
function mpi_collective(user::Ptr{Cvoid}, op::Int32, send::Ptr{Cvoid}, recv::Ptr{Cvoid}, count::Int64, dtype::Int32, root::Int32,
                        stream::Ptr{Cvoid})::Int32
    try
        comm = unsafe_pointer_to_objref(user)::MPI.Comm
        CUDA.synchronize()
        if op == 0                                                    # bcast of `count` BYTES from root
            MPI.Bcast!(unsafe_wrap(CuArray, reinterpret(CuPtr{UInt8}, recv), Int(count)), comm; root = Int(root))
            return Int32(0)
        end
        ET = (Float32, Float64, ComplexF32, ComplexF64)[dtype + 1]
        s = unsafe_wrap(CuArray, reinterpret(CuPtr{ET}, send), Int(count)); r = unsafe_wrap(CuArray, reinterpret(CuPtr{ET}, recv), Int(count))
        op == 1 ? MPI.Reduce!(s, r, MPI.SUM, comm; root = Int(root)) : MPI.Allreduce!(s, r, MPI.SUM, comm)
        return Int32(0)
    catch
        return Int32(1)
    end
end
const COMM_KEEP = Any[]                                               # communicators handed to C must stay rooted
function comm_init_mpi!(comm::MPI.Comm = MPI.COMM_WORLD)
    push!(COMM_KEEP, comm)
    ex = @cfunction(mpi_exchange, Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Ptr{Csize_t}, Int32, Ptr{Int32}, Ptr{Ptr{Cvoid}}, Ptr{Csize_t}, Ptr{Cvoid}))
    co = @cfunction(mpi_collective, Int32, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Int32, Ptr{Cvoid}))
    check(ccall(sym(:po_comm_init_external), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                ctx(), MPI.Comm_size(comm), MPI.Comm_rank(comm), ex, co, pointer_from_objref(comm)))
end
# NCCL + NVLink peer-store path (preferred on one NVSwitch box): ship the 128-byte id with MPI
function comm_init_nccl!(comm::MPI.Comm = MPI.COMM_WORLD)
    id = MPI.Comm_rank(comm) == 0 ? comm_unique_id() : zeros(UInt8, 128)
    MPI.Bcast!(id, comm; root = 0)
    comm_init!(MPI.Comm_size(comm), MPI.Comm_rank(comm), id)
end

# ---- FNO block epilogue (examples/fno.jl:34-44): gelu.(S x .+ (I_grid ⊗ W) x .+ b) in one pass ----------------
# W: the c_out x c_in parameter of the pointwise ParMatrix, x: (c_in * npts, batch), s = S x: (c_out * npts, batch)
function block_epilogue(W::CuMatrix{T}, x::CuMatrix{T}, s::Union{CuMatrix{T},Nothing} = nothing,
                        b::Union{CuVector{T},Nothing} = nothing; gelu::Bool = true) where {T<:Real}
    c_out, c_in = size(W)
    npts = (size(x, 1) ÷ c_in) * size(x, 2)
    y = CuArray{T}(undef, (size(x, 1) ÷ c_in) * c_out, size(x, 2))
    GC.@preserve W x s b y check(ccall(sym(:po_block_epilogue), Int32,
        (Ptr{Cvoid}, Int32, Int64, Int64, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int32, CuPtr{Cvoid}, Ptr{Cvoid}),
        ctx(), dtcode(T), c_in, c_out, npts, pointer(W), pointer(x), s === nothing ? CU_NULL : pointer(s),
        b === nothing ? CU_NULL : pointer(b), gelu ? 1 : 0, pointer(y), stream_ptr()))
    y
end
# fno_block(x) = block_epilogue(θ[mix_channels], x, S(θ) * x)        # replaces  gelu.(S(θ)*x .+ W(θ)*x)

# reverse pass of the block (what Zygote does to examples/fno.jl:40-44): g = gelu'(z) .* ȳ with z recomputed, s̄ = g, x̄ = (I ⊗ W') g,
# W̄ = Σ g x', b̄ = Σ g  (po_block_epilogue_pullback; nothing is taped)
function block_epilogue_pullback(W::CuMatrix{T}, x::CuMatrix{T}, ȳ::CuMatrix{T}, s::Union{CuMatrix{T},Nothing} = nothing,
                                 b::Union{CuVector{T},Nothing} = nothing; gelu::Bool = true) where {T<:Real}
    c_out, c_in = size(W)
    npts = (size(x, 1) ÷ c_in) * size(x, 2)
    s̄ = similar(ȳ); x̄ = similar(x); W̄ = similar(W)
    b̄ = b === nothing ? nothing : similar(b)
    GC.@preserve W x s b ȳ s̄ x̄ W̄ b̄ check(ccall(sym(:po_block_epilogue_pullback), Int32,
        (Ptr{Cvoid}, Int32, Int64, Int64, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Int32, CuPtr{Cvoid}, CuPtr{Cvoid},
         CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
        ctx(), dtcode(T), c_in, c_out, npts, pointer(W), pointer(x), s === nothing ? CU_NULL : pointer(s), b === nothing ? CU_NULL : pointer(b),
        gelu ? 1 : 0, pointer(ȳ), pointer(s̄), pointer(x̄), pointer(W̄), b̄ === nothing ? CU_NULL : pointer(b̄), stream_ptr()))
    return s̄, x̄, W̄, b̄
end
function ChainRulesCore.rrule(::typeof(block_epilogue), W::CuMatrix{T}, x::CuMatrix{T}, s::Union{CuMatrix{T},Nothing} = nothing,
                              b::Union{CuVector{T},Nothing} = nothing; gelu::Bool = true) where {T<:Real}
    y = block_epilogue(W, x, s, b; gelu = gelu)
    function block_epilogue_back(ȳ)
        s̄, x̄, W̄, b̄ = block_epilogue_pullback(W, x, CuArray{T}(unthunk(ȳ)), s, b; gelu = gelu)
        return NoTangent(), W̄, x̄, (s === nothing ? NoTangent() : s̄), (b === nothing ? NoTangent() : b̄)
    end
    return y, block_epilogue_back
end

# ---- the reference's JSON wire format handed to the library (po_plan_create_json: src/ASTSerialization.jl:6-58 parsed and lowered in C) --
# `json` = to_json(A) of the un-parameterised tree; returns the plan and the parameter "id" bound to each slot of `params`
function plan_from_json(json::AbstractString, batch::Integer)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall(sym(:po_plan_create_json), Int32, (Ptr{Cvoid}, Cstring, Int64, UInt32, Ptr{Ptr{Cvoid}}), ctx(), json, batch, 0, out))
    wb, tb, np = Ref{Csize_t}(0), Ref{Csize_t}(0), Ref{Int32}(0)
    check(ccall(sym(:po_plan_workspace_bytes), Int32, (Ptr{Cvoid}, Ptr{Csize_t}), out[], wb))
    check(ccall(sym(:po_plan_tape_bytes), Int32, (Ptr{Cvoid}, Ptr{Csize_t}), out[], tb))
    check(ccall(sym(:po_plan_num_params), Int32, (Ptr{Cvoid}, Ptr{Int32}), out[], np))
    ids = String[]
    for slot in 0:np[]-1
        buf = zeros(UInt8, 256)
        check(ccall(sym(:po_plan_param_id), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt8}, Csize_t), out[], slot, buf, length(buf)))
        push!(ids, unsafe_string(pointer(buf)))
    end
    return Plan(out[], CUDA.zeros(UInt8, max(1, wb[])), wb[], tb[]), ids
end

# one-node conveniences (cached plans inside the library): ParRepartition apply / adjoint without building a tree
function repartition(x::CuMatrix{T}, global_size::NTuple{N,Integer}, dims_in, dims_out, rank::Integer, n_out::Integer; adjoint::Bool = false) where {T,N}
    y = CuArray{T}(undef, n_out, size(x, 2))
    gs, di, dout = Int64.(collect(global_size)), Int32.(collect(dims_in)), Int32.(collect(dims_out))
    GC.@preserve x y gs di dout check(ccall(sym(:po_repartition), Int32,
        (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}, Ptr{Int32}, Ptr{Int32}, Int32, Int32, Int64, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Cvoid}),
        ctx(), dtcode(T), N, gs, di, dout, rank, adjoint ? 1 : 0, size(x, 2), pointer(x), pointer(y), stream_ptr()))
    y
end

export apply, apply_host, apply_host_async!, load, block_epilogue, block_epilogue_pullback, plan_from_json, repartition,
       comm_init_mpi!, comm_init_nccl!, device_status!

end # module
