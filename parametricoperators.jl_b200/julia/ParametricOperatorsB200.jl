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
    ParRepartition, ParException, Applicable, DDT, RDT, Domain, Range, children
using CUDA
using ChainRulesCore
using Libdl

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
    dims_in, _, _ = MPI.Cart_get(R.comm_in); dims_out, _, _ = MPI.Cart_get(R.comm_out)
    aux = vcat(N, collect(R.global_size), dims_in, dims_out, MPI.Comm_rank(R.comm_union))
    add!(lw, NODE_REPARTITION, T, T, 0, -1, Domain(R), Range(R); aux = aux)
end

# ---- plans -----------------------------------------------------------------------------------------------
mutable struct Plan
    handle::Ptr{Cvoid}; ws::CuVector{UInt8}; ws_bytes::Csize_t; tape_bytes::Csize_t
    lw::Lowered
end
const PLANS = IdDict{Any,Dict{Int,Plan}}()           # operator => batch => plan

function plan_for(A::ParOperator, batch::Int)
    per_op = get!(() -> Dict{Int,Plan}(), PLANS, A)
    get!(per_op, batch) do
        lw = Lowered(); root = lower!(lw, A)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall(sym(:po_plan_create), Int32,
                    (Ptr{Cvoid}, Ptr{PoNode}, Int32, Ptr{Int32}, Int32, Ptr{Int64}, Int32, Int32, Int64, UInt32, Ptr{Ptr{Cvoid}}),
                    ctx(), lw.nodes, length(lw.nodes), lw.child_index, length(lw.child_index), lw.aux, length(lw.aux), root, batch, 0, out))
        wb, tb = Ref{Csize_t}(0), Ref{Csize_t}(0)
        check(ccall(sym(:po_plan_workspace_bytes), Int32, (Ptr{Cvoid}, Ptr{Csize_t}), out[], wb))
        check(ccall(sym(:po_plan_tape_bytes), Int32, (Ptr{Cvoid}, Ptr{Csize_t}), out[], tb))
        Plan(out[], CUDA.zeros(UInt8, max(1, wb[])), wb[], tb[], lw)
    end
end

param_ptrs(lw::Lowered) = Ptr{Cvoid}[reinterpret(Ptr{Cvoid}, pointer(θ)) for θ in lw.params]
stream_ptr() = reinterpret(Ptr{Cvoid}, CUDA.stream().handle)

# ---- A * x for CuArray operands (src/ParOperator.jl:221-223) ------------------------------------------------
function apply(A::ParOperator{D,R,L,<:Applicable,T}, x::CuMatrix{D}) where {D,R,L,T}
    size(x, 1) == Domain(A) || throw(ParException("operand has $(size(x,1)) rows but Domain(A) = $(Domain(A))"))
    p = plan_for(A, size(x, 2))
    y = CuArray{R}(undef, Range(A), size(x, 2))
    ptrs = param_ptrs(p.lw)
    GC.@preserve x y ptrs p begin
        check(ccall(sym(:po_plan_apply), Int32,
                    (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Ptr{Cvoid}}, Int32, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                    p.handle, pointer(x), pointer(y), ptrs, length(ptrs), pointer(p.ws), p.ws_bytes, stream_ptr()))
    end
    return y
end
apply(A::ParOperator{D,R,L,<:Applicable,T}, x::CuVector{D}) where {D,R,L,T} = vec(apply(A, reshape(x, length(x), 1)))

# Route the internal-node applies of the reference to the library when the operand lives on the GPU.
(A::ParCompose{D,R,L,<:Applicable,F,N})(x::CuMatrix{D}) where {D,R,L,F,N} = apply(A, x)
(A::ParCompose{D,R,L,<:Applicable,F,N})(x::CuVector{D}) where {D,R,L,F,N} = apply(A, x)
(A::ParKron{D,R,<:Applicable,F,N})(x::CuMatrix{D}) where {D,R,F,N} = apply(A, x)
(A::ParKron{D,R,<:Applicable,F,N})(x::CuVector{D}) where {D,R,F,N} = apply(A, x)
(A::ParRepartition{T,N})(x::CuMatrix{T}) where {T,N} = apply(A, x)

# ---- reverse mode: a ccall is opaque to Zygote, so the apply carries its own rrule (SURVEY 3.5) ---------------
function ChainRulesCore.rrule(::typeof(apply), A::ParOperator{D,R,L,<:Applicable,T}, x::CuMatrix{D}) where {D,R,L,T}
    p = plan_for(A, size(x, 2))
    y = CuArray{R}(undef, Range(A), size(x, 2))
    tape = CUDA.zeros(UInt8, max(1, p.tape_bytes))
    ptrs = param_ptrs(p.lw)
    GC.@preserve x y ptrs p tape begin
        check(ccall(sym(:po_plan_apply_taped), Int32,
                    (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Ptr{Cvoid}}, Int32, CuPtr{Cvoid}, Csize_t, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                    p.handle, pointer(x), pointer(y), ptrs, length(ptrs), pointer(tape), p.tape_bytes, pointer(p.ws), p.ws_bytes, stream_ptr()))
    end
    function apply_pullback(ȳ)
        ȳ = CuArray{R}(unthunk(ȳ))
        x̄ = CuArray{D}(undef, Domain(A), size(x, 2))
        ḡ = [CUDA.zeros(eltype(θ), size(θ)) for θ in p.lw.params]
        gptrs = Ptr{Cvoid}[reinterpret(Ptr{Cvoid}, pointer(g)) for g in ḡ]
        GC.@preserve ȳ x̄ ḡ gptrs ptrs tape begin
            check(ccall(sym(:po_plan_pullback), Int32,
                        (Ptr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, CuPtr{Cvoid}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Int32, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
                        p.handle, pointer(tape), pointer(ȳ), pointer(x̄), ptrs, gptrs, length(ptrs), pointer(p.ws), p.ws_bytes, stream_ptr()))
        end
        # cotangent of the operator = Dict(leaf => θ̄), what Zygote's Dict getindex adjoint yields for A(θ)
        Ā = Tangent{typeof(A)}(; params = Dict(leaf => g for (leaf, g) in zip(p.lw.param_ops, ḡ)))
        return NoTangent(), Ā, x̄
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

export apply, load, block_epilogue

end # module
