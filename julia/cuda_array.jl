# cuda_array.jl — CUDAArray: a third PartitionedArrays backend next to DebugArray (src/debug_array.jl) and MPIArray
# (src/mpi_array.jl), binding libpa_b200.so (include/pa_b200.h) with ccall.
#
# STATUS: written against PartitionedArrays v0.5.7 @ 8b2b2014.  NOT executed in the build image (no Julia toolchain, no
# network there); every C entry point used below is exercised through ctypes by tests/ and by examples/hpcg_cg.c.
# julia/runtests.jl runs the reference's own test bodies against this backend when a toolchain is available.
#
# Shape of the backend (what src/debug_array.jl:34-255 and src/mpi_array.jl:105-117,221-239,525-614 define for theirs):
#   * `CUDAArray{T,N} <: AbstractArray{T,N}` — the array of parts.  ONE Julia process holds all parts (the DebugArray
#     execution model) and drives one GPU per part through pa_ctx_create_multi: context k = part k = device k.  Host-side
#     items (index partitions, COO triplets, neighbour lists — setup-time metadata) live in `items` exactly as in
#     DebugArray; `map`, `foreach`, `gather_impl!`, `scatter_impl`, `multicast_impl`, `scan_impl`, `reduction_impl`,
#     `exchange_impl!`, ... operate on them sequentially, so every generic algorithm of the package runs unchanged.
#   * device-resident payloads: `DeviceVector` (local values of one part of a PVector) and `DeviceMatrix` (local matrix of
#     one part of a PSparseMatrix) are the per-part ITEM types.  A `PVector` / `PSparseMatrix` whose partition is a
#     `CUDAArray` of those items dispatches to ONE library call per part and operation, issued from one task per part
#     (Threads.@spawn): mul!, consistent!, assemble!, dot, norm, sum, fill!, copy!, rmul!, broadcast updates, ref_cg!.
#   * `with_cuda(f) = f(distribute_with_cuda)` mirrors with_debug (src/debug_array.jl:7-9); `to_device(v)` / `to_device(A)`
#     move a host PVector / PSparseMatrix built by the package's own constructors onto the GPUs.
module PartitionedArraysB200

using PartitionedArrays, SparseArrays, SparseMatricesCSR, LinearAlgebra
import PartitionedArrays: partition, local_values, own_values, ghost_values, consistent!, assemble!, linear_indices, cartesian_indices,
    gather_impl!, scatter_impl, scatter_impl!, multicast_impl, multicast_impl!, scan_impl, reduction_impl, is_consistent,
    allocate_exchange_impl, setup_exchange_impl, exchange_impl!, scalar_indexing_action, getany, i_am_main, ExchangeGraph, @fake_async

export CUDAArray, with_cuda, distribute_with_cuda, to_device, to_host, DeviceVector, DeviceMatrix, ref_cg!, device_gauss_seidel, smooth!, device_mg_preconditioner

const LIB = get(ENV, "PA_B200_LIB", "libpa_b200.so")

pa_error() = unsafe_string(ccall((:pa_last_error, LIB), Cstring, ()))
"Turns a status code into a Julia error, like the reference's @assert / @boundscheck failures (src/p_sparse_matrix.jl:2091-2093)."
check(rc) = rc == 0 ? nothing : error("pa_b200: ", pa_error())

# ------------------------------------------------------------------------------------------ the set of device contexts
mutable struct Contexts
    h::Vector{Ptr{Cvoid}}     # pa_ctx* of part k (device k)
    function Contexts(devices::Vector{Int32}; arena_bytes::UInt64 = UInt64(8) << 30)
        h = Vector{Ptr{Cvoid}}(undef, length(devices))
        check(ccall((:pa_ctx_create_multi, LIB), Cint, (Int32, Ptr{Int32}, UInt64, Ptr{Ptr{Cvoid}}), length(devices), devices, arena_bytes, h))
        c = new(h)
        finalizer(x -> foreach(p -> ccall((:pa_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), p), x.h), c)
    end
end
const CTX = Ref{Union{Nothing,Contexts}}(nothing)
contexts() = something(CTX[])

"One task per part: every library call is collective over the parts (the device-side waits of part k need part q's call to be enqueued)."
function foreach_part(f, n::Integer)
    tasks = [Threads.@spawn f(k) for k in 1:n]
    foreach(wait, tasks)
end

# ------------------------------------------------------------------------------------------ the array of parts
"""
    CUDAArray{T,N} <: AbstractArray{T,N}

Array of parts of the CUDA backend (cf. `DebugArray`, src/debug_array.jl:34-52).  Scalar indexing is disallowed
(`scalar_indexing_action`, src/primitives.jl:4-11); the array is immutable like `DebugArray`.
"""
struct CUDAArray{T,N} <: AbstractArray{T,N}
    items::Array{T,N}
    CUDAArray{T,N}(a) where {T,N} = new{T,N}(convert(Array{T,N}, a))
    CUDAArray(a) = new{eltype(a),ndims(a)}(convert(Array{eltype(a),ndims(a)}, a))
end
Base.size(a::CUDAArray) = size(a.items)
Base.IndexStyle(::Type{<:CUDAArray}) = IndexLinear()
Base.getindex(a::CUDAArray, i::Int) = (scalar_indexing_action(a); a.items[i])
Base.setindex!(a::CUDAArray, v, i::Int) = error("CUDAArray is inmutable for performance reasons")
Base.similar(a::CUDAArray, ::Type{T}, dims::Dims) where T = error("CUDAArray is inmutable for performance reasons")
Base.copyto!(b::CUDAArray, a::CUDAArray) = error("CUDAArray is inmutable for performance reasons")
Base.map!(f, r::CUDAArray, args::CUDAArray...) = error("CUDAArray is inmutable for performance reasons")
linear_indices(a::CUDAArray) = CUDAArray(collect(LinearIndices(a)))
cartesian_indices(a::CUDAArray) = CUDAArray(collect(CartesianIndices(a)))
Base.map(f, args::CUDAArray...) = CUDAArray(map(f, map(i -> i.items, args)...))
Base.foreach(f, args::CUDAArray...) = (foreach(f, map(i -> i.items, args)...); nothing)
Base.all(a::CUDAArray) = reduce(&, a; init = true)
Base.all(p::Function, a::CUDAArray) = all(map(p, a))
Base.reduce(op, a::CUDAArray; kwargs...) = reduce(op, a.items; kwargs...)
Base.sum(a::CUDAArray) = reduce(+, a)
Base.collect(a::CUDAArray) = collect(a.items)
getany(a::CUDAArray) = first(a.items)
i_am_main(::CUDAArray) = true
function Base.show(io::IO, k::MIME"text/plain", data::CUDAArray)
    println(io, "$(length(data))-element CUDAArray (one part per GPU):")
    for (i, item) in enumerate(data.items)
        println(io, "[$i] = ", item)
    end
end
Base.show(io::IO, data::CUDAArray) = print(io, "CUDAArray(", data.items, ")")

# primitives on host items: delegated to the sequential implementations, like DebugArray (src/debug_array.jl:138-255)
gather_impl!(rcv::CUDAArray, snd::CUDAArray, destination, ::Type{T}) where T = gather_impl!(rcv.items, snd.items, destination, T)
scatter_impl(snd::CUDAArray, source) = CUDAArray(scatter_impl(snd.items, source))
scatter_impl!(rcv::CUDAArray, snd::CUDAArray, source, ::Type{T}) where T = scatter_impl!(rcv.items, snd.items, source, T)
multicast_impl(snd::CUDAArray, source) = CUDAArray(multicast_impl(snd.items, source))
multicast_impl!(rcv::CUDAArray, snd::CUDAArray, source, ::Type{T}) where T = multicast_impl!(rcv.items, snd.items, source, T)
scan_impl(op, a::CUDAArray, init, type) = CUDAArray(scan_impl(op, a.items, init, type))
reduction_impl(op, a::CUDAArray, destination; kwargs...) = CUDAArray(reduction_impl(op, a.items, destination; kwargs...))
is_consistent(graph::ExchangeGraph{<:CUDAArray}) = is_consistent(ExchangeGraph(graph.snd.items, graph.rcv.items))
allocate_exchange_impl(snd::CUDAArray, graph::ExchangeGraph{<:CUDAArray}) =
    CUDAArray(allocate_exchange_impl(snd.items, ExchangeGraph(graph.snd.items, graph.rcv.items)))
setup_exchange_impl(rcv::CUDAArray, snd::CUDAArray, graph::ExchangeGraph{<:CUDAArray}) =
    setup_exchange_impl(rcv.items, snd.items, ExchangeGraph(graph.snd.items, graph.rcv.items))
function exchange_impl!(rcv::CUDAArray, snd::CUDAArray, graph::ExchangeGraph{<:CUDAArray}, setup)
    exchange_impl!(rcv.items, snd.items, ExchangeGraph(graph.snd.items, graph.rcv.items), setup)
    @fake_async rcv
end

"`distribute_with_cuda(a)`: the backend's `distribute` (cf. distribute_with_debug, src/debug_array.jl:24-31)."
distribute_with_cuda(a) = CUDAArray(collect(a))

"""
    with_cuda(f; devices = 0:ngpus-1, arena_bytes)

`with_cuda(f) = f(distribute_with_cuda)` (cf. with_debug, src/debug_array.jl:7-9): creates one device context per part.
"""
function with_cuda(f; devices = nothing, arena_bytes::UInt64 = UInt64(8) << 30)
    devs = devices === nothing ? Int32.(0:parse(Int, get(ENV, "PA_B200_NGPUS", "1"))-1) : Int32.(collect(devices))
    CTX[] = Contexts(devs; arena_bytes)
    try
        f(distribute_with_cuda)
    finally
        finalize(CTX[]); CTX[] = nothing
    end
end

# ------------------------------------------------------------------------------------------ plan: PRange + VectorAssemblyCache
mutable struct DevicePlan
    h::Vector{Ptr{Cvoid}}   # pa_plan* per part
end
const PLANS = IdDict{Any,DevicePlan}()   # memoised per partition object, like assembly_cache (src/p_range.jl:354-376)

"Upload the exchange plan of `index_partition` (assembly_neighbors / assembly_local_indices, src/p_range.jl:417-531)."
function device_plan(index_partition::CUDAArray)
    get!(PLANS, index_partition) do
        ctx = contexts()
        np = length(index_partition)
        nsnd, nrcv = assembly_neighbors(index_partition)
        lsnd, lrcv = assembly_local_indices(index_partition, nsnd, nrcv)
        graph = ExchangeGraph(nsnd, nrcv)
        # neighbour-side local ids: one exchange of the lid lists (as the gid exchange at src/p_range.jl:517-518)
        rl_snd = exchange(lrcv, reverse(graph)) |> fetch    # for my snd entries: the neighbour's rcv lids
        rl_rcv = exchange(lsnd, graph) |> fetch             # for my rcv entries: the neighbour's snd lids
        sym = maximum(map(local_length, index_partition).items)
        h = Vector{Ptr{Cvoid}}(undef, np)
        for k in 1:np
            r = Ref{Ptr{Cvoid}}()
            check(ccall((:pa_plan_create, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), ctx.h[k], r)); h[k] = r[]
            ids = index_partition.items[k]
            o2l = collect(Int32, own_to_local(ids)); g2l = collect(Int32, ghost_to_local(ids))
            ls, lr, rs, rr = lsnd.items[k], lrcv.items[k], rl_snd.items[k], rl_rcv.items[k]
            ns, nr = collect(Int32, nsnd.items[k]), collect(Int32, nrcv.items[k])
            check(ccall((:pa_plan_set_part, LIB), Cint,
                (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32},
                 Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                h[k], 0, local_length(ids), own_length(ids), o2l, g2l, length(ns), ns, ls.ptrs, ls.data, rs.data,
                length(nr), nr, lr.ptrs, lr.data, rr.data))
            check(ccall((:pa_plan_commit, LIB), Cint, (Ptr{Cvoid}, Int64), h[k], sym))
        end
        p = DevicePlan(h)
        finalizer(x -> foreach(q -> ccall((:pa_plan_destroy, LIB), Cint, (Ptr{Cvoid},), q), x.h), p)
    end
end

# ------------------------------------------------------------------------------------------ device items
"Local values of one part of a PVector, resident in that part's HBM (the item type of a device PVector's partition)."
mutable struct DeviceVector <: AbstractVector{Float64}
    h::Ptr{Cvoid}
    part::Int
    n::Int
end
Base.size(v::DeviceVector) = (v.n,)
Base.getindex(v::DeviceVector, i::Int) = error("scalar indexing of a DeviceVector: use to_host(v)")
"Local matrix of one part of a PSparseMatrix (own rows x local columns, CSR) in that part's HBM."
mutable struct DeviceMatrix <: AbstractMatrix{Float64}
    h::Ptr{Cvoid}
    part::Int
    dims::Tuple{Int,Int}
end
Base.size(a::DeviceMatrix) = a.dims
Base.getindex(a::DeviceMatrix, i::Int, j::Int) = error("scalar indexing of a DeviceMatrix")

const DevicePVector = PVector{DeviceVector}
const DevicePSparseMatrix = PSparseMatrix{DeviceMatrix}
handles(v::PVector) = map(i -> i.h, partition(v).items)
nparts(v) = length(partition(v))

"PVector(undef, index_partition) on the device (src/p_vector.jl:334-344)."
function device_pvector(index_partition::CUDAArray)
    plan = device_plan(index_partition)
    items = map(1:length(index_partition)) do k
        r = Ref{Ptr{Cvoid}}()
        check(ccall((:pa_vec_create, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), plan.h[k], r))
        v = DeviceVector(r[], k, local_length(index_partition.items[k]))
        finalizer(x -> ccall((:pa_vec_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), v)   # symmetric heap: freed in creation order per part
    end
    PVector(CUDAArray(items), index_partition)
end
Base.similar(v::DevicePVector) = device_pvector(partition(axes(v, 1)))

"Move a host PVector (Vector{Float64} items) onto the GPUs / back."
function to_device(v::PVector)
    d = device_pvector(partition(axes(v, 1)))
    foreach_part(nparts(d)) do k
        a = partition(v).items[k]
        check(ccall((:pa_vec_upload, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64), partition(d).items[k].h, 0, a, length(a)))
    end
    d
end
function to_host(d::DevicePVector)
    items = map(i -> zeros(Float64, i.n), partition(d).items)
    foreach_part(nparts(d)) do k
        check(ccall((:pa_vec_download, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64), partition(d).items[k].h, 0, items[k], length(items[k])))
    end
    PVector(CUDAArray(items), partition(axes(d, 1)))
end

"Upload an assembled PSparseMatrix (SparseMatrixCSR{1}/CSC local matrices, split or not) — src/p_sparse_matrix.jl:971-991."
function to_device(A::PSparseMatrix)
    rows, cols = device_plan(partition(axes(A, 1))), device_plan(partition(axes(A, 2)))
    items = map(1:length(partition(A))) do k
        r = Ref{Ptr{Cvoid}}()
        check(ccall((:pa_mat_create, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), rows.h[k], cols.h[k], r))
        a = partition(A).items[k]
        upload_local!(r[], a)
        check(ccall((:pa_mat_commit, LIB), Cint, (Ptr{Cvoid},), r[]))
        m = DeviceMatrix(r[], k, size(a))
        finalizer(x -> ccall((:pa_mat_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), m)
    end
    PSparseMatrix(CUDAArray(items), partition(axes(A, 1)), partition(axes(A, 2)), A.assembled)
end
bits(::Type{T}) where T = Int32(8sizeof(T))
function upload_local!(h, a::PartitionedArrays.AbstractSplitMatrix)            # src/p_sparse_matrix.jl:588-593
    oo, oh = a.blocks.own_own, a.blocks.own_ghost
    upload_split!(h, oo, oh)
end
upload_split!(h, oo::SparseMatrixCSR{Bi,Float64,Ti}, oh::SparseMatrixCSR{Bi,Float64,Ti}) where {Bi,Ti} =
    check(ccall((:pa_mat_set_csr_split, LIB), Cint,
        (Ptr{Cvoid}, Int32, Int64, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}),
        h, 0, size(oo, 1), Bi, bits(Ti), bits(Ti), oo.rowptr, oo.colval, oo.nzval, oh.rowptr, oh.colval, oh.nzval))
upload_split!(h, oo::SparseMatrixCSC{Float64,Ti}, oh::SparseMatrixCSC{Float64,Ti}) where Ti =
    check(ccall((:pa_mat_set_csc_split, LIB), Cint,
        (Ptr{Cvoid}, Int32, Int64, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}),
        h, 0, size(oo, 1), 1, bits(Ti), bits(Ti), oo.colptr, oo.rowval, oo.nzval, oh.colptr, oh.rowval, oh.nzval))
upload_local!(h, a::SparseMatrixCSR{Bi,Float64,Ti}) where {Bi,Ti} =       # HPCG layout (HPCG/src/sparse_matrix.jl:115-121)
    check(ccall((:pa_mat_set_csr, LIB), Cint, (Ptr{Cvoid}, Int32, Int64, Int64, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}),
        h, 0, size(a, 1), size(a, 2), Bi, bits(Ti), bits(Ti), a.rowptr, a.colval, a.nzval))
upload_local!(h, a::SparseMatrixCSC{Float64,Ti}) where Ti =
    check(ccall((:pa_mat_set_csc, LIB), Cint, (Ptr{Cvoid}, Int32, Int64, Int64, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}),
        h, 0, size(a, 1), size(a, 2), 1, bits(Ti), bits(Ti), a.colptr, a.rowval, a.nzval))

# ------------------------------------------------------------------------------------------ PVector / PSparseMatrix level overloads
struct DeviceTask; n::Int; end
function Base.wait(t::DeviceTask)
    ctx = contexts()
    foreach_part(k -> check(ccall((:pa_ctx_sync, LIB), Cint, (Ptr{Cvoid},), ctx.h[k])), t.n)
end
Base.fetch(t::DeviceTask) = wait(t)

each(f, v::DevicePVector) = foreach_part(k -> check(f(partition(v).items[k].h)), nparts(v))

Base.fill!(v::DevicePVector, a) = (each(h -> ccall((:pa_vec_fill, LIB), Cint, (Ptr{Cvoid}, Float64), h, a), v); v)           # src/p_vector.jl:816-821
LinearAlgebra.rmul!(v::DevicePVector, a::Number) = (each(h -> ccall((:pa_vec_scale, LIB), Cint, (Ptr{Cvoid}, Float64), h, a), v); v)  # :1194-1199
function Base.copy!(d::DevicePVector, s::DevicePVector)                                                                     # :800-814
    foreach_part(k -> check(ccall((:pa_vec_copy, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), partition(d).items[k].h, partition(s).items[k].h)), nparts(d)); d
end
Base.copyto!(d::DevicePVector, s::DevicePVector) = copy!(d, s)
"consistent!(v) (src/p_vector.jl:747-755) / assemble!([op,] v) (:695-708): peer-load kernels; returns a waitable like the reference's task."
consistent!(v::DevicePVector) = (each(h -> ccall((:pa_vec_consistent, LIB), Cint, (Ptr{Cvoid},), h), v); DeviceTask(nparts(v)))
const PA_OP = IdDict{Any,Int32}(+ => 0, max => 1, min => 2, PartitionedArrays.insert => 6)
assemble!(v::DevicePVector) = assemble!(+, v)
assemble!(op, v::DevicePVector) = (each(h -> ccall((:pa_vec_assemble_op, LIB), Cint, (Ptr{Cvoid}, Int32), h, PA_OP[op]), v); DeviceTask(nparts(v)))

function reduce_scalar(sym::Symbol, x::DevicePVector, y = nothing)
    out = zeros(Float64, nparts(x))
    foreach_part(nparts(x)) do k
        r = Ref{Float64}()
        hx = partition(x).items[k].h
        rc = y === nothing ? ccall((sym, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), hx, r) :
                             ccall((sym, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}), hx, partition(y).items[k].h, r)
        check(rc); out[k] = r[]
    end
    out[1]   # every part holds the all-reduced value (part order sum, bitwise identical on all parts)
end
LinearAlgebra.dot(a::DevicePVector, b::DevicePVector) = reduce_scalar(:pa_vec_dot, a, b)                                    # :1189-1192
Base.sum(a::DevicePVector) = reduce_scalar(:pa_vec_sum, a)                                                                  # :1178-1187
function LinearAlgebra.norm(a::DevicePVector, p::Real = 2)                                                                  # :1201-1206
    p == 2 && return sqrt(reduce_scalar(:pa_vec_norm2, a))
    out = zeros(Float64, nparts(a))
    foreach_part(nparts(a)) do k
        r = Ref{Float64}()
        check(ccall((:pa_vec_reduce_parts, LIB), Cint, (Ptr{Cvoid}, Int32, Float64, Ptr{Float64}), partition(a).items[k].h, p == 1 ? Int32(3) : Int32(5), Float64(p), r))
        out[k] = r[]
    end
    sum(out)^(1 / p)
end
function Base.reduce(op, a::DevicePVector; kwargs...)                                                                        # :1178-1183
    out = zeros(Float64, nparts(a))
    foreach_part(nparts(a)) do k
        r = Ref{Float64}()
        check(ccall((:pa_vec_reduce_parts, LIB), Cint, (Ptr{Cvoid}, Int32, Float64, Ptr{Float64}), partition(a).items[k].h, PA_OP[op], 0.0, r))
        out[k] = r[]
    end
    reduce(op, out; kwargs...)
end
Base.maximum(a::DevicePVector) = reduce(max, a)
Base.minimum(a::DevicePVector) = reduce(min, a)

"w .= a.*x .+ b.*y — the broadcast updates of the CG loop (materialize!, src/p_vector.jl:1208-1277; own AND ghost entries when the partitions are identical)."
function waxpby!(w::DevicePVector, a::Number, x::DevicePVector, b::Number, y::DevicePVector)
    foreach_part(nparts(w)) do k
        check(ccall((:pa_vec_waxpby, LIB), Cint, (Ptr{Cvoid}, Float64, Ptr{Cvoid}, Float64, Ptr{Cvoid}),
                    partition(w).items[k].h, a, partition(x).items[k].h, b, partition(y).items[k].h))
    end
    w
end
LinearAlgebra.axpy!(a::Number, x::DevicePVector, y::DevicePVector) = waxpby!(y, a, x, 1.0, y)
LinearAlgebra.axpby!(a::Number, x::DevicePVector, b::Number, y::DevicePVector) = waxpby!(y, a, x, b, y)

"mul!(c,A,b[,α,β]) (src/p_sparse_matrix.jl:2090-2142; sub-assembled matrices end with assemble!(c)) and HPCG mul_no_lat! (HPCG/src/hpcg_utils.jl:6-17)."
function LinearAlgebra.mul!(c::DevicePVector, A::DevicePSparseMatrix, b::DevicePVector, α::Number = 1.0, β::Number = 0.0)
    foreach_part(nparts(c)) do k
        check(ccall((:pa_spmv, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, UInt32),
                    partition(A).items[k].h, partition(b).items[k].h, partition(c).items[k].h, α, β, 0))
    end
    c
end
function LinearAlgebra.mul!(c::DevicePVector, At::Transpose{T,<:DevicePSparseMatrix} where T, b::DevicePVector, α::Number = 1.0, β::Number = 0.0)   # :2144-2162
    A = At.parent
    foreach_part(nparts(c)) do k
        check(ccall((:pa_spmv_transpose, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64),
                    partition(A).items[k].h, partition(b).items[k].h, partition(c).items[k].h, α, β))
    end
    c
end
function Base.:*(A::DevicePSparseMatrix, b::DevicePVector)                                                                   # :2042-2049
    c = device_pvector(partition(axes(A, 1))); mul!(c, A, b); c
end
LinearAlgebra.fillstored!(A::DevicePSparseMatrix, a) =
    (foreach_part(k -> check(ccall((:pa_mat_fill_stored, LIB), Cint, (Ptr{Cvoid}, Float64), partition(A).items[k].h, a)), length(partition(A))); A)

struct PaCgResult; iters::Int32; converged::Int32; residual0::Float64; residual::Float64; end
"ref_cg!(x,A,b; tolerance, maxiter, Pl=Identity) (HPCG/src/ref_cg.jl:119-134) — the whole loop on the devices (3 launches per iteration)."
function ref_cg!(x::DevicePVector, A::DevicePSparseMatrix, b::DevicePVector; tolerance = 0.0, maxiter = length(b))
    np = nparts(x)
    res = [Ref{PaCgResult}() for _ in 1:np]; hist = [zeros(Float64, maxiter + 1) for _ in 1:np]
    foreach_part(np) do k
        check(ccall((:pa_cg, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, UInt32, Ptr{PaCgResult}, Ptr{Float64}),
                    partition(A).items[k].h, partition(x).items[k].h, partition(b).items[k].h, maxiter, tolerance, 0, res[k], hist[k]))
    end
    x, hist[1][1:res[1][].iters+1], res[1][].residual0, res[1][].residual, res[1][].iters
end

# ------------------------------------------------------------------ HPCG preconditioner (HPCG/src/mg_preconditioner.jl, PartitionedSolvers smoothers)
# gauss_seidel(p; iterations=1, sweep=:symmetric) state on the devices.  `order = :lexicographic` (default) runs the reference's
# sequential sweeps as a wavefront dataflow (bit-identical iterates); `:multicolor` is the fast, convergence-level-parity order.
mutable struct DeviceGaussSeidel
    h::Vector{Ptr{Cvoid}}      # one pa_gs per part
    A::DevicePSparseMatrix
end
function device_gauss_seidel(A::DevicePSparseMatrix; box_dims = nothing, kind = 27, order = :lexicographic)
    np = length(partition(A)); h = Vector{Ptr{Cvoid}}(undef, np)
    foreach_part(np) do k
        r = Ref{Ptr{Cvoid}}()
        check(ccall((:pa_gs_create, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), partition(A).items[k].h, r)); h[k] = r[]
        if box_dims !== nothing    # local box of the stencil operator, x fastest: closed-form wavefront levels / colours
            d = Int64[box_dims[k]...]
            check(ccall((:pa_gs_set_box, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Int64}), h[k], 0, kind, d))
        end
        check(ccall((:pa_gs_commit, LIB), Cint, (Ptr{Cvoid},), h[k]))
        order === :multicolor && check(ccall((:pa_gs_set_order, LIB), Cint, (Ptr{Cvoid}, Int32), h[k], 1))
    end
    g = DeviceGaussSeidel(h, A)
    finalizer(x -> foreach(q -> ccall((:pa_gs_destroy, LIB), Cint, (Ptr{Cvoid},), q), x.h), g)
    g
end
"smooth!(x, state, b; zero_guess) — one symmetric Gauss-Seidel iteration (PartitionedSolvers/src/smoothers.jl:98-125)"
function smooth!(x::DevicePVector, g::DeviceGaussSeidel, b::DevicePVector; zero_guess = false)
    foreach_part(k -> check(ccall((:pa_gs_smooth, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32),
                                   g.h[k], partition(x).items[k].h, partition(b).items[k].h, zero_guess ? 1 : 0)), length(g.h))
    x
end

# Mg_preconditioner (HPCG/src/mg_preconditioner.jl:44-63): A_vec[1] coarsest ... A_vec[l] finest, one smoother per level
mutable struct DeviceMgPreconditioner
    h::Vector{Ptr{Cvoid}}      # one pa_mg per part
    A_vec::Vector{DevicePSparseMatrix}
    gs::Vector{DeviceGaussSeidel}
end
function device_mg_preconditioner(A_vec::Vector{<:DevicePSparseMatrix}, gs::Vector{DeviceGaussSeidel}, dims)   # dims[level][part] = (nx, ny, nz)
    l = length(A_vec); np = length(partition(A_vec[1])); h = Vector{Ptr{Cvoid}}(undef, np)
    foreach_part(np) do k
        mats = Ptr{Cvoid}[partition(A_vec[i]).items[k].h for i in 1:l]; sm = Ptr{Cvoid}[gs[i].h[k] for i in 1:l]
        d = Int64[dims[i][k][q] for q in 1:3, i in 1:l][:]
        r = Ref{Ptr{Cvoid}}()
        check(ccall((:pa_mg_create, LIB), Cint, (Int32, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, Ptr{Int64}, Ptr{Ptr{Cvoid}}), l, mats, sm, d, r)); h[k] = r[]
    end
    P = DeviceMgPreconditioner(h, A_vec, gs)
    finalizer(x -> foreach(q -> ccall((:pa_mg_destroy, LIB), Cint, (Ptr{Cvoid},), q), x.h), P)
    P
end
"ldiv!(x, P, b) = fill!(x,0); pc_solve!(x,P,b,l; zero_guess=true) (mg_preconditioner.jl:202-206, 314-328)"
function LinearAlgebra.ldiv!(x::DevicePVector, P::DeviceMgPreconditioner, b::DevicePVector)
    foreach_part(k -> check(ccall((:pa_mg_apply, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), P.h[k], partition(x).items[k].h, partition(b).items[k].h)), length(P.h))
    x
end
"ref_cg!(x,A,b; Pl = P) (HPCG/src/ref_cg.jl:40-134) with the multigrid preconditioner on the devices"
function ref_cg!(x::DevicePVector, A::DevicePSparseMatrix, b::DevicePVector, P::DeviceMgPreconditioner; tolerance = 0.0, maxiter = length(b))
    np = nparts(x)
    res = [Ref{PaCgResult}() for _ in 1:np]; hist = [zeros(Float64, maxiter + 1) for _ in 1:np]
    foreach_part(np) do k
        check(ccall((:pa_cg_precond, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, UInt32, Ptr{PaCgResult}, Ptr{Float64}),
                    partition(A).items[k].h, partition(x).items[k].h, partition(b).items[k].h, P.h[k], maxiter, tolerance, 0, res[k], hist[k]))
    end
    x, hist[1][1:res[1][].iters+1], res[1][].residual0, res[1][].residual, res[1][].iters
end

end # module
