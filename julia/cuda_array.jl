# cuda_array.jl — a third PartitionedArrays backend next to DebugArray (src/debug_array.jl) and
# MPIArray (src/mpi_array.jl), binding libpa_b200.so (include/pa_b200.h) with ccall.
#
# STATUS: written against PartitionedArrays v0.5.7 @ 8b2b2014; NOT executed in the build image
# (no Julia toolchain there).  The same C entry points are exercised through ctypes by tests/.
#
# Model: one MPI rank (or one Julia process) per GPU holds one part, exactly like MPIArray
# (src/mpi_array.jl:105-117).  Index metadata stays in Julia (PRange, AssemblyCache); vector and
# matrix payloads live on the GPU behind opaque handles; mul!/consistent!/assemble!/dot/norm/
# broadcast updates and the HPCG CG loop are forwarded to the library.
module PartitionedArraysB200

using PartitionedArrays, SparseMatricesCSR, LinearAlgebra, MPI
import PartitionedArrays: partition, local_values, own_values, ghost_values, consistent!, assemble!

const LIB = get(ENV, "PA_B200_LIB", "libpa_b200.so")

pa_error() = unsafe_string(ccall((:pa_last_error, LIB), Cstring, ()))
macro pacall(ex)   # @pacall ccall(...)  -> throws like the reference's @assert/@boundscheck failures
    :(rc = $(esc(ex)); rc == 0 || error("pa_b200: ", pa_error()); nothing)
end

# ---------------------------------------------------------------- backend instance (with_cuda)
mutable struct CUDABackend
    h::Ptr{Cvoid}
    comm::MPI.Comm
    rank::Int32
    nparts::Int32
end

"with_cuda(f) = f(distribute) — mirrors with_mpi (src/mpi_array.jl:64-83)."
function with_cuda(f; comm=MPI.COMM_WORLD, arena_bytes::UInt64=UInt64(8) << 30)
    MPI.Initialized() || MPI.Init()
    rank, np = MPI.Comm_rank(comm), MPI.Comm_size(comm)
    h = Ref{Ptr{Cvoid}}()
    ids = Int32[rank + 1]
    dev = Int32(parse(Int, get(ENV, "LOCAL_RANK", string(rank))))
    @pacall ccall((:pa_ctx_create, LIB), Cint, (Int32, Int32, Ptr{Int32}, Int32, UInt64, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
                  np, 1, ids, dev, arena_bytes, C_NULL, h)
    b = CUDABackend(h[], comm, rank, np)
    if np > 1
        # peer-map every arena (CUDA IPC) and create the NCCL communicator used for scalar all-reduces
        handle = zeros(UInt8, 64)
        @pacall ccall((:pa_ctx_arena_export, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{UInt8}), b.h, 0, handle)
        all = MPI.Allgather(handle, comm)
        for q in 0:np-1
            q == rank && continue
            @pacall ccall((:pa_ctx_arena_import, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{UInt8}), b.h, q + 1, all[64q+1:64q+64])
        end
        uid = zeros(UInt8, 128)
        rank == 0 && @pacall ccall((:pa_nccl_unique_id, LIB), Cint, (Ptr{UInt8},), uid)
        MPI.Bcast!(uid, 0, comm)
        @pacall ccall((:pa_ctx_nccl_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), b.h, uid, rank, np)
    end
    try
        # index metadata keeps using the MPI backend of the reference; payloads go to the GPU
        f(a -> distribute_with_mpi(a; comm), b)
    finally
        ccall((:pa_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), b.h)
    end
end

# ---------------------------------------------------------------- plan: PRange + VectorAssemblyCache
mutable struct DevicePlan
    h::Ptr{Cvoid}
end

"Upload the exchange plan of `index_partition` (assembly_neighbors / assembly_local_indices, src/p_range.jl:417-531)."
function DevicePlan(b::CUDABackend, index_partition)
    h = Ref{Ptr{Cvoid}}()
    @pacall ccall((:pa_plan_create, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), b.h, h)
    nsnd, nrcv = assembly_neighbors(index_partition)
    lsnd, lrcv = assembly_local_indices(index_partition, nsnd, nrcv)
    # neighbour-side local ids (one exchange of the lid lists, like the gid exchange at src/p_range.jl:517-518)
    graph = ExchangeGraph(nsnd, nrcv)
    rl_snd = exchange_fetch(lrcv, reverse(graph))   # for my snd entries: the neighbour's rcv lids
    rl_rcv = exchange_fetch(lsnd, graph)            # for my rcv entries: the neighbour's snd lids
    map(index_partition, nsnd, nrcv, lsnd, lrcv, rl_snd, rl_rcv) do ids, ns, nr, ls, lr, rs, rr
        o2l = collect(Int32, own_to_local(ids)); g2l = collect(Int32, ghost_to_local(ids))
        @pacall ccall((:pa_plan_set_part, LIB), Cint,
            (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Int32}, Ptr{Int32},
             Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32},
             Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
            h[], 0, local_length(ids), own_length(ids), o2l, g2l,
            length(ns), collect(Int32, ns), ls.ptrs, ls.data, rs.data,
            length(nr), collect(Int32, nr), lr.ptrs, lr.data, rr.data)
    end
    sym = reduction(max, map(local_length, index_partition); destination=:all, init=0)
    @pacall ccall((:pa_plan_commit, LIB), Cint, (Ptr{Cvoid}, Int64), h[], PartitionedArrays.getany(sym))
    p = DevicePlan(h[])
    finalizer(x -> ccall((:pa_plan_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), p)
end

# ---------------------------------------------------------------- PVector payload on the GPU
mutable struct DeviceVector
    h::Ptr{Cvoid}
    plan::DevicePlan
    n_local::Int
end
function DeviceVector(plan::DevicePlan, n_local)
    h = Ref{Ptr{Cvoid}}()
    @pacall ccall((:pa_vec_create, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), plan.h, h)
    v = DeviceVector(h[], plan, n_local)
    finalizer(x -> ccall((:pa_vec_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), v)   # symmetric heap: free in SPMD order
end
upload!(v::DeviceVector, a::Vector{Float64}) = @pacall ccall((:pa_vec_upload, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64), v.h, 0, a, length(a))
download!(a::Vector{Float64}, v::DeviceVector) = @pacall ccall((:pa_vec_download, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int64), v.h, 0, a, length(a))

Base.fill!(v::DeviceVector, a) = (@pacall ccall((:pa_vec_fill, LIB), Cint, (Ptr{Cvoid}, Float64), v.h, a); v)                       # src/p_vector.jl:816-821
Base.copy!(d::DeviceVector, s::DeviceVector) = (@pacall ccall((:pa_vec_copy, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), d.h, s.h); d)      # :800-814
LinearAlgebra.rmul!(v::DeviceVector, a::Number) = (@pacall ccall((:pa_vec_scale, LIB), Cint, (Ptr{Cvoid}, Float64), v.h, a); v)       # :1194-1199
"y .= a.*x .+ b.*y  (broadcast materialize!, src/p_vector.jl:1208-1277)"
axpby!(a, x::DeviceVector, b, y::DeviceVector) = (@pacall ccall((:pa_vec_axpby, LIB), Cint, (Ptr{Cvoid}, Float64, Ptr{Cvoid}, Float64), y.h, a, x.h, b); y)
function LinearAlgebra.dot(x::DeviceVector, y::DeviceVector)                                                                          # :1189-1192
    r = Ref{Float64}(); @pacall ccall((:pa_vec_dot, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}), x.h, y.h, r); r[]
end
function LinearAlgebra.norm(x::DeviceVector)                                                                                           # :1201-1206
    r = Ref{Float64}(); @pacall ccall((:pa_vec_norm2, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), x.h, r); sqrt(r[])
end
struct DeviceTask; ctx::Ptr{Cvoid}; end
Base.wait(t::DeviceTask) = @pacall ccall((:pa_ctx_sync, LIB), Cint, (Ptr{Cvoid},), t.ctx)
consistent!(v::DeviceVector, b::CUDABackend) = (@pacall ccall((:pa_vec_consistent, LIB), Cint, (Ptr{Cvoid},), v.h); DeviceTask(b.h))  # :747-755
assemble!(v::DeviceVector, b::CUDABackend) = (@pacall ccall((:pa_vec_assemble, LIB), Cint, (Ptr{Cvoid},), v.h); DeviceTask(b.h))      # :695-708
# assemble!(op, v) (:699-708); insert(a,b) = b (:755)
const PA_OP = Dict{Any,Int32}(+ => 0, max => 1, min => 2, PartitionedArrays.insert => 6)
assemble!(op, v::DeviceVector, b::CUDABackend) = (@pacall ccall((:pa_vec_assemble_op, LIB), Cint, (Ptr{Cvoid}, Int32), v.h, PA_OP[op]); DeviceTask(b.h))
"reduce(op, a) (src/p_vector.jl:1178-1183): the per-part reduction runs on the device; the reduction over parts is the backend's `reduce`"
function reduce_own(op, v::DeviceVector)
    r = Ref{Float64}(); @pacall ccall((:pa_vec_reduce_parts, LIB), Cint, (Ptr{Cvoid}, Int32, Float64, Ptr{Float64}), v.h, PA_OP[op], 0.0, r); r[]
end
"norm(a,p) (:1201-1206): sum over parts of norm(own,p)^p, then ^(1/p)  (op 3 = sum|x|, op 5 = sum|x|^p)"
function norm_p_own(v::DeviceVector, p::Real)
    r = Ref{Float64}(); @pacall ccall((:pa_vec_reduce_parts, LIB), Cint, (Ptr{Cvoid}, Int32, Float64, Ptr{Float64}), v.h, p == 1 ? Int32(3) : Int32(5), Float64(p), r); r[]
end

# ---------------------------------------------------------------- exchange!(rcv, snd, graph) with device-resident buffers
# exchange_impl!(rcv, snd, graph, setup, ::Type{<:AbstractVector}) (src/primitives.jl:1020-1042; src/mpi_array.jl:525-614):
# snd/rcv are JaggedArrays of 8-byte elements; the receiver pulls its segments from the senders' HBM.
mutable struct DeviceExchange
    h::Ptr{Cvoid}
end
function DeviceExchange(b::CUDABackend, graph::ExchangeGraph, snd_ptrs::Vector{Int64}, rcv_ptrs::Vector{Int64}, rcv_src_offsets::Vector{Int64}, sym_snd_len)
    h = Ref{Ptr{Cvoid}}()
    @pacall ccall((:pa_xchg_create, LIB), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), b.h, h)
    map(graph.snd, graph.rcv) do s, r   # one part per process: a single item
        @pacall ccall((:pa_xchg_set_part, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Int32}, Ptr{Int64}, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}),
                      h[], 0, length(s), Int32.(s), snd_ptrs, length(r), Int32.(r), rcv_ptrs, rcv_src_offsets)
    end
    @pacall ccall((:pa_xchg_commit, LIB), Cint, (Ptr{Cvoid}, Int64), h[], sym_snd_len)
    x = DeviceExchange(h[])
    finalizer(y -> ccall((:pa_xchg_destroy, LIB), Cint, (Ptr{Cvoid},), y.h), x)
end
function exchange!(rcv::Vector{T}, snd::Vector{T}, x::DeviceExchange, b::CUDABackend) where T<:Union{Float64,Int64}
    @pacall ccall((:pa_xchg_upload_snd, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int64), x.h, 0, snd, length(snd))
    @pacall ccall((:pa_xchg_exchange, LIB), Cint, (Ptr{Cvoid},), x.h)
    @pacall ccall((:pa_xchg_download_rcv, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int64), x.h, 0, rcv, length(rcv))   # = fetch(t)
    rcv
end

# ---------------------------------------------------------------- PSparseMatrix payload on the GPU
mutable struct DeviceMatrix
    h::Ptr{Cvoid}
end
"Upload an assembled PSparseMatrix whose local matrices are SparseMatrixCSR{1,Float64,Ti} (split or not)."
function DeviceMatrix(A::PSparseMatrix, rows::DevicePlan, cols::DevicePlan)
    h = Ref{Ptr{Cvoid}}()
    @pacall ccall((:pa_mat_create, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), rows.h, cols.h, h)
    map(partition(A)) do a
        if a isa PartitionedArrays.AbstractSplitMatrix      # src/p_sparse_matrix.jl:588-593
            oo, oh = a.blocks.own_own, a.blocks.own_ghost
            Ti = eltype(oo.rowptr); bits = Int32(8sizeof(Ti))
            @pacall ccall((:pa_mat_set_csr_split, LIB), Cint,
                (Ptr{Cvoid}, Int32, Int64, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}),
                h[], 0, size(oo, 1), 1, bits, bits, oo.rowptr, oo.colval, oo.nzval, oh.rowptr, oh.colval, oh.nzval)
        else                                                  # HPCG layout (HPCG/src/sparse_matrix.jl:115-121)
            Ti = eltype(a.rowptr); bits = Int32(8sizeof(Ti))
            @pacall ccall((:pa_mat_set_csr, LIB), Cint,
                (Ptr{Cvoid}, Int32, Int64, Int64, Int32, Int32, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}),
                h[], 0, size(a, 1), size(a, 2), 1, bits, bits, a.rowptr, a.colval, a.nzval)
        end
    end
    @pacall ccall((:pa_mat_commit, LIB), Cint, (Ptr{Cvoid},), h[])
    m = DeviceMatrix(h[])
    finalizer(x -> ccall((:pa_mat_destroy, LIB), Cint, (Ptr{Cvoid},), x.h), m)
end

"mul!(c,A,b[,α,β]) (src/p_sparse_matrix.jl:2090-2142) and HPCG mul_no_lat! (HPCG/src/hpcg_utils.jl:6-17)"
LinearAlgebra.mul!(c::DeviceVector, A::DeviceMatrix, b::DeviceVector, α::Number=1.0, β::Number=0.0) =
    (@pacall ccall((:pa_spmv, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Float64, Float64, UInt32), A.h, b.h, c.h, α, β, 0); c)

struct PaCgResult; iters::Int32; converged::Int32; residual0::Float64; residual::Float64; end
"ref_cg!(x,A,b; tolerance, maxiter, Pl=Identity) (HPCG/src/ref_cg.jl:119-134) — the whole loop on the device"
function ref_cg!(x::DeviceVector, A::DeviceMatrix, b::DeviceVector; tolerance=0.0, maxiter=50)
    res = Ref{PaCgResult}(); hist = zeros(Float64, maxiter + 1)
    @pacall ccall((:pa_cg, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Float64, UInt32, Ptr{PaCgResult}, Ptr{Float64}),
                  A.h, x.h, b.h, maxiter, tolerance, 0, res, hist)
    x, hist, res[].residual0, res[].residual, res[].iters
end

end # module
