/* pa_b200.h — C ABI of libpa_b200.so, the B200-native engine behind the PartitionedArrays.jl
 * PSparseMatrix x PVector hot path (mul!, consistent!/assemble!, dot/norm/axpy, HPCG CG loop).
 *
 * The reference is pure Julia: the path sits behind multiple dispatch on the "array of parts"
 * backend type (DebugArray src/debug_array.jl, MPIArray src/mpi_array.jl).  A third backend
 * (CUDAArray, see INTEGRATION.md for the Julia ccall stubs) binds exactly the entry points
 * below; each one cites the reference interface it replaces (path:line in the reference tree).
 *
 * Conventions
 *  - every function returns int: 0 = PA_OK, negative = PA_E*; text via pa_last_error() (thread
 *    local).  Nothing throws or aborts across the ABI (Julia shim turns codes into error(...),
 *    mirroring the reference's @assert / @boundscheck failures, src/p_sparse_matrix.jl:2091-2093).
 *  - all index arrays arriving here are 1-based exactly as Julia stores them (local ids, part
 *    ids, JaggedArray ptrs, SparseMatrixCSR{1} rowptr/colval) unless index_base says otherwise.
 *  - host arrays are borrowed for the duration of the call only; the library owns all device
 *    memory behind the opaque handles.  Handles are not thread-safe.
 *  - a pa_ctx is the backend instance: it holds the parts that live in THIS process (one per
 *    process in distributed runs = the MPIArray model, src/mpi_array.jl:105-117; all parts in one
 *    process = the DebugArray model, src/debug_array.jl:7-9).  Every operation is collective over
 *    the local parts and must be called in the same order by every process (SPMD), like MPI.
 *  - vectors live in a symmetric, peer-mapped arena (CUDA IPC over NVLink/NVSwitch): a ghost
 *    value is read straight from the owner's HBM by the consuming kernel; there is no message.
 */
#ifndef PA_B200_H
#define PA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PA_ABI_VERSION 1

#define PA_OK 0
#define PA_EINVAL (-1)  /* bad argument / dimension mismatch (reference: @assert, @boundscheck) */
#define PA_ECUDA (-2)   /* CUDA runtime error (sticky: the context is unusable afterwards)      */
#define PA_ENOMEM (-3)  /* arena or device memory exhausted                                     */
#define PA_ENCCL (-4)   /* NCCL missing or failed                                               */
#define PA_ESTATE (-5)  /* object not committed / wrong state                                   */

typedef struct pa_ctx pa_ctx;   /* backend instance (array of parts)                               */
typedef struct pa_plan pa_plan; /* PRange partition + exchange plan (AssemblyCache/VectorAssemblyCache) */
typedef struct pa_vec pa_vec;   /* PVector{Vector{Float64}}                                        */
typedef struct pa_mat pa_mat;   /* PSparseMatrix (per-part CSR, Int32 columns, Float64 values)     */

int pa_abi_version(void);
const char *pa_last_error(void);

/* ------------------------------------------------------------------ backend / context -------
 * Replaces distribute_with_debug / with_debug (src/debug_array.jl:7-31) and distribute_with_mpi /
 * with_mpi (src/mpi_array.jl:42-83).
 *  nparts_global : number of parts of the whole job (length of `ranks`)
 *  nlocal        : parts held by this process; part_ids[k] (1-based) their global ids
 *  device        : CUDA device ordinal for all local parts (one device per process)
 *  arena_bytes   : size of the symmetric vector arena per local part (0 = 1 GiB)
 *  stream        : cudaStream_t to enqueue on (NULL = the library creates a non-blocking stream) */
int pa_ctx_create(int32_t nparts_global, int32_t nlocal, const int32_t *part_ids, int32_t device,
                  uint64_t arena_bytes, void *stream, pa_ctx **out);
int pa_ctx_destroy(pa_ctx *ctx);
/* Block until everything enqueued so far has finished (wait(t) of the reference's tasks). */
int pa_ctx_sync(pa_ctx *ctx);
int pa_ctx_stream(pa_ctx *ctx, void **stream_out);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int pa_ctx_launch_count(pa_ctx *ctx, int64_t *out);

/* One process, several GPUs (the DebugArray execution model across the box): creates ndev contexts — context k holds part
 * k+1 on devices[k] — and links them (peer access between the devices, scalar all-reduce over peer memory, no NCCL, no
 * IPC).  Drive context k from its own host thread: every operation is collective over the contexts exactly as it is over
 * processes in the one-process-per-GPU model.  out: ndev handles; destroy each with pa_ctx_destroy. */
int pa_ctx_create_multi(int32_t ndev, const int32_t *devices, uint64_t arena_bytes, pa_ctx **out);

/* Peer mapping of remote parts (distributed runs).  Export the 64-byte CUDA IPC handle of local
 * part k's arena, all-gather the handles on the host (torch.distributed / MPI), then import the
 * handle of every remote part.  Parts held by this process are linked automatically. */
int pa_ctx_arena_export(pa_ctx *ctx, int32_t k, void *handle64);
int pa_ctx_arena_import(pa_ctx *ctx, int32_t part_id, const void *handle64);

/* Scalar all-reduce across processes for dot/norm/sum — replaces reduction_impl / MPI.Allreduce!
 * (src/mpi_array.jl:478-507).  NCCL is loaded with dlopen only when this is called. */
int pa_nccl_unique_id(void *id128);
int pa_ctx_nccl_init(pa_ctx *ctx, const void *id128, int32_t rank, int32_t world);

/* ------------------------------------------------------------------ plan (PRange + cache) ----
 * Ingests, verbatim, the arrays the reference computes once per partition:
 * AbstractLocalIndices accessors (src/p_range.jl:32-160) and VectorAssemblyCache
 * (src/p_vector.jl:418-468; built by assembly_neighbors / assembly_local_indices,
 * src/p_range.jl:417-531).  All ids 1-based.
 *  own_to_local / ghost_to_local : NULL means the block layout own = 1..n_own, ghost = the rest.
 *  *_remote_lids : for every entry of snd_lids (rcv_lids) the local id of the same global id on
 *     the neighbour (the neighbour's matching rcv (snd) entry).  May be NULL when the neighbour is
 *     held by this process (derived internally); required for remote neighbours. */
int pa_plan_create(pa_ctx *ctx, pa_plan **out);
int pa_plan_set_part(pa_plan *plan, int32_t k, int64_t n_local, int64_t n_own, const int32_t *own_to_local,
                     const int32_t *ghost_to_local, int32_t n_nbr_snd, const int32_t *nbr_snd,
                     const int32_t *snd_ptrs, const int32_t *snd_lids, const int32_t *snd_remote_lids,
                     int32_t n_nbr_rcv, const int32_t *nbr_rcv, const int32_t *rcv_ptrs,
                     const int32_t *rcv_lids, const int32_t *rcv_remote_lids);
/* sym_n_local: max n_local over ALL parts of the job (0 = max over the local parts; only valid
 * when every part is local).  Fixes the symmetric arena stride of vectors on this plan. */
int pa_plan_commit(pa_plan *plan, int64_t sym_n_local);
int pa_plan_destroy(pa_plan *plan);

/* ------------------------------------------------------------------ PVector -----------------
 * PVector(undef, index_partition) / similar (src/p_vector.jl:334-344,758-772). */
int pa_vec_create(pa_plan *plan, pa_vec **out);
int pa_vec_destroy(pa_vec *v);
/* local_values(v)[k] <- host (n_local doubles, reference local order); asynchronous w.r.t. the
 * device but the host buffer may be reused on return unless it is pinned (then call pa_ctx_sync). */
int pa_vec_upload(pa_vec *v, int32_t k, const double *host, int64_t n);
/* host <- local_values(v)[k]; synchronises. */
int pa_vec_download(const pa_vec *v, int32_t k, double *host, int64_t n);
/* The same copies on a caller-chosen stream (pinned host memory; returns at once): lets a caller overlap the upload of the next
 * right-hand side and the download of the previous solution with the running solve.  The caller orders them against the
 * context's stream with events; the vector must not take part in an exchange that is in flight. */
int pa_vec_upload_async(pa_vec *v, int32_t k, const double *host_pinned, int64_t n, void *stream);
int pa_vec_download_async(const pa_vec *v, int32_t k, double *host_pinned, int64_t n, void *stream);
/* fill!(v,a) (src/p_vector.jl:816-821), copy!(dst,src) (:800-814), rmul!(v,a) (:1194-1199) —
 * all local entries (own and ghost). */
int pa_vec_fill(pa_vec *v, double a);
int pa_vec_copy(pa_vec *dst, const pa_vec *src);
int pa_vec_scale(pa_vec *v, double a);
/* Broadcast updates (src/p_vector.jl:1208-1277; own AND ghost entries are written, :1271-1276):
 *   axpby : y .= a.*x .+ b.*y        waxpby : w .= a.*x .+ b.*y   (w may alias x or y) */
int pa_vec_axpby(pa_vec *y, double a, const pa_vec *x, double b);
int pa_vec_waxpby(pa_vec *w, double a, const pa_vec *x, double b, const pa_vec *y);
/* dot (src/p_vector.jl:1189-1192), norm(v,2)^2 (:1201-1206), sum (:1178-1187): own entries only,
 * per-part partial then sum over parts in part order (+ all-reduce across processes). Synchronise. */
int pa_vec_dot(const pa_vec *x, const pa_vec *y, double *out);
int pa_vec_norm2(const pa_vec *x, double *out_sumsq);
int pa_vec_sum(const pa_vec *x, double *out);
/* consistent!(v) (src/p_vector.jl:747-755) and assemble!(+,v) (:695-708).  Asynchronous: the
 * returned state is the reference's task; pa_ctx_sync is its wait(). */
int pa_vec_consistent(pa_vec *v);
int pa_vec_assemble(pa_vec *v);
/* assemble!(op, v) (src/p_vector.jl:699-708): values[lid] = op(values[lid], ghost copy) in neighbour order (:605-609);
 * op = PA_OP_SUM (+), PA_OP_INSERT (insert(a,b) = b, :755), PA_OP_MAX, PA_OP_MIN. */
#define PA_OP_SUM 0
#define PA_OP_MAX 1
#define PA_OP_MIN 2
#define PA_OP_ABSSUM 3 /* reductions only: sum |x|   = norm(x,1)   */
#define PA_OP_ABSMAX 4 /* reductions only: max |x|                  */
#define PA_OP_ABSPOW 5 /* reductions only: sum |x|^p = norm(x,p)^p  */
#define PA_OP_INSERT 6 /* assemble only                             */
int pa_vec_assemble_op(pa_vec *v, int32_t op);
/* reduce(op, a::PVector) / maximum / minimum / norm(a,p) (src/p_vector.jl:1178-1183, :1201-1206): out[k] =
 * reduce(op, own_values(a)[k]; init = neutral_element(op)) (:1170-1175) for every local part k -- one single-pass
 * kernel per part; the (tiny) reduction over parts is the caller's second step (reduce(op,b), :1182).  p is the
 * exponent of PA_OP_ABSPOW.  Synchronises. */
int pa_vec_reduce_parts(const pa_vec *x, int32_t op, double p, double *out);

/* ------------------------------------------------------------------ exchange!(rcv, snd, graph) ----------
 * ExchangeGraph (src/primitives.jl:728-741) + the vector-payload exchange (exchange!/exchange_impl! :992-1042;
 * src/debug_array.jl:250-255, src/mpi_array.jl:525-614) for 1/2/4/8-byte elements (Float64 / Int64 by default; Int32, Float32, ...).
 * snd_ids / rcv_ids: 1-based part ids (graph.snd[p], graph.rcv[p]); snd_ptrs / rcv_ptrs: the 1-based JaggedArray ptrs
 * of the send / receive buffers (n+1 entries; what allocate_exchange_impl computes, :921-947).
 * rcv_src_offsets (nullable): for every source i the 0-based position of the segment addressed to this part inside the
 * SOURCE's send buffer; derived internally when the source is held by this process, required otherwise.
 * The send buffers live in the symmetric arena: a receiver reads its segment straight from the sender's HBM. */
typedef struct pa_xchg pa_xchg;
int pa_xchg_create(pa_ctx *ctx, pa_xchg **out);
/* bytes per payload element: 8 (default), 4 (the reference's Int32 index lists, src/p_range.jl:489-531; Float32), 2 or 1;
 * before pa_xchg_commit.  ptrs, offsets and the n of upload/download count elements. */
int pa_xchg_set_elem_size(pa_xchg *x, int32_t bytes);
int pa_xchg_set_part(pa_xchg *x, int32_t k, int32_t n_snd, const int32_t *snd_ids, const int64_t *snd_ptrs, int32_t n_rcv,
                     const int32_t *rcv_ids, const int64_t *rcv_ptrs, const int64_t *rcv_src_offsets);
/* sym_snd_len: longest send buffer over ALL parts of the job (0 = over the local parts; only when every part is local) */
int pa_xchg_commit(pa_xchg *x, int64_t sym_snd_len);
int pa_xchg_destroy(pa_xchg *x);
int pa_xchg_upload_snd(pa_xchg *x, int32_t k, const void *data, int64_t n);
int pa_xchg_exchange(pa_xchg *x);
int pa_xchg_download_rcv(pa_xchg *x, int32_t k, void *data, int64_t n);
/* v[lid] = hash(gid, seed) in [-1,1) for a box partition (own box lo..hi, 0-based, hi exclusive,
 * of a gn grid; column-major ids); ghost entries are set to 0.  Test/bench input generator. */
int pa_vec_fill_hash_box(pa_vec *v, int32_t k, const int64_t *gn, const int64_t *lo, const int64_t *hi,
                         uint64_t seed);

/* ------------------------------------------------------------------ PSparseMatrix -----------
 * PSparseMatrix (src/p_sparse_matrix.jl:971-991).  Row plan / column plan = partition(axes(A,1)),
 * partition(axes(A,2)).  A part stores either its own rows only (assembled matrices: the ghost-row blocks are empty,
 * :1704-1705) or ALL local rows of an own-first row partition (sub-assembled matrices, psparse(...; assemble=false)): mul!
 * then multiplies own and ghost rows and finishes with assemble!(c) (:2109-2142). */
int pa_mat_create(pa_plan *rows, pa_plan *cols, pa_mat **out);
int pa_mat_destroy(pa_mat *A);
/* Unsplit local CSR, the HPCG layout (HPCG/src/sparse_matrix.jl:115-121): n_own_rows x n_local_cols,
 * row i = i-th own row, column ids = local ids of the column partition.
 * index_base 0/1; ptr_bits and col_bits 32 or 64 (SparseMatrixCSR{Bi,Float64,Ti}). */
int pa_mat_set_csr(pa_mat *A, int32_t k, int64_t nrows, int64_t ncols, int32_t index_base, int32_t ptr_bits,
                   int32_t col_bits, const void *rowptr, const void *colval, const double *nzval);
/* Split format (src/p_sparse_matrix.jl:588-593): own_own (n_own_rows x n_own_cols, own ids) and
 * own_ghost (n_own_rows x n_ghost_cols, ghost ids).  Rows are merged on upload, own-block entries
 * first, so the summation order of mul! (:2099-2101) is kept. */
int pa_mat_set_csr_split(pa_mat *A, int32_t k, int64_t nrows, int32_t index_base, int32_t ptr_bits,
                         int32_t col_bits, const void *rowptr_oo, const void *colval_oo,
                         const double *nzval_oo, const void *rowptr_oh, const void *colval_oh,
                         const double *nzval_oh);
/* SparseMatrixCSC local matrices — the reference's DEFAULT storage (src/p_sparse_matrix.jl:1132-1135; spmv_csc!
 * src/sparse_utils.jl:671-690).  Converted to CSR at upload; rows list their entries by ascending column, which is
 * the order in which spmv_csc! scatters into b[row], so results stay bit-identical.  Unsplit: n_local_rows x
 * n_local_cols with own rows first and empty ghost rows.  Split: own_own / own_ghost CSC blocks. */
int pa_mat_set_csc(pa_mat *A, int32_t k, int64_t nrows, int64_t ncols, int32_t index_base, int32_t ptr_bits,
                   int32_t idx_bits, const void *colptr, const void *rowval, const double *nzval);
int pa_mat_set_csc_split(pa_mat *A, int32_t k, int64_t nrows, int32_t index_base, int32_t ptr_bits, int32_t idx_bits,
                         const void *colptr_oo, const void *rowval_oo, const double *nzval_oo,
                         const void *colptr_oh, const void *rowval_oh, const double *nzval_oh);
/* sparse_matrix(T,I,J,V,m,n; reuse=true) on the device (src/sparse_utils.jl:392-405; used by psparse, src/p_sparse_matrix.jl
 * :1196-1203): COO with 1-based OWN row ids and LOCAL column ids (ids < 1 are skipped like the reference: they become a
 * stored (1,1,0)); columns sorted within rows, duplicates added in input order.  The pattern cache (the reference's K,
 * precompute_nzindex :434-455) is kept so that pa_mat_update_coo_values = sparse_matrix!(A,V,K) / psparse! (:457-469,
 * src/p_sparse_matrix.jl:1291-1305) refreshes the values of the same pattern with one kernel. */
int pa_mat_set_coo(pa_mat *A, int32_t k, int64_t n, int32_t idx_bits, const void *I, const void *J, const double *V);
/* perm[0..n) = stable ascending order of 64-bit keys (device radix sort; host arrays in and out).  Setup-time helper for the
 * host-side assembly stages: with key = (row << 32) | col it is the entry order of sparse_matrix / compresscoo
 * (SparseArrays / SparseMatricesCSR, called by src/p_sparse_matrix.jl:1186-1222), duplicates in input order. */
int pa_sort_perm_u64(pa_ctx *ctx, const uint64_t *keys, int64_t n, int32_t *perm);
int pa_mat_update_coo_values(pa_mat *A, int32_t k, const double *V, int64_t n);
/* On-device generators of the benchmark operators for a box partition (own box lo..hi of a gn grid,
 * 0-based, hi exclusive): kind 7 = gallery laplacian_fdm (src/gallery.jl:12-86), kind 27 = HPCG
 * build_matrix (HPCG/src/sparse_matrix.jl:27-80).  ghost_gid_sorted / ghost_id_of_sorted: the ng
 * ghost global ids (0-based) sorted ascending and their 0-based ghost ids (reference order).
 * rhs (nullable): kind 27 -> b = 27 - nnz_row ; kind 7 -> A*ones.  Never materialises COO. */
int pa_mat_set_stencil(pa_mat *A, int32_t k, int32_t kind, const int64_t *gn, const int64_t *lo,
                       const int64_t *hi, int64_t ng, const int64_t *ghost_gid_sorted,
                       const int32_t *ghost_id_of_sorted, pa_vec *rhs);
int pa_mat_commit(pa_mat *A);
int pa_mat_nnz(const pa_mat *A, int32_t k, int64_t *out);
/* stored rows of part k: the own rows, or all local rows for a sub-assembled matrix */
int pa_mat_nrows(const pa_mat *A, int32_t k, int64_t *out);
/* Copy the device CSR of part k back (0-based, int64 rowptr; any pointer may be NULL). */
int pa_mat_download_csr(const pa_mat *A, int32_t k, int64_t *rowptr, int32_t *colval, double *nzval);
/* LinearAlgebra.fillstored!(A,a) (used by test/p_sparse_matrix_tests.jl:285). */
int pa_mat_fill_stored(pa_mat *A, double a);

/* flags for pa_spmv / pa_cg */
#define PA_SPMV_DEFAULT 0u
#define PA_SPMV_EXPLICIT_EXCHANGE 1u  /* consistent!(x) kernel first, then a purely local SpMV       */
#define PA_SPMV_SKIP_GHOST_REFRESH 2u /* fused path: do not also write x's local ghost slots          */
#define PA_CG_REFERENCE_OPS 4u        /* op-for-op sequence of ref_cg.jl (copy,dot,axpby,spmv,dot,...) */
#define PA_SPMV_INLINE_PEER_LOADS 8u  /* one kernel: ghost columns dereference the owner's arena inside the SpMV */
#define PA_SPMV_OVERLAP 16u           /* consistent!(x) on a side stream || own-block product, then ghost-block product */
#define PA_SPMV_FUSED_EXCHANGE 32u    /* ONE kernel: the SpMV's producer warps pull the ghost values while the first tiles stream */
#define PA_CG_TIMING 64u              /* pa_cg*: record per-operation device times (see pa_cg_timings); runs the loop eagerly */

/* mul!(y,A,x) (src/p_sparse_matrix.jl:2090-2103) when alpha=1,beta=0; mul!(y,A,x,alpha,beta)
 * (:2105-2142) otherwise; HPCG mul_no_lat! (HPCG/src/hpcg_utils.jl:6-17) is the same call.
 * Ghost values are always read straight from the owner's HBM over NVLink by a kernel of this call — no
 * message, no pack/unpack, no MPI/NCCL.  Four schedules (results are bit-identical):
 *   default                    consistent!(x) as a peer-load gather kernel, then ONE local SpMV over own|ghost
 *                              columns (the HPCG mul_no_lat! schedule; fastest measured on B200)
 *   PA_SPMV_OVERLAP            the reference mul! latency hiding: the gather runs on a side stream while the
 *                              own-block product A_oo*x_own streams from HBM; then A_oh*x_ghost is added in order
 *   PA_SPMV_FUSED_EXCHANGE     one kernel: the producer warp of every CTA pulls its share of the ghost values into x's
 *                              ghost slots (NVLink peer loads) while the first matrix tiles are in flight; only rows that
 *                              touch a ghost column wait for the gather (needs regular rows; else falls back to default)
 *   PA_SPMV_INLINE_PEER_LOADS  one kernel: ghost columns dereference the owner's arena inside the SpMV
 * x's ghost slots are consistent on return, like after the reference's mul! (except with
 * PA_SPMV_INLINE_PEER_LOADS|PA_SPMV_SKIP_GHOST_REFRESH). */
int pa_spmv(pa_mat *A, pa_vec *x, pa_vec *y, double alpha, double beta, uint32_t flags);

/* mul!(c, transpose(A), b, alpha, beta) (src/p_sparse_matrix.jl:2144-2162): b on axes(A,1) (own entries read), c on
 * axes(A,2): c_own = beta*c_own + alpha*A_oo^T b_own + the ghost contributions alpha*A_oh^T b_own shipped to their owners
 * by assemble!; c's ghost entries are zero on return.  The local transposes are built once on the device. */
int pa_spmv_transpose(pa_mat *A, pa_vec *b, pa_vec *c, double alpha, double beta);

/* Sparse x sparse products (spmm / spmtm / rap, src/p_sparse_matrix.jl:2212-2307; what AMG needs).
 * pa_mat_spmm_local: D_k = A_k * C_k for every local part on the device (expand - stable sort - in-order compression: sums
 * in ascending local k, one rounding per product and per addition).  C holds one row per LOCAL column of A: the own rows of
 * B followed by the rows of B that A's ghost columns refer to (C = consistent(B, axes(A,2)), :2243; built by the host mirror
 * with exchange!).  D: created on (axes(A,1), axes(C,2)), not committed.
 * pa_mat_transpose_local: T_k = (A_k)^T as a matrix on (axes(A,2), axes(A,1)) — with pa_mat_spmm_local(T, B, D) the local
 * product of spmtm (transpose(A)*B, :2276-2290), whose ghost rows then travel to their owners (assemble). */
int pa_mat_spmm_local(pa_mat *A, pa_mat *C, pa_mat *D);
int pa_mat_transpose_local(pa_mat *A, pa_mat *T);

typedef struct {
  int32_t iters;
  int32_t converged;
  double residual0; /* ||b - A x0||  */
  double residual;  /* ||r|| at exit */
} pa_cg_result;

/* ref_cg!(x,A,b; tolerance, maxiter, Pl=Identity) (HPCG/src/ref_cg.jl:40-134).  history (nullable)
 * receives maxiter+1 residual norms (residual0 first).  Stops at maxiter or ||r||/||r0|| <= tol. */
int pa_cg(pa_mat *A, pa_vec *x, const pa_vec *b, int32_t maxiter, double tol, uint32_t flags,
          pa_cg_result *result, double *history);

/* Per-operation device times (CUDA events) of the last solve that ran with PA_CG_TIMING, in ms — the timing_data slots of
 * HPCG/src/ref_cg.jl:46-67 that HPCG/src/report_results.jl:89-152 reports: out6 = { DDOT, WAXPBY, SPMV, preconditioner
 * (ldiv!), total of the timed operations, iterations }.  With PA_CG_REFERENCE_OPS the split is the reference's own
 * op-for-op sequence; without it the fused kernels are binned by their dominant operation. */
int pa_cg_timings(pa_ctx *ctx, double *out6);
/* pa_cg keeps its three work vectors, the residual history buffer and the captured CUDA graph of the iteration cached per
 * (A, x, b) — released when A or the partition is destroyed, or all at once here. */
int pa_ctx_release_workspaces(pa_ctx *ctx);

/* ------------------------------------------------------------------ HPCG multigrid preconditioner ----
 * (the caller on either side of the SpMV in HPCG/src/ref_cg.jl:48.)
 * pa_gs  = gauss_seidel(p; iterations=1, sweep=:symmetric) state (PartitionedSolvers/src/smoothers.jl:82-125).
 *          The sweeps reproduce the reference's sequential per-part sweeps bit for bit (wavefront schedule).
 * pa_mg  = Mg_preconditioner (HPCG/src/mg_preconditioner.jl:44-63): levels[0] coarsest ... levels[n-1] finest. */
typedef struct pa_gs pa_gs;
typedef struct pa_mg pa_mg;
int pa_gs_create(pa_mat *A, pa_gs **out);
/* geometry hint (stencil on a local box, x fastest): closed-form wavefront levels; kind 7 or 27 */
int pa_gs_set_box(pa_gs *gs, int32_t k, int32_t kind, const int64_t *dims);
int pa_gs_commit(pa_gs *gs);
int pa_gs_destroy(pa_gs *gs);
/* Sweep order.  PA_GS_LEXICOGRAPHIC (default): the reference's sequential order (1:n, then n:-1:1,
 * PartitionedSolvers/src/smoothers.jl:162-176) executed as a wavefront dataflow: iterates bit-identical to the reference.
 * PA_GS_MULTICOLOR: colour by colour (27-pt: 8 colours, 7-pt: red/black; needs pa_gs_set_box) — the standard GPU order of
 * HPCG; same fixed point and per-row arithmetic, different iterates: convergence-level parity, gated by the reference's
 * own test (HPCG/test/hpcg_benchmark_tests.jl:31-41: scaled residual < 1e-12 after 50 iterations). */
#define PA_GS_LEXICOGRAPHIC 0
#define PA_GS_MULTICOLOR 1
int pa_gs_set_order(pa_gs *gs, int32_t order);
/* smooth!(x, state, b; zero_guess) — one symmetric Gauss-Seidel iteration (smoothers.jl:98-125) */
int pa_gs_smooth(pa_gs *gs, pa_vec *x, const pa_vec *b, int32_t zero_guess);
/* dims: nlevels x nlocal x 3 local box dims; restrict!/prolongate! are the f2c injections (:81-101,:224-251) */
int pa_mg_create(int32_t nlevels, pa_mat **A, pa_gs **gs, const int64_t *dims, pa_mg **out);
int pa_mg_destroy(pa_mg *mg);
/* ldiv!(x, P, b) = fill!(x,0); pc_solve!(x,P,b,l; zero_guess=true) (:202-206, :314-328) */
int pa_mg_apply(pa_mg *mg, pa_vec *x, const pa_vec *b);
/* ref_cg!(x,A,b; Pl = mg) — mg == NULL is pa_cg */
int pa_cg_precond(pa_mat *A, pa_vec *x, const pa_vec *b, pa_mg *mg, int32_t maxiter, double tol, uint32_t flags,
                  pa_cg_result *result, double *history);

/* spmv!(b,A,x) / spmtv!(b,A,x) on ONE local matrix in the storage and element types the reference tests
 * (src/sparse_utils.jl:609-690; test/sparse_utils_tests.jl:14-45,113-118): host arrays in, b out.
 *   kind 0 = spmv_csr!(b,x,ptr,idx,val) (:649-669): spmv! of a SparseMatrixCSR{Bi}, spmtv! of a SparseMatrixCSC
 *   kind 1 = spmv_csc!(b,x,ptr,idx,val) (:671-690): spmv! of a SparseMatrixCSC,     spmtv! of a SparseMatrixCSR{Bi}
 * ncomp = length(ptr)-1; index_base 0/1; idx_bits 32/64 (ptr and idx); val_bits 32/64 (val, x, b: Float32/Float64).
 * Sequential per-entry multiply and add in the value type, contributions in the reference's order: same bits. */
int pa_local_spmv(pa_ctx *ctx, int32_t kind, int32_t index_base, int32_t idx_bits, int32_t val_bits, int64_t ncomp, int64_t nb,
                  const void *ptr, const void *idx, const void *val, const void *x, int64_t nx, void *b);

/* Pinned host memory for the end-to-end path (cudaHostAlloc / cudaFreeHost). */
int pa_host_alloc(void **ptr, size_t bytes);
int pa_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* PA_B200_H */
