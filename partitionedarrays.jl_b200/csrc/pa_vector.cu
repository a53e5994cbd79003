// PVector storage and BLAS-1: fill!/copy!/rmul!/broadcast updates, single-pass dot/norm/sum
// reductions, and the consistent!/assemble! peer-pull kernels.
// Reference: src/p_vector.jl:587-612 (assemble_impl!), :695-708 (assemble!), :747-755 (consistent!),
// :800-821 (copy!/fill!), :1178-1206 (sum/dot/norm), :1194-1199 (rmul!), :1208-1277 (broadcast).
#include <iterator>

#include "pa_device.cuh"
#include "pa_internal.h"

// ------------------------------------------------------------------ symmetric arena
// Deterministic first-fit allocator with coalescing: every process performs the same sequence of allocations and releases
// (SPMD) with the same sizes (sizes derive from the job-wide symmetric lengths), so the offsets agree on all parts — the
// NVSHMEM symmetric-heap rule — whatever mixture of vector sizes the caller creates and frees.
int pa_arena_alloc(pa_ctx *c, uint64_t bytes, uint64_t *off) {
  for (auto it = c->freeblocks.begin(); it != c->freeblocks.end(); ++it) {
    if (it->second >= bytes) {
      *off = it->first;
      const uint64_t rest = it->second - bytes;
      const uint64_t at = it->first + bytes;
      c->freeblocks.erase(it);
      if (rest) c->freeblocks[at] = rest;
      return PA_OK;
    }
  }
  PA_CHECK(c->bump + bytes <= c->arena_bytes, PA_ENOMEM,
           "vector arena exhausted (%.2f GiB in use of %.2f GiB): pass a larger arena_bytes to pa_ctx_create",
           c->bump / 1073741824.0, c->arena_bytes / 1073741824.0);
  *off = c->bump;
  c->bump += bytes;
  return PA_OK;
}

void pa_arena_free(pa_ctx *c, uint64_t off, uint64_t bytes) {
  auto it = c->freeblocks.emplace(off, bytes).first;
  auto nx = std::next(it);
  if (nx != c->freeblocks.end() && it->first + it->second == nx->first) {  // merge with the block behind
    it->second += nx->second;
    c->freeblocks.erase(nx);
  }
  if (it != c->freeblocks.begin()) {  // merge with the block in front
    auto pv = std::prev(it);
    if (pv->first + pv->second == it->first) {
      pv->second += it->second;
      c->freeblocks.erase(it);
      it = pv;
    }
  }
  if (it->first + it->second == c->bump) {  // the top of the heap: give it back
    c->bump = it->first;
    c->freeblocks.erase(it);
  }
}

extern "C" int pa_vec_create(pa_plan *plan, pa_vec **out) {
  PA_CHECK(plan && out && plan->committed, PA_ESTATE, "pa_vec_create: plan missing or not committed");
  pa_ctx *c = plan->ctx;
  pa_vec *v = new pa_vec();
  v->uid = c->next_uid++;
  v->plan = plan;
  int r = pa_arena_alloc(c, plan->vec_bytes, &v->offset);
  if (r != PA_OK) {
    delete v;
    return r;
  }
  for (int k = 0; k < c->nlocal; ++k) v->d.push_back((double *)(c->arena[k] + v->offset));
  *out = v;
  return PA_OK;
}

extern "C" int pa_vec_destroy(pa_vec *v) {
  if (!v) return PA_OK;
  pa_ctx *c = v->plan->ctx;
  // the slot may be handed out again: every earlier reader (local or remote) must be finished
  cudaSetDevice(c->device);
  pa_before_write(c);
  pa_arena_free(c, v->offset, v->plan->vec_bytes);
  delete v;
  return PA_OK;
}

PeerPtrs pa_peer_ptrs(const pa_vec *v, int k) {
  PeerPtrs pp;
  pa_ctx *c = v->plan->ctx;
  const PlanPart &p = v->plan->parts[k];
  for (size_t i = 0; i < PA_MAX_NBR; ++i)
    pp.p[i] = i < p.nbrs.size() ? (double *)(c->peer_base[p.nbrs[i]] + v->offset) : nullptr;
  return pp;
}

extern "C" int pa_vec_upload(pa_vec *v, int32_t k, const double *host, int64_t n) {
  PA_CHECK(v && host, PA_EINVAL, "pa_vec_upload: null argument");
  pa_ctx *c = v->plan->ctx;
  PA_CHECK(k >= 0 && k < c->nlocal && n == v->plan->parts[k].n_local, PA_EINVAL,
           "pa_vec_upload: length %lld != n_local %lld", (long long)n, (long long)(k >= 0 && k < c->nlocal ? v->plan->parts[k].n_local : -1));
  PA_CUDA(cudaSetDevice(c->device));
  PA_TRY(pa_before_write(c));
  PA_CUDA(cudaMemcpyAsync(v->d[k], host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return PA_OK;
}

extern "C" int pa_vec_download(const pa_vec *v, int32_t k, double *host, int64_t n) {
  PA_CHECK(v && host, PA_EINVAL, "pa_vec_download: null argument");
  pa_ctx *c = v->plan->ctx;
  PA_CHECK(k >= 0 && k < c->nlocal && n == v->plan->parts[k].n_local, PA_EINVAL, "pa_vec_download: length mismatch");
  PA_CUDA(cudaSetDevice(c->device));
  PA_CUDA(cudaMemcpyAsync(host, v->d[k], n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  return pa_check_device_error(c);
}

/* Copies on a caller-chosen stream (pinned host memory), for pipelines that overlap the transfer of the NEXT right-hand side
 * and of the PREVIOUS solution with the running solve.  No ordering is added here: the caller orders the copy against the
 * context's stream with events (the vector must not take part in an exchange that is in flight). */
extern "C" int pa_vec_upload_async(pa_vec *v, int32_t k, const double *host, int64_t n, void *stream) {
  PA_CHECK(v && host && stream, PA_EINVAL, "pa_vec_upload_async: null argument");
  pa_ctx *c = v->plan->ctx;
  PA_CHECK(k >= 0 && k < c->nlocal && n == v->plan->parts[k].n_local, PA_EINVAL, "pa_vec_upload_async: length mismatch");
  PA_CUDA(cudaSetDevice(c->device));
  PA_CUDA(cudaMemcpyAsync(v->d[k], host, n * sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return PA_OK;
}
extern "C" int pa_vec_download_async(const pa_vec *v, int32_t k, double *host, int64_t n, void *stream) {
  PA_CHECK(v && host && stream, PA_EINVAL, "pa_vec_download_async: null argument");
  pa_ctx *c = v->plan->ctx;
  PA_CHECK(k >= 0 && k < c->nlocal && n == v->plan->parts[k].n_local, PA_EINVAL, "pa_vec_download_async: length mismatch");
  PA_CUDA(cudaSetDevice(c->device));
  PA_CUDA(cudaMemcpyAsync(host, v->d[k], n * sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return PA_OK;
}

// ------------------------------------------------------------------ elementwise kernels
// All updates are written as separate IEEE multiply and add (no FMA contraction) so that they are
// bit-identical to the reference's Julia broadcasts (a.*x .+ b.*y evaluates mul, mul, add).
static inline int ew_grid(int64_t n) {
  int64_t b = (n + 2 * PA_RED_THREADS - 1) / (2 * PA_RED_THREADS);
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

__global__ void __launch_bounds__(PA_RED_THREADS) k_fill(double *__restrict__ v, int64_t n, double a) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2, st = (int64_t)gridDim.x * blockDim.x * 2;
  for (; i + 1 < n; i += st) *reinterpret_cast<double2 *>(v + i) = make_double2(a, a);
  if (i < n) v[i] = a;
}

// w = ca*x + cb*y ; coefficients either immediate or num/den read from device scalars
// num == 0 (an EXACTLY zero residual: the reference's CG stops there, ref_cg.jl:22-26) gives 0 instead of 0/0, so the
// iterations a device-resident loop still runs after exact convergence leave x, r and u unchanged
__device__ __forceinline__ double coef_value(const Coef &c) {
  if (!c.num) return c.imm;
  const double n = *c.num;
  return n == 0.0 ? 0.0 : c.sign * (n / *c.den);
}

__global__ void __launch_bounds__(PA_RED_THREADS)
    k_waxpby(double *w, Coef ca, const double *x, Coef cb, const double *y, int64_t n) {
  const double a = coef_value(ca), b = coef_value(cb);
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2, st = (int64_t)gridDim.x * blockDim.x * 2;
  for (; i + 1 < n; i += st) {
    double2 xv = *reinterpret_cast<const double2 *>(x + i), yv = *reinterpret_cast<const double2 *>(y + i);
    double2 r;
    r.x = __dadd_rn(__dmul_rn(a, xv.x), __dmul_rn(b, yv.x));
    r.y = __dadd_rn(__dmul_rn(a, xv.y), __dmul_rn(b, yv.y));
    *reinterpret_cast<double2 *>(w + i) = r;
  }
  if (i < n) w[i] = __dadd_rn(__dmul_rn(a, x[i]), __dmul_rn(b, y[i]));
}

__global__ void __launch_bounds__(PA_RED_THREADS) k_scale(double *v, double a, int64_t n) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2, st = (int64_t)gridDim.x * blockDim.x * 2;
  for (; i + 1 < n; i += st) {
    double2 xv = *reinterpret_cast<double2 *>(v + i);
    xv.x = __dmul_rn(a, xv.x);
    xv.y = __dmul_rn(a, xv.y);
    *reinterpret_cast<double2 *>(v + i) = xv;
  }
  if (i < n) v[i] = __dmul_rn(a, v[i]);
}

static int same_plan(const pa_vec *a, const pa_vec *b, const char *who) {
  PA_CHECK(a && b, PA_EINVAL, "%s: null vector", who);
  if (a->plan == b->plan) return PA_OK;
  PA_CHECK(a->plan->ctx == b->plan->ctx && a->plan->parts.size() == b->plan->parts.size(), PA_EINVAL, "%s: vectors live on different backends", who);
  for (size_t k = 0; k < a->plan->parts.size(); ++k) {
    const PlanPart &pa_ = a->plan->parts[k], &pb_ = b->plan->parts[k];
    // same own ids are required (matching_own_indices, src/p_range.jl:1813-1817); ghost layouts may differ, in which
    // case broadcast updates touch own entries only (src/p_vector.jl:1271-1276) — needs the own-first layout
    PA_CHECK(pa_.n_own == pb_.n_own && (pa_.n_local == pb_.n_local || (pa_.prefix && pb_.prefix)), PA_EINVAL,
             "%s: partitions do not match (matching_own_indices, src/p_range.jl:1813-1823)", who);
  }
  return PA_OK;
}

// number of leading local entries a broadcast update touches on part k: all local entries when the vectors share the
// partition, the own entries only otherwise (BroadcastedPVector materialize!, src/p_vector.jl:1271-1276)
// "Identical partition" = the same plan object (the reference tests `a.index_partition === b.index_partition`), or two
// plans whose layout signature (own/ghost permutation, neighbours, local and owner-side ids of every ghost: computed
// once at pa_plan_commit) is equal — e.g. the column partitions of two matrices built from the same ghost set.
static int64_t bcast_extent(const pa_vec *w, const pa_vec *x, const pa_vec *y, int k) {
  const PlanPart &pw = w->plan->parts[k], &px = x->plan->parts[k], &py = y->plan->parts[k];
  const bool same = (w->plan == x->plan || (pw.n_local == px.n_local && pw.signature == px.signature)) &&
                    (w->plan == y->plan || (pw.n_local == py.n_local && pw.signature == py.signature));
  return same ? pw.n_local : pw.n_own;
}

extern "C" int pa_vec_fill(pa_vec *v, double a) {
  PA_CHECK(v, PA_EINVAL, "pa_vec_fill: null vector");
  pa_ctx *c = v->plan->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  PA_TRY(pa_before_write(c));
  for (int k = 0; k < c->nlocal; ++k) {
    int64_t n = v->plan->parts[k].n_local;
    if (!n) continue;
    k_fill<<<ew_grid(n), PA_RED_THREADS, 0, c->stream>>>(v->d[k], n, a);
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}

extern "C" int pa_vec_copy(pa_vec *dst, const pa_vec *src) {
  PA_TRY(same_plan(dst, src, "pa_vec_copy"));
  pa_ctx *c = dst->plan->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  PA_TRY(pa_before_write(c));
  for (int k = 0; k < c->nlocal; ++k) {
    // same partition: all local values; only matching own indices: own values (copyto!, src/p_vector.jl:805-814)
    int64_t n = bcast_extent(dst, src, src, k);
    if (n && dst->d[k] != src->d[k])
      PA_CUDA(cudaMemcpyAsync(dst->d[k], src->d[k], n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  }
  return PA_OK;
}

extern "C" int pa_vec_scale(pa_vec *v, double a) {
  PA_CHECK(v, PA_EINVAL, "pa_vec_scale: null vector");
  pa_ctx *c = v->plan->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  PA_TRY(pa_before_write(c));
  for (int k = 0; k < c->nlocal; ++k) {
    int64_t n = v->plan->parts[k].n_local;
    if (!n) continue;
    k_scale<<<ew_grid(n), PA_RED_THREADS, 0, c->stream>>>(v->d[k], a, n);
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}

int pa_waxpby_dev(pa_vec *w, Coef ca, const pa_vec *x, Coef cb, const pa_vec *y) {
  pa_ctx *c = w->plan->ctx;
  PA_TRY(pa_before_write(c));
  for (int k = 0; k < c->nlocal; ++k) {
    int64_t n = bcast_extent(w, x, y, k);
    if (!n) continue;
    k_waxpby<<<ew_grid(n), PA_RED_THREADS, 0, c->stream>>>(w->d[k], ca, x->d[k], cb, y->d[k], n);
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}

extern "C" int pa_vec_waxpby(pa_vec *w, double a, const pa_vec *x, double b, const pa_vec *y) {
  PA_TRY(same_plan(w, x, "pa_vec_waxpby"));
  PA_TRY(same_plan(w, y, "pa_vec_waxpby"));
  PA_CUDA(cudaSetDevice(w->plan->ctx->device));
  return pa_waxpby_dev(w, coef_imm(a), x, coef_imm(b), y);
}

extern "C" int pa_vec_axpby(pa_vec *y, double a, const pa_vec *x, double b) { return pa_vec_waxpby(y, a, x, b, y); }

// ------------------------------------------------------------------ reductions
// One pass over the data: per-thread accumulation with 128-bit loads, warp-shuffle tree, one partial
// per CTA; the last CTA to finish (atomic ticket) folds the partials in a fixed order, so the result
// is deterministic run to run.  Grid = 148 SMs x 8 CTAs.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double block_sum(double v, double *sm) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sm[w] = v;
  __syncthreads();
  double r = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0.0;
  if (w == 0) r = warp_sum(r);
  return r;  // valid in warp 0
}

__device__ __forceinline__ void grid_finish(double s, double *blockpart, unsigned *ticket, double *out, double *sm) {
  __shared__ bool last;
  if (threadIdx.x == 0) {
    blockpart[blockIdx.x] = s;
    __threadfence();
    unsigned t = atomicInc(ticket, gridDim.x - 1);  // wraps to 0: self resetting
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double a = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) a += __ldcg(blockpart + i);
    a = block_sum(a, sm);
    if (threadIdx.x == 0) *out = a;
  }
}

// MODE 0: dot(x,y)  1: sum(x.^2)  2: sum(x)
template <int MODE>
__global__ void __launch_bounds__(PA_RED_THREADS)
    k_reduce(const double *__restrict__ x, const double *__restrict__ y, int64_t n, const int32_t *__restrict__ idx,
             double *blockpart, unsigned *ticket, double *out) {
  __shared__ double sm[PA_RED_THREADS / 32];
  double acc = 0.0;
  if (idx == nullptr) {
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2, st = (int64_t)gridDim.x * blockDim.x * 2;
    double a0 = 0.0, a1 = 0.0;
    for (; i + 1 < n; i += st) {
      double2 xv = *reinterpret_cast<const double2 *>(x + i);
      if (MODE == 0) {
        double2 yv = *reinterpret_cast<const double2 *>(y + i);
        a0 = fma(xv.x, yv.x, a0);
        a1 = fma(xv.y, yv.y, a1);
      } else if (MODE == 1) {
        a0 = fma(xv.x, xv.x, a0);
        a1 = fma(xv.y, xv.y, a1);
      } else {
        a0 += xv.x;
        a1 += xv.y;
      }
    }
    if (i < n) a0 += MODE == 0 ? x[i] * y[i] : (MODE == 1 ? x[i] * x[i] : x[i]);
    acc = a0 + a1;
  } else {  // own entries scattered in the local numbering (permuted layouts)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      int32_t l = idx[i];
      acc += MODE == 0 ? x[l] * y[l] : (MODE == 1 ? x[l] * x[l] : x[l]);
    }
  }
  double s = block_sum(acc, sm);
  grid_finish(s, blockpart, ticket, out, sm);
}

static int launch_reduce(pa_ctx *c, int k, const PlanPart &pp, int mode, const double *x, const double *y, double *d_out) {
  double *out = c->nlocal == 1 ? d_out : c->d_partial + k;
  int64_t n = pp.n_own;
  int64_t per = 2 * PA_RED_THREADS * 4;
  int grid = (int)((n + per - 1) / per);
  grid = grid < 1 ? 1 : (grid > PA_RED_BLOCKS ? PA_RED_BLOCKS : grid);
  const int32_t *idx = pp.prefix ? nullptr : pp.d_own_to_local;
  double *bp = c->d_blockpart + (size_t)k * PA_RED_BLOCKS;
  if (mode == 0)
    k_reduce<0><<<grid, PA_RED_THREADS, 0, c->stream>>>(x, y, n, idx, bp, c->d_ticket + k, out);
  else if (mode == 1)
    k_reduce<1><<<grid, PA_RED_THREADS, 0, c->stream>>>(x, y, n, idx, bp, c->d_ticket + k, out);
  else
    k_reduce<2><<<grid, PA_RED_THREADS, 0, c->stream>>>(x, y, n, idx, bp, c->d_ticket + k, out);
  c->launches++;
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}

// device-resident reduction into d_scal[slot] (no host sync): used by the CG loop
int pa_reduce_dev_to(const pa_vec *x, const pa_vec *y, int mode, double *d_out) {
  pa_ctx *c = x->plan->ctx;
  for (int k = 0; k < c->nlocal; ++k) PA_TRY(launch_reduce(c, k, x->plan->parts[k], mode, x->d[k], y ? y->d[k] : nullptr, d_out));
  return pa_reduce_finish(c, d_out);
}

static int reduce_host(const pa_vec *x, const pa_vec *y, int mode, double *out, const char *who) {
  PA_CHECK(x && out, PA_EINVAL, "%s: null argument", who);
  if (y) PA_TRY(same_plan(x, y, who));
  pa_ctx *c = x->plan->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  PA_TRY(pa_reduce_dev_to(x, y, mode, c->d_scal + S_TMP));
  PA_TRY(pa_read_scalars(c, S_TMP, 1, out));
  return pa_check_device_error(c);
}

extern "C" int pa_vec_dot(const pa_vec *x, const pa_vec *y, double *out) {
  PA_CHECK(y, PA_EINVAL, "pa_vec_dot: null vector");
  return reduce_host(x, y, 0, out, "pa_vec_dot");
}
extern "C" int pa_vec_norm2(const pa_vec *x, double *out) { return reduce_host(x, nullptr, 1, out, "pa_vec_norm2"); }
extern "C" int pa_vec_sum(const pa_vec *x, double *out) { return reduce_host(x, nullptr, 2, out, "pa_vec_sum"); }

// ------------------------------------------------------------------ consistent! / assemble!
// consistent!: every ghost slot pulls the owner's value straight from the owner's HBM (NVLink peer
// load, L2 bypassed on the reading side: ld.global.cg avoids stale L1 lines).
__global__ void k_consistent(double *v, const int32_t *__restrict__ lid, const int32_t *__restrict__ slot,
                             const int32_t *__restrict__ rlid, int64_t n, PeerPtrs peers) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
    v[lid[j]] = __ldcg(peers.p[slot[j]] + rlid[j]);
}

// assemble!: each owned destination combines the neighbours' ghost copies in neighbour order
// (values[lid] = f(values[lid], buf[p]), src/p_vector.jl:605-609), one thread per destination.
// OP 0: +   1: insert(a,b) = b (src/p_vector.jl:755)   2: max   3: min
template <int OP>
__global__ void k_assemble(double *v, const int32_t *__restrict__ dst, const int32_t *__restrict__ ptr,
                           const int32_t *__restrict__ slot, const int32_t *__restrict__ rlid, int64_t ndst, PeerPtrs peers) {
  for (int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; d < ndst; d += (int64_t)gridDim.x * blockDim.x) {
    double acc = v[dst[d]];
    for (int t = ptr[d]; t < ptr[d + 1]; ++t) {
      const double b = __ldcg(peers.p[slot[t]] + rlid[t]);
      acc = OP == 0 ? __dadd_rn(acc, b) : (OP == 1 ? b : (OP == 2 ? (b > acc ? b : acc) : (b < acc ? b : acc)));
    }
    v[dst[d]] = acc;
  }
}

__global__ void k_zero_idx(double *v, const int32_t *__restrict__ idx, int64_t base, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[idx ? idx[i] : base + i] = 0.0;
}

static int small_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

// consistent! with its signalling folded in (one local part): signal "my values are final" to the neighbours, wait for
// theirs, gather, and — by the last CTA to finish — tell the neighbours "done reading" and advance the epoch.
// One launch instead of k_signal_wait + k_consistent + k_signal(done).
__global__ void __launch_bounds__(256) k_consistent_sync(double *v, const int32_t *__restrict__ lid, const int32_t *__restrict__ slot,
                                                          const int32_t *__restrict__ rlid, int64_t n, PeerPtrs peers,
                                                          unsigned long long *epoch, FlagPtrs arrive_dst, FlagPtrs arrive_src,
                                                          FlagPtrs done_dst, int nnbr, unsigned *ticket, int *err) {
  __shared__ bool last;
  const unsigned long long e = *epoch + 1ull;  // nobody writes *epoch before every CTA has passed its ticket
  if ((int)threadIdx.x < nnbr) {
    if (blockIdx.x == 0) {
      // all earlier kernels of this stream are complete (stream order): publish their writes system wide, then the epoch
      __threadfence_system();
      pa_st_release_sys(arrive_dst.p[threadIdx.x], e);
    }
    pa_spin_until(arrive_src.p[threadIdx.x], e, err);
  }
  __syncthreads();
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
    v[lid[j]] = __ldcg(peers.p[slot[j]] + rlid[j]);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;  // wraps to 0: self resetting
  }
  __syncthreads();
  if (last) {
    if ((int)threadIdx.x < nnbr) pa_st_release_sys(done_dst.p[threadIdx.x], e);
    if (threadIdx.x == 0) *epoch = e;
  }
}

int pa_consistent_sync(pa_vec *v) {
  pa_ctx *c = v->plan->ctx;
  const PlanPart &pp = v->plan->parts[0];
  FlagPtrs adst, asrc, ddst;
  int n = 0;
  PA_TRY(pa_sync_flags(v->plan, &adst, &asrc, &ddst, &n));
  if (!n && c->nparts == 1) return PA_OK;  // a job of one part: nothing to order, nothing to gather
  // the grid is capped so that every CTA is resident (each CTA polls the neighbours' flags before it gathers)
  int64_t g = (pp.n_cons + 255) / 256;
  g = g < 1 ? 1 : (g > 148 * 4 ? 148 * 4 : g);
  k_consistent_sync<<<(unsigned)g, 256, 0, c->stream>>>(v->d[0], pp.d_ghost_lid, pp.d_ghost_slot, pp.d_ghost_rlid, pp.n_cons, pa_peer_ptrs(v, 0),
                                                        c->d_epoch, adst, asrc, ddst, n, c->d_cons_ticket, c->d_err);
  c->launches++;
  PA_CUDA(cudaGetLastError());
  pa_mark_pending_done(v->plan);
  return PA_OK;
}

int pa_launch_consistent(pa_vec *v) {
  pa_ctx *c = v->plan->ctx;
  for (int k = 0; k < c->nlocal; ++k) {
    const PlanPart &pp = v->plan->parts[k];
    if (!pp.n_cons) continue;
    k_consistent<<<small_grid(pp.n_cons), 256, 0, c->stream>>>(v->d[k], pp.d_ghost_lid, pp.d_ghost_slot, pp.d_ghost_rlid,
                                                              pp.n_cons, pa_peer_ptrs(v, k));
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}

extern "C" int pa_vec_consistent(pa_vec *v) {
  PA_CHECK(v, PA_EINVAL, "pa_vec_consistent: null vector");
  pa_ctx *c = v->plan->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  PA_TRY(pa_before_write(c));
  if (pa_fold_ok(c)) return pa_consistent_sync(v);
  PA_TRY(pa_collective_begin(v->plan));
  PA_TRY(pa_launch_consistent(v));
  return pa_collective_end(v->plan);
}

extern "C" int pa_vec_assemble(pa_vec *v) { return pa_vec_assemble_op(v, PA_OP_SUM); }

extern "C" int pa_vec_assemble_op(pa_vec *v, int32_t op) {
  PA_CHECK(v, PA_EINVAL, "pa_vec_assemble: null vector");
  PA_CHECK(op == PA_OP_SUM || op == PA_OP_INSERT || op == PA_OP_MAX || op == PA_OP_MIN, PA_EINVAL, "pa_vec_assemble_op: unknown operation %d", op);
  pa_ctx *c = v->plan->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  PA_TRY(pa_before_write(c));
  PA_TRY(pa_collective_begin(v->plan));
  for (int k = 0; k < c->nlocal; ++k) {
    const PlanPart &pp = v->plan->parts[k];
    if (!pp.n_asm_dst) continue;
    auto kern = op == PA_OP_SUM ? k_assemble<0> : (op == PA_OP_INSERT ? k_assemble<1> : (op == PA_OP_MAX ? k_assemble<2> : k_assemble<3>));
    kern<<<small_grid(pp.n_asm_dst), 256, 0, c->stream>>>(v->d[k], pp.d_asm_dst, pp.d_asm_ptr, pp.d_asm_slot, pp.d_asm_rlid,
                                                        pp.n_asm_dst, pa_peer_ptrs(v, k));
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  PA_TRY(pa_collective_end(v->plan));
  // "After the transfer, the source ghost values are set to zero" (src/p_vector.jl:699-707): only once
  // every neighbour has finished reading them.
  PA_TRY(pa_before_write(c));
  for (int k = 0; k < c->nlocal; ++k) {
    const PlanPart &pp = v->plan->parts[k];
    if (!pp.n_ghost) continue;
    k_zero_idx<<<small_grid(pp.n_ghost), 256, 0, c->stream>>>(v->d[k], pp.prefix ? nullptr : pp.d_ghost_to_local, pp.n_own, pp.n_ghost);
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}

// ------------------------------------------------------------------ deterministic input generator
__device__ __forceinline__ double hash_uniform(uint64_t gid0, uint64_t seed) {
  uint64_t z = gid0 + seed * 0x9E3779B97F4A7C15ull;
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * 0x1.0p-52 - 1.0;
}

__global__ void k_fill_hash_box(double *v, int64_t n_own, int64_t n_local, int64_t bx, int64_t by, int64_t lox, int64_t loy,
                                int64_t loz, int64_t gnx, int64_t gny, uint64_t seed) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_local; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n_own) {
      int64_t ix = i % bx, iy = (i / bx) % by, iz = i / (bx * by);
      uint64_t gid = (uint64_t)((lox + ix) + gnx * ((loy + iy) + gny * (loz + iz)));
      v[i] = hash_uniform(gid, seed);
    } else {
      v[i] = 0.0;
    }
  }
}

extern "C" int pa_vec_fill_hash_box(pa_vec *v, int32_t k, const int64_t *gn, const int64_t *lo, const int64_t *hi, uint64_t seed) {
  PA_CHECK(v && gn && lo && hi, PA_EINVAL, "pa_vec_fill_hash_box: null argument");
  pa_ctx *c = v->plan->ctx;
  PA_CHECK(k >= 0 && k < c->nlocal, PA_EINVAL, "pa_vec_fill_hash_box: bad part");
  const PlanPart &pp = v->plan->parts[k];
  int64_t bx = hi[0] - lo[0], by = hi[1] - lo[1], bz = hi[2] - lo[2];
  PA_CHECK(pp.prefix && bx * by * bz == pp.n_own, PA_EINVAL, "pa_vec_fill_hash_box: box does not match the own block");
  PA_CUDA(cudaSetDevice(c->device));
  PA_TRY(pa_before_write(c));
  if (pp.n_local) {
    k_fill_hash_box<<<ew_grid(pp.n_local), 256, 0, c->stream>>>(v->d[k], pp.n_own, pp.n_local, bx, by, lo[0], lo[1], lo[2], gn[0], gn[1], seed);
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}
