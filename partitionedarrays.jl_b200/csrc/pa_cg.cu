// Device-resident conjugate gradients: ref_cg!(x,A,b; Pl) of HPCG/src/ref_cg.jl:40-134.
// alpha/beta/rho never leave the GPU; the host only enqueues.  Three schedules, same arithmetic:
//  * PA_CG_REFERENCE_OPS: the reference's sequence op for op (ldiv!=copy, dot, u.=c.+beta.*u, mul_no_lat!,
//    dot, x.+=alpha.*u, r.-=alpha.*c, norm) — 8 passes per iteration.
//  * fused (Pl = Identity): rho = ||r||^2 is carried over from the previous norm (c == r under the identity
//    preconditioner, so dot(c,r) and norm(r)^2 are the same sum), dot(u,c) is the SpMV's epilogue and x/r updates +
//    the new ||r||^2 are one pass: 3 passes per iteration.
//  * fused + folded (one part per process, the production path at 1..8 GPUs): additionally the scalar all-reduces and
//    the exchange signalling ride in the prologue/epilogue of those kernels (RedPush/RedWait/DoneWait): per iteration
//    k_cg_direction, k_consistent_sync (absent on one GPU), k_spmv_tma, k_cg_update_fold — 4 launches, captured once
//    in a CUDA graph.  The work vectors, the history buffer and the graph are cached per (A, x, b).
#include <math.h>

#include <algorithm>

#include "pa_device.cuh"
#include "pa_internal.h"

__device__ __forceinline__ double cg_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// A coefficient num/den of the CG recurrences.  num == 0 means the residual is EXACTLY zero: the reference stops there
// (done(): residual/residual0 <= tol, ref_cg.jl:22-26); a device-resident loop that cannot stop turns the remaining
// iterations into no-ops instead of dividing 0/0 (alpha = beta = 0 keeps x, r and u fixed).
__device__ __forceinline__ double cg_ratio(double num, double den) { return num == 0.0 ? 0.0 : num / den; }

// x += alpha*u ; r -= alpha*c ; partial ||r_own||^2   (alpha = *num / *den)
__global__ void __launch_bounds__(PA_RED_THREADS)
    k_cg_update(double *x, const double *u, double *r, const double *c, int64_t n_local, int64_t n_own, const double *num,
                const double *den, double *blockpart, unsigned *ticket, double *out) {
  __shared__ double sm[PA_RED_THREADS / 32];
  __shared__ bool last;
  const double alpha = cg_ratio(*num, *den), nalpha = -alpha;
  double a0 = 0.0, a1 = 0.0;
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2, st = (int64_t)gridDim.x * blockDim.x * 2;
  for (; i + 1 < n_local; i += st) {
    double2 xv = *reinterpret_cast<double2 *>(x + i), uv = *reinterpret_cast<const double2 *>(u + i);
    double2 rv = *reinterpret_cast<double2 *>(r + i), cv = *reinterpret_cast<const double2 *>(c + i);
    xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, uv.x));
    xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, uv.y));
    rv.x = __dadd_rn(rv.x, __dmul_rn(nalpha, cv.x));
    rv.y = __dadd_rn(rv.y, __dmul_rn(nalpha, cv.y));
    *reinterpret_cast<double2 *>(x + i) = xv;
    *reinterpret_cast<double2 *>(r + i) = rv;
    if (i < n_own) a0 = fma(rv.x, rv.x, a0);
    if (i + 1 < n_own) a1 = fma(rv.y, rv.y, a1);
  }
  if (i < n_local) {
    x[i] = __dadd_rn(x[i], __dmul_rn(alpha, u[i]));
    const double rr = __dadd_rn(r[i], __dmul_rn(nalpha, c[i]));
    r[i] = rr;
    if (i < n_own) a0 = fma(rr, rr, a0);
  }
  double s = cg_warp_sum(a0 + a1);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sm[w] = s;
  __syncthreads();
  if (w == 0) {
    s = (l < PA_RED_THREADS / 32) ? sm[l] : 0.0;
    s = cg_warp_sum(s);
    if (l == 0) {
      blockpart[blockIdx.x] = s;
      __threadfence();
      unsigned t = atomicInc(ticket, gridDim.x - 1);
      last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double a = 0.0;
    for (unsigned j = threadIdx.x; j < gridDim.x; j += blockDim.x) a += __ldcg(blockpart + j);
    a = cg_warp_sum(a);
    __syncthreads();
    if (l == 0) sm[w] = a;
    __syncthreads();
    if (w == 0) {
      a = (l < PA_RED_THREADS / 32) ? sm[l] : 0.0;
      a = cg_warp_sum(a);
      if (l == 0) *out = a;
    }
  }
}

static int cg_update(pa_vec *x, const pa_vec *u, pa_vec *r, const pa_vec *cvec, const double *num, const double *den, double *d_out) {
  pa_ctx *c = x->plan->ctx;
  PA_TRY(pa_before_write(c));
  for (int k = 0; k < c->nlocal; ++k) {
    const PlanPart &pp = x->plan->parts[k];
    int64_t per = 2 * PA_RED_THREADS * 4;
    int grid = (int)((pp.n_local + per - 1) / per);
    grid = grid < 1 ? 1 : (grid > PA_RED_BLOCKS ? PA_RED_BLOCKS : grid);
    double *out = c->nlocal == 1 ? d_out : c->d_partial + k;
    k_cg_update<<<grid, PA_RED_THREADS, 0, c->stream>>>(x->d[k], u->d[k], r->d[k], cvec->d[k], pp.n_local, pp.n_own, num, den,
                                                        c->d_blockpart + (size_t)k * PA_RED_BLOCKS, c->d_ticket + k, out);
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  return pa_reduce_finish(c, d_out);
}

// ------------------------------------------------------------------ the folded schedule (one part per process)
// hist[i] = ||r_i||^2 (= rho_i under Pl = Identity) on every part; *it = iterations completed.
// Iteration i:  k_cg_direction  every CTA: rho_i = (i == 0 ? hist[0] : sum over parts of the posted ||r_i||^2 partials);
//                               CTA 0 records hist[i]; beta = rho_i / rho_{i-1} (1 at i == 0); u = r + beta*u
//               k_consistent_sync(u) (jobs of several parts), k_spmv_tma: c = A*u, last CTA posts the part's u.c
//               k_cg_update_fold every CTA: waits for the neighbours' "done reading u" and for all parts' u.c;
//                               alpha = rho_i / u.c; x += alpha*u; r -= alpha*c; last CTA posts ||r_{i+1}||^2, ++*it
__global__ void __launch_bounds__(PA_RED_THREADS)
    k_cg_direction(double *u, const double *r, int64_t n, RedWait w, double *hist, const int *it) {
  __shared__ double sm[PA_MAX_NBR + 1];
  const int i = *it;
  double rho, rho_prev = 1.0;
  if (i == 0) {
    rho = hist[0];
  } else {
    rho = pa_red_sum(w, sm);
    rho_prev = hist[i - 1];
    if (blockIdx.x == 0 && threadIdx.x == 0) hist[i] = rho;
  }
  const double beta = cg_ratio(rho, rho_prev);
  int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2, st = (int64_t)gridDim.x * blockDim.x * 2;
  for (; j + 1 < n; j += st) {
    double2 rv = *reinterpret_cast<const double2 *>(r + j), uv = *reinterpret_cast<double2 *>(u + j);
    // u .= c .+ beta.*u  evaluated as 1.0*c + beta*u (mul, mul, add: the arithmetic of k_waxpby)
    uv.x = __dadd_rn(__dmul_rn(1.0, rv.x), __dmul_rn(beta, uv.x));
    uv.y = __dadd_rn(__dmul_rn(1.0, rv.y), __dmul_rn(beta, uv.y));
    *reinterpret_cast<double2 *>(u + j) = uv;
  }
  if (j < n) u[j] = __dadd_rn(__dmul_rn(1.0, r[j]), __dmul_rn(beta, u[j]));
}

// The same update with consistent!(u) fused in (jobs of several parts): the ghost exchange rides inside the kernel that
// PRODUCES u, so the SpMV that follows is purely local and no exchange latency is exposed:
//   A  every CTA updates its share of the BOUNDARY entries of u (the own entries some neighbour reads: the plan's assemble
//      destinations); the last CTA through the ticket publishes "my boundary is final" to the neighbours (system scope);
//   B  the bulk: all other own entries (a bitmap marks the boundary entries: they must not be updated twice) — ~0.5 ms of
//      pure HBM streaming during which the neighbours' boundaries become final;
//   C  every CTA waits for the neighbours' signals (they arrived long ago) and pulls its share of the ghost values from
//      the owners' HBM over NVLink into u's ghost slots; the last CTA tells the neighbours "done reading" and advances
//      the exchange epoch.
struct XchgArgs {
  const int32_t *bnd;        // boundary local ids
  int64_t n_bnd;
  const uint32_t *bitmap;    // bit l: local entry l is a boundary entry
  const int32_t *lid, *slot, *rlid;  // the consistent! table (ghost slot <- neighbour slot, local id on the owner)
  int64_t n_cons;
  PeerPtrs peers;
  unsigned long long *epoch;
  FlagPtrs arrive_dst, arrive_src, done_dst;
  int nnbr;
  unsigned *tickets;         // [2]
  int *err;
};

__global__ void __launch_bounds__(PA_RED_THREADS)
    k_cg_direction_xchg(double *u, const double *r, int64_t n_own, RedWait w, double *hist, const int *it, const XchgArgs xa) {
  __shared__ double sm[PA_MAX_NBR + 1];
  __shared__ bool last;
  const int i = *it;
  double rho, rho_prev = 1.0;
  if (i == 0) {
    rho = hist[0];
  } else {
    rho = pa_red_sum(w, sm);
    rho_prev = hist[i - 1];
    if (blockIdx.x == 0 && threadIdx.x == 0) hist[i] = rho;
  }
  const double beta = cg_ratio(rho, rho_prev);
  const unsigned long long e = *xa.epoch + 1ull;  // nobody writes *epoch before every CTA has passed the second ticket
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  // ---- A: boundary entries first
  for (int64_t j = tid; j < xa.n_bnd; j += nth) {
    const int32_t l = xa.bnd[j];
    u[l] = __dadd_rn(__dmul_rn(1.0, r[l]), __dmul_rn(beta, u[l]));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();  // this CTA's boundary stores, before it is counted in
    last = atomicInc(xa.tickets, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && (int)threadIdx.x < xa.nnbr) {
    __threadfence_system();
    pa_st_release_sys(xa.arrive_dst.p[threadIdx.x], e);
  }
  // ---- B: the bulk of the own entries (boundary entries are skipped)
  // (the loads do not wait for the bitmap word: r and u are read for every pair, only the STORE of a boundary entry is
  // suppressed — a boundary entry of u read here may be the old or the new value, it is not used)
  int64_t j = tid * 2;
  for (; j + 1 < n_own; j += nth * 2) {
    const unsigned bits = xa.bitmap ? ((__ldg(xa.bitmap + (j >> 5)) >> (j & 31)) & 3u) : 0u;  // j is even: both bits in one word
    const double2 rv = *reinterpret_cast<const double2 *>(r + j);
    double2 uv = *reinterpret_cast<double2 *>(u + j);
    uv.x = __dadd_rn(__dmul_rn(1.0, rv.x), __dmul_rn(beta, uv.x));
    uv.y = __dadd_rn(__dmul_rn(1.0, rv.y), __dmul_rn(beta, uv.y));
    if (bits == 0u) {
      *reinterpret_cast<double2 *>(u + j) = uv;
    } else {
      if (!(bits & 1u)) u[j] = uv.x;
      if (!(bits & 2u)) u[j + 1] = uv.y;
    }
  }
  if (j < n_own) {
    const unsigned bit = xa.bitmap ? ((xa.bitmap[j >> 5] >> (j & 31)) & 1u) : 0u;
    if (!bit) u[j] = __dadd_rn(__dmul_rn(1.0, r[j]), __dmul_rn(beta, u[j]));
  }
  // ---- C: the ghost values
  if ((int)threadIdx.x < xa.nnbr) pa_spin_until(xa.arrive_src.p[threadIdx.x], e, xa.err);
  __syncthreads();
  for (int64_t j = tid; j < xa.n_cons; j += nth) u[xa.lid[j]] = __ldcg(xa.peers.p[xa.slot[j]] + xa.rlid[j]);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicInc(xa.tickets + 1, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {
    if ((int)threadIdx.x < xa.nnbr) pa_st_release_sys(xa.done_dst.p[threadIdx.x], e);
    if (threadIdx.x == 0) *xa.epoch = e;
  }
}

__global__ void __launch_bounds__(PA_RED_THREADS)
    k_cg_update_fold(double *x, const double *u, double *r, const double *c, int64_t n_local, int64_t n_own, RedWait w, DoneWait dw,
                     RedPush push, const double *hist, int *it, double *blockpart, unsigned *ticket) {
  __shared__ double sm[PA_MAX_NBR + 1];
  __shared__ double red[PA_RED_THREADS / 32];
  __shared__ bool last;
  pa_wait_done(dw);  // r and x are not read by the neighbours, u is: conservative, and free (the flags arrived long ago)
  const double uc = pa_red_sum(w, sm);
  const int i = *it;
  const double alpha = cg_ratio(hist[i], uc), nalpha = -alpha;
  double a0 = 0.0, a1 = 0.0;
  int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2, st = (int64_t)gridDim.x * blockDim.x * 2;
  for (; j + 1 < n_local; j += st) {
    double2 xv = *reinterpret_cast<double2 *>(x + j), uv = *reinterpret_cast<const double2 *>(u + j);
    double2 rv = *reinterpret_cast<double2 *>(r + j), cv = *reinterpret_cast<const double2 *>(c + j);
    xv.x = __dadd_rn(xv.x, __dmul_rn(alpha, uv.x));
    xv.y = __dadd_rn(xv.y, __dmul_rn(alpha, uv.y));
    rv.x = __dadd_rn(rv.x, __dmul_rn(nalpha, cv.x));
    rv.y = __dadd_rn(rv.y, __dmul_rn(nalpha, cv.y));
    *reinterpret_cast<double2 *>(x + j) = xv;
    *reinterpret_cast<double2 *>(r + j) = rv;
    if (j < n_own) a0 = fma(rv.x, rv.x, a0);
    if (j + 1 < n_own) a1 = fma(rv.y, rv.y, a1);
  }
  if (j < n_local) {
    x[j] = __dadd_rn(x[j], __dmul_rn(alpha, u[j]));
    const double rr = __dadd_rn(r[j], __dmul_rn(nalpha, c[j]));
    r[j] = rr;
    if (j < n_own) a0 = fma(rr, rr, a0);
  }
  double s = cg_warp_sum(a0 + a1);
  const int wp = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[wp] = s;
  __syncthreads();
  if (wp == 0) {
    s = (l < PA_RED_THREADS / 32) ? red[l] : 0.0;
    s = cg_warp_sum(s);
    if (l == 0) {
      blockpart[blockIdx.x] = s;
      __threadfence();
      last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double a = 0.0;
    for (unsigned q = threadIdx.x; q < gridDim.x; q += blockDim.x) a += __ldcg(blockpart + q);
    a = cg_warp_sum(a);
    __syncthreads();
    if (l == 0) red[wp] = a;
    __syncthreads();
    if (wp == 0) {
      a = (l < PA_RED_THREADS / 32) ? red[l] : 0.0;
      a = cg_warp_sum(a);
      pa_red_post(push, a, l);  // ||r_{i+1}||^2 of this part -> every part
      if (l == 0) *it = i + 1;
    }
  }
}

// after the loop (or per iteration when the host checks the tolerance): hist[*it] = sum over parts of the last post
__global__ void k_cg_finish(RedWait w, double *hist, const int *it) {
  __shared__ double sm[PA_MAX_NBR + 1];
  const double rho = pa_red_sum(w, sm);
  if (threadIdx.x == 0) hist[*it] = rho;
}

// ------------------------------------------------------------------ cached workspaces
struct CgWork {
  uint64_t a_uid = 0, x_uid = 0, b_uid = 0;
  pa_plan *plan = nullptr;
  pa_vec *r = nullptr, *cv = nullptr, *u = nullptr;
  double *d_hist = nullptr, *d_rho = nullptr;
  int *d_it = nullptr;
  int cap = 0;  // iterations the history buffers hold
  cudaGraphExec_t exec = nullptr;
  int64_t exec_launches = 0;  // kernels per replay
  int exec_kind = 0;          // 1 generic fused, 2 folded
};

static void cg_free_work(CgWork *w) {
  if (w->exec) cudaGraphExecDestroy(w->exec);
  cudaFree(w->d_hist);
  cudaFree(w->d_rho);
  cudaFree(w->d_it);
  pa_vec_destroy(w->u);  // reverse creation order (symmetric heap discipline)
  pa_vec_destroy(w->cv);
  pa_vec_destroy(w->r);
  delete w;
}

void pa_cg_drop_work(pa_ctx *c, const pa_mat *A, const pa_plan *plan, bool all) {
  auto &v = c->cg_work;
  for (size_t i = 0; i < v.size();) {
    if (all || (A && v[i]->a_uid == A->uid) || (plan && v[i]->plan == plan)) {
      cg_free_work(v[i]);
      v.erase(v.begin() + i);
    } else {
      ++i;
    }
  }
}

/* Release every cached CG workspace of the context (three work vectors per (A, x, b) stay in the arena otherwise). */
extern "C" int pa_ctx_release_workspaces(pa_ctx *c) {
  PA_CHECK(c, PA_EINVAL, "pa_ctx_release_workspaces: null context");
  PA_CUDA(cudaSetDevice(c->device));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  pa_cg_drop_work(c, nullptr, nullptr, true);
  return PA_OK;
}

static int cg_get_work(pa_ctx *c, pa_mat *A, pa_vec *x, const pa_vec *b, int maxiter, CgWork **out) {
  CgWork *w = nullptr;
  for (CgWork *q : c->cg_work)
    if (q->a_uid == A->uid && q->x_uid == x->uid && q->b_uid == b->uid && q->plan == x->plan) w = q;
  if (!w) {
    // at most two workspaces per matrix (a caller that alternates between two (x, b) pairs to overlap transfers with solves
    // keeps both graphs); a third pair replaces the older one — the arena holds 3 vectors per workspace
    int cnt = 0;
    for (CgWork *q : c->cg_work) cnt += q->a_uid == A->uid;
    if (cnt >= 2) {
      for (size_t i = 0; i < c->cg_work.size(); ++i)
        if (c->cg_work[i]->a_uid == A->uid) {
          cg_free_work(c->cg_work[i]);
          c->cg_work.erase(c->cg_work.begin() + i);
          break;
        }
    }
    w = new CgWork();
    w->a_uid = A->uid;
    w->x_uid = x->uid;
    w->b_uid = b->uid;
    w->plan = x->plan;
    int rc = pa_vec_create(x->plan, &w->r);
    if (rc == PA_OK) rc = pa_vec_create(x->plan, &w->cv);
    if (rc == PA_OK) rc = pa_vec_create(x->plan, &w->u);
    if (rc == PA_OK && cudaMalloc((void **)&w->d_it, sizeof(int)) != cudaSuccess) rc = PA_ENOMEM;
    if (rc != PA_OK) {
      cg_free_work(w);
      return rc;
    }
    c->cg_work.push_back(w);
  }
  if (w->cap < maxiter) {
    cudaFree(w->d_hist);
    cudaFree(w->d_rho);
    w->d_hist = w->d_rho = nullptr;
    w->cap = 0;
    PA_CUDA(cudaMalloc((void **)&w->d_hist, (size_t)(maxiter + 2) * sizeof(double)));
    PA_CUDA(cudaMalloc((void **)&w->d_rho, (size_t)(maxiter + 2) * sizeof(double)));
    w->cap = maxiter;
  }
  *out = w;
  return PA_OK;
}

struct pa_mg;
extern "C" int pa_mg_apply(pa_mg *M, pa_vec *x, const pa_vec *b);

// end of a generic fused iteration: rotate rho_prev <- rho_cur <- ||r_new||^2, record the history, advance the
// device-side iteration counter (iteration-independent kernel arguments: one captured graph replays the loop body)
__global__ void k_cg_commit(double *rho_cur, double *rho_prev, const double *nrm2_new, double *hist, int *it) {
  const double v = *nrm2_new;
  *rho_prev = *rho_cur;
  *rho_cur = v;
  const int i = *it + 1;
  hist[i] = v;
  *it = i;
}

// per-operation device times of one solve (PA_CG_TIMING): an event after every operation, binned by category
enum { T_DDOT = 0, T_WAXPBY = 1, T_SPMV = 2, T_PRECOND = 3, T_NCAT = 4 };
struct OpTimer {
  bool on = false;
  cudaStream_t st = nullptr;
  std::vector<cudaEvent_t> ev;
  std::vector<int> cat;
  void start() {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
    cat.push_back(-1);
  }
  void mark(int category) {  // the operation enqueued since the previous mark belongs to `category`
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
    cat.push_back(category);
  }
  void finish(double *out /* [T_NCAT + 1] ms; last = total */) {
    if (!on) return;
    cudaEventSynchronize(ev.back());
    for (int i = 0; i <= T_NCAT; ++i) out[i] = 0.0;
    for (size_t i = 1; i < ev.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
      if (cat[i] >= 0) out[cat[i]] += ms;
      out[T_NCAT] += ms;
    }
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear();
  }
};

/* ref_cg!(x,A,b; Pl) with Pl = Identity (mg == NULL) or the HPCG multigrid preconditioner */
extern "C" int pa_cg_precond(pa_mat *A, pa_vec *x, const pa_vec *b, pa_mg *mg, int32_t maxiter, double tol, uint32_t flags,
                             pa_cg_result *result, double *history) {
  PA_CHECK(A && x && b && result, PA_EINVAL, "pa_cg: null argument");
  PA_CHECK(A->committed, PA_ESTATE, "pa_cg: matrix not committed");
  PA_CHECK(!A->subassembled, PA_EINVAL, "pa_cg: needs an assembled matrix");
  PA_CHECK(maxiter >= 0, PA_EINVAL, "pa_cg: negative maxiter");
  pa_ctx *c = A->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  bool all_prefix = true;
  for (int k = 0; k < c->nlocal; ++k) {
    const PlanPart &xp = x->plan->parts[k], &bp = b->plan->parts[k], &cp = A->cols->parts[k], &rp = A->rows->parts[k];
    PA_CHECK(xp.n_local == cp.n_local && xp.n_own == cp.n_own && bp.n_local == xp.n_local && bp.n_own == xp.n_own && rp.n_own == cp.n_own,
             PA_EINVAL, "pa_cg: x, b and A do not share one square partition on part %d", c->part_ids[k] + 1);
    all_prefix = all_prefix && xp.prefix;
  }
  const bool timing = (flags & PA_CG_TIMING) != 0;
  const bool ref_ops = (flags & PA_CG_REFERENCE_OPS) || !all_prefix;
  CgWork *W = nullptr;
  PA_TRY(cg_get_work(c, A, x, b, maxiter, &W));
  pa_vec *r = W->r, *cv = W->cv, *u = W->u;
  double *d_hist = W->d_hist, *d_rho = W->d_rho;
  int *d_it = W->d_it;
  double *d_one = c->d_scal + S_RHO0;
  const double one = 1.0;
  PA_CUDA(cudaMemcpyAsync(d_one, &one, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  double *d_uc = c->d_scal + S_UC;
  double *d_rho_cur = c->d_scal + S_RHO1, *d_rho_prev = c->d_scal + S_NRM0, *d_nrm2 = c->d_scal + S_NRM2;
  PA_CUDA(cudaMemsetAsync(d_it, 0, sizeof(int), c->stream));
  std::vector<double> hist((size_t)maxiter + 1, 0.0);
  int iters = 0, converged = 0;
  double nrm0 = 0.0, nrm = 0.0;
  OpTimer tm;
  tm.on = timing;
  tm.st = c->stream;
  // the folded schedule: one local part with mapped peers, the dot epilogue available, Pl = Identity
  const bool folded = !mg && !ref_ops && !timing && pa_spmv_dot_foldable(A, u) && pa_knob(c, "cg_fold", 1) != 0;
  auto body = [&]() -> int {
    // cg_iterator! (ref_cg.jl:76-97): u .= 0 ; r = b ; c = A*x ; r .-= c ; residual = norm(r)
    tm.start();
    PA_TRY(pa_vec_fill(u, 0.0));
    PA_TRY(pa_vec_copy(r, b));
    tm.mark(T_WAXPBY);
    PA_TRY(pa_spmv(A, x, cv, 1.0, 0.0, PA_SPMV_DEFAULT));
    tm.mark(T_SPMV);
    PA_TRY(pa_waxpby_dev(r, coef_imm(1.0), r, coef_imm(-1.0), cv));
    tm.mark(T_WAXPBY);
    PA_TRY(pa_reduce_dev_to(r, nullptr, 1, d_hist));
    tm.mark(T_DDOT);
    double h0 = 0.0;
    PA_CUDA(cudaMemcpyAsync(&h0, d_hist, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PA_CUDA(cudaStreamSynchronize(c->stream));
    nrm0 = nrm = sqrt(h0);
    hist[0] = nrm0;
    PA_CUDA(cudaMemcpyAsync(d_rho_cur, d_hist, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    PA_CUDA(cudaMemcpyAsync(d_rho_prev, &one, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (nrm0 == 0.0) {  // exact initial guess: done() holds before the first iteration (0/0 <= tol is false in the reference,
      converged = 1;    // which would then iterate on NaNs; we stop)
      return PA_OK;
    }
    if (folded) {
      const PlanPart &xp = x->plan->parts[0];
      const RedWait rw = pa_red_wait(c);
      const RedPush rpush = pa_red_push(c);
      const DoneWait dw = pa_done_wait(x->plan);
      const int64_t per = 2 * PA_RED_THREADS * 4;
      int ugrid = (int)((xp.n_local + per - 1) / per);
      ugrid = ugrid < 1 ? 1 : (ugrid > PA_RED_BLOCKS ? PA_RED_BLOCKS : ugrid);
      // consistent!(u) fused into the kernel that produces u (several parts, own-first layout, default mul! schedule)
      const bool fuse_x = c->nparts > 1 && xp.prefix && pa_knob(c, "cg_fuse_exchange", 1) != 0 && pa_knob(c, "spmv_strategy", -1) <= 0;
      XchgArgs xa;
      if (fuse_x) {
        xa.bnd = xp.d_asm_dst;
        xa.n_bnd = xp.n_asm_dst;
        xa.bitmap = xp.d_bnd_bitmap;
        xa.lid = xp.d_ghost_lid;
        xa.slot = xp.d_ghost_slot;
        xa.rlid = xp.d_ghost_rlid;
        xa.n_cons = xp.n_cons;
        xa.peers = pa_peer_ptrs(u, 0);
        xa.epoch = c->d_epoch;
        PA_TRY(pa_sync_flags(x->plan, &xa.arrive_dst, &xa.arrive_src, &xa.done_dst, &xa.nnbr));
        xa.tickets = c->d_cons_ticket;
        xa.err = c->d_err;
      }
      int xgrid = ugrid;
      if (fuse_x) {  // every CTA must be resident: a CTA waits (phase C) for signals that need all CTAs of the peers' grids
        int per_sm = 0, nsm = 148;
        PA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_direction_xchg, PA_RED_THREADS, 0));
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
        xgrid = std::max(1, std::min(ugrid, nsm * std::max(per_sm, 1)));
      }
      auto iteration = [&]() -> int {
        PA_TRY(pa_before_write(c));  // (pending "done" of the setup product; nothing inside the loop)
        if (fuse_x) {
          k_cg_direction_xchg<<<xgrid, PA_RED_THREADS, 0, c->stream>>>(u->d[0], r->d[0], xp.n_own, rw, d_hist, d_it, xa);
          c->launches++;
          PA_TRY(pa_spmv_local(A, u, cv, 1.0, 0.0, 0, u, nullptr, /*fold=*/1));  // purely local: the ghosts are in place
          k_cg_update_fold<<<ugrid, PA_RED_THREADS, 0, c->stream>>>(x->d[0], u->d[0], r->d[0], cv->d[0], xp.n_local, xp.n_own, rw, dw, rpush,
                                                                    d_hist, d_it, c->d_blockpart, c->d_ticket);
          c->launches++;
          PA_CUDA(cudaGetLastError());
          return PA_OK;
        }
        k_cg_direction<<<ugrid, PA_RED_THREADS, 0, c->stream>>>(u->d[0], r->d[0], xp.n_local, rw, d_hist, d_it);
        c->launches++;
        PA_TRY(pa_spmv_dot(A, u, cv, 1.0, 0.0, PA_SPMV_SKIP_GHOST_REFRESH, u, nullptr, /*fold=*/1));
        c->pending_done.clear();  // the neighbours' "done reading u" is awaited in the prologue of k_cg_update_fold
        k_cg_update_fold<<<ugrid, PA_RED_THREADS, 0, c->stream>>>(x->d[0], u->d[0], r->d[0], cv->d[0], xp.n_local, xp.n_own, rw, dw, rpush,
                                                                  d_hist, d_it, c->d_blockpart, c->d_ticket);
        c->launches++;
        PA_CUDA(cudaGetLastError());
        return PA_OK;
      };
      auto finish = [&]() -> int {
        k_cg_finish<<<1, 32, 0, c->stream>>>(rw, d_hist, d_it);
        c->launches++;
        PA_CUDA(cudaGetLastError());
        return PA_OK;
      };
      int it = 0;
      if (tol <= 0.0 && maxiter >= 4 && pa_knob(c, "cg_graph", 1) != 0) {
        PA_TRY(iteration());  // iteration 0 eagerly (lazy allocations, tile tables)
        it = 1;
        const int kind_id = fuse_x ? 3 : 2;
        if (!(W->exec && W->exec_kind == kind_id)) {
          if (W->exec) cudaGraphExecDestroy(W->exec);
          W->exec = nullptr;
          const int64_t l0 = c->launches;
          cudaGraph_t graph = nullptr;
          bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
          int rcap = ok ? iteration() : PA_ECUDA;
          if (ok) ok = cudaStreamEndCapture(c->stream, &graph) == cudaSuccess && rcap == PA_OK && graph;
          if (ok) ok = cudaGraphInstantiate(&W->exec, graph, 0) == cudaSuccess;
          if (graph) cudaGraphDestroy(graph);
          W->exec_launches = c->launches - l0;
          c->launches = l0;
          if (!ok) {
            cudaGetLastError();
            W->exec = nullptr;
            PA_CHECK(rcap == PA_OK || rcap == PA_ECUDA, rcap, "%s", pa_last_error());
            c->pending_done.clear();
            PA_CUDA(cudaStreamSynchronize(c->stream));
          } else {
            W->exec_kind = kind_id;
          }
        }
        if (W->exec) {
          for (; it < maxiter; ++it) PA_CUDA(cudaGraphLaunch(W->exec, c->stream));
          c->launches += W->exec_launches * (maxiter - 1);
        }
      }
      for (; it < maxiter; ++it) {
        if (tol > 0.0) {
          if (nrm / nrm0 <= tol) {
            converged = 1;
            break;
          }
        }
        PA_TRY(iteration());
        if (tol > 0.0) {  // the reference checks every iteration (done(), ref_cg.jl:22-26)
          PA_TRY(finish());
          double h = 0.0;
          PA_CUDA(cudaMemcpyAsync(&h, d_hist + it + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
          PA_CUDA(cudaStreamSynchronize(c->stream));
          nrm = sqrt(h);
          hist[it + 1] = nrm;
        }
      }
      iters = it;
      if (tol <= 0.0 && iters > 0) PA_TRY(finish());
    } else {
      for (int it = 0; it < maxiter; ++it) {
        if (nrm / nrm0 <= tol) {
          converged = 1;
          break;
        }
        const double *rho, *rho_prev;
        if (mg || ref_ops) {
          if (mg) {
            PA_TRY(pa_mg_apply(mg, cv, r));  // ldiv!(c, Pl, r): one V-cycle
            tm.mark(T_PRECOND);
          } else {
            PA_TRY(pa_vec_copy(cv, r));  // ldiv!(c, Identity, r)
            tm.mark(T_PRECOND);
          }
          PA_TRY(pa_reduce_dev_to(cv, r, 0, d_rho + it + 1));  // rho = dot(c,r)
          tm.mark(T_DDOT);
          rho = d_rho + it + 1;
          rho_prev = it ? d_rho + it : d_one;
          PA_TRY(pa_waxpby_dev(u, coef_imm(1.0), cv, coef_ratio(rho, rho_prev, 1.0), u));  // u .= c .+ beta.*u
          tm.mark(T_WAXPBY);
          if (ref_ops || timing) {
            PA_TRY(pa_spmv(A, u, cv, 1.0, 0.0, PA_SPMV_DEFAULT));  // mul_no_lat!
            tm.mark(T_SPMV);
            PA_TRY(pa_reduce_dev_to(u, cv, 0, d_uc));  // uc = dot(u,c)
            tm.mark(T_DDOT);
            PA_TRY(pa_waxpby_dev(x, coef_imm(1.0), x, coef_ratio(rho, d_uc, 1.0), u));    // x .+= alpha.*u
            PA_TRY(pa_waxpby_dev(r, coef_imm(1.0), r, coef_ratio(rho, d_uc, -1.0), cv));  // r .-= alpha.*c
            tm.mark(T_WAXPBY);
            PA_TRY(pa_reduce_dev_to(r, nullptr, 1, d_hist + it + 1));  // norm(r)
            tm.mark(T_DDOT);
          } else {
            PA_TRY(pa_spmv_dot(A, u, cv, 1.0, 0.0, PA_SPMV_DEFAULT, u, d_uc));  // c = A*u ; uc = dot(u,c)
            PA_TRY(cg_update(x, u, r, cv, rho, d_uc, d_hist + it + 1));         // x += alpha u ; r -= alpha c ; ||r||^2
          }
        } else {
          // generic fused schedule; rho_cur = ||r||^2 = dot(c,r) with c == r, rho_prev from the previous iteration (1 at start)
          auto fused_iteration = [&]() -> int {
            PA_TRY(pa_waxpby_dev(u, coef_imm(1.0), r, coef_ratio(d_rho_cur, d_rho_prev, 1.0), u));
            tm.mark(T_WAXPBY);
            PA_TRY(pa_spmv_dot(A, u, cv, 1.0, 0.0, PA_SPMV_SKIP_GHOST_REFRESH, u, d_uc));  // c = A*u and u.c in one pass
            tm.mark(T_SPMV);
            PA_TRY(cg_update(x, u, r, cv, d_rho_cur, d_uc, d_nrm2));
            k_cg_commit<<<1, 1, 0, c->stream>>>(d_rho_cur, d_rho_prev, d_nrm2, d_hist, d_it);
            c->launches++;
            PA_CUDA(cudaGetLastError());
            tm.mark(T_WAXPBY);
            return PA_OK;
          };
          const bool want_graph = !timing && tol <= 0.0 && it == 1 && maxiter >= 4 && pa_knob(c, "cg_graph", 1) != 0;
          if (want_graph) {
            // iteration 0 ran eagerly (warm caches, lazy allocations); capture iteration 1 once and replay it for the rest
            const int64_t l0 = c->launches;
            if (!(W->exec && W->exec_kind == 1)) {
              if (W->exec) cudaGraphExecDestroy(W->exec);
              W->exec = nullptr;
              cudaGraph_t graph = nullptr;
              bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
              int rcap = ok ? fused_iteration() : PA_ECUDA;
              if (ok) ok = cudaStreamEndCapture(c->stream, &graph) == cudaSuccess && rcap == PA_OK && graph;
              if (ok) ok = cudaGraphInstantiate(&W->exec, graph, 0) == cudaSuccess;
              if (graph) cudaGraphDestroy(graph);
              W->exec_launches = c->launches - l0;
              c->launches = l0;
              if (!ok) {
                cudaGetLastError();  // capture unsupported here: fall through to the eager loop
                W->exec = nullptr;
                PA_CHECK(rcap == PA_OK || rcap == PA_ECUDA, rcap, "%s", pa_last_error());
                c->pending_done.clear();
                PA_CUDA(cudaStreamSynchronize(c->stream));
              } else {
                W->exec_kind = 1;
              }
            }
            if (W->exec) {
              bool ok = true;
              for (int j = it; j < maxiter && ok; ++j) ok = cudaGraphLaunch(W->exec, c->stream) == cudaSuccess;
              PA_CHECK(ok, PA_ECUDA, "pa_cg: cudaGraphLaunch failed (%s)", cudaGetErrorString(cudaGetLastError()));
              c->launches = l0 + W->exec_launches * (maxiter - it);
              iters = maxiter;
              break;
            }
          }
          PA_TRY(fused_iteration());
        }
        iters = it + 1;
        if (tol > 0.0) {  // the reference checks every iteration (done(), ref_cg.jl:22-26)
          double h = 0.0;
          PA_CUDA(cudaMemcpyAsync(&h, d_hist + it + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
          PA_CUDA(cudaStreamSynchronize(c->stream));
          nrm = sqrt(h);
          hist[it + 1] = nrm;
        }
      }
    }
    if (tol <= 0.0 && iters > 0) {
      PA_CUDA(cudaMemcpyAsync(hist.data() + 1, d_hist + 1, (size_t)iters * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      PA_CUDA(cudaStreamSynchronize(c->stream));
      for (int i = 1; i <= iters; ++i) hist[i] = sqrt(hist[i]);
      // the loop never looks at the residual when tol <= 0; the reference does and stops at an EXACT zero
      // (0/residual0 <= 0).  The device loop made the later iterations no-ops: report the iteration it happened in.
      for (int i = 1; i <= iters; ++i)
        if (hist[i] == 0.0) {
          iters = i;
          converged = 1;
          break;
        }
      nrm = hist[iters];
    }
    if (!converged && nrm / nrm0 <= tol) converged = 1;
    if (timing) {
      tm.finish(c->cg_timing);
      c->cg_timing[T_NCAT + 1] = (double)iters;
    }
    return pa_check_device_error(c);
  };
  int rc = body();
  cudaStreamSynchronize(c->stream);
  if (rc != PA_OK) {
    pa_cg_drop_work(c, A, nullptr, false);  // do not keep a workspace (or a graph) of a failed solve
    return rc;
  }
  result->iters = iters;
  result->converged = converged;
  result->residual0 = nrm0;
  result->residual = nrm;
  if (history)
    for (int i = 0; i <= maxiter; ++i) history[i] = i <= iters ? hist[i] : 0.0;
  return PA_OK;
}

extern "C" int pa_cg(pa_mat *A, pa_vec *x, const pa_vec *b, int32_t maxiter, double tol, uint32_t flags, pa_cg_result *result,
                     double *history) {
  return pa_cg_precond(A, x, b, nullptr, maxiter, tol, flags, result, history);
}

/* Device times of the last solve run with PA_CG_TIMING, in ms: [0] DDOT, [1] WAXPBY, [2] SPMV, [3] preconditioner
 * (ldiv!), [4] total of all timed operations, [5] iterations — the `timing_data` slots of HPCG/src/ref_cg.jl:46-67. */
extern "C" int pa_cg_timings(pa_ctx *c, double *out6) {
  PA_CHECK(c && out6, PA_EINVAL, "pa_cg_timings: null argument");
  for (int i = 0; i < 6; ++i) out6[i] = c->cg_timing[i];
  return PA_OK;
}
