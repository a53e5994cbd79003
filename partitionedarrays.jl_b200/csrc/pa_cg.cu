// Device-resident conjugate gradients: ref_cg!(x,A,b; Pl=Identity) of HPCG/src/ref_cg.jl:40-134.
// alpha/beta/rho never leave the GPU; the host only enqueues.  Two schedules:
//  * PA_CG_REFERENCE_OPS: the reference's sequence op for op (ldiv!=copy, dot, u.=c.+beta.*u, mul_no_lat!,
//    dot, x.+=alpha.*u, r.-=alpha.*c, norm) — 8 passes per iteration.
//  * default: same arithmetic, fewer passes: rho = ||r||^2 is carried over from the previous norm
//    (c == r under the identity preconditioner, so dot(c,r) and norm(r)^2 are the same sum), and
//    x/r updates + the new ||r||^2 are one fused pass.
#include <math.h>

#include "pa_internal.h"

__device__ __forceinline__ double cg_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// x += alpha*u ; r -= alpha*c ; partial ||r_own||^2   (alpha = *num / *den)
__global__ void __launch_bounds__(PA_RED_THREADS)
    k_cg_update(double *x, const double *u, double *r, const double *c, int64_t n_local, int64_t n_own, const double *num,
                const double *den, double *blockpart, unsigned *ticket, double *out) {
  __shared__ double sm[PA_RED_THREADS / 32];
  __shared__ bool last;
  const double alpha = *num / *den, nalpha = -alpha;
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

struct pa_mg;
extern "C" int pa_mg_apply(pa_mg *M, pa_vec *x, const pa_vec *b);

// end of a fused iteration: rotate rho_prev <- rho_cur <- ||r_new||^2, record the history, advance the device-side
// iteration counter.  With this every kernel of the iteration has iteration-independent arguments, so ONE captured
// CUDA graph replays the whole loop body (the scalars, epochs and the counter live on the device).
__global__ void k_cg_commit(double *rho_cur, double *rho_prev, const double *nrm2_new, double *hist, int *it) {
  const double v = *nrm2_new;
  *rho_prev = *rho_cur;
  *rho_cur = v;
  const int i = *it + 1;
  hist[i] = v;
  *it = i;
}

/* ref_cg!(x,A,b; Pl) with Pl = Identity (mg == NULL) or the HPCG multigrid preconditioner */
extern "C" int pa_cg_precond(pa_mat *A, pa_vec *x, const pa_vec *b, pa_mg *mg, int32_t maxiter, double tol, uint32_t flags,
                             pa_cg_result *result, double *history) {
  PA_CHECK(A && x && b && result, PA_EINVAL, "pa_cg: null argument");
  PA_CHECK(A->committed, PA_ESTATE, "pa_cg: matrix not committed");
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
  const bool ref_ops = (flags & PA_CG_REFERENCE_OPS) || !all_prefix;
  pa_vec *r = nullptr, *cv = nullptr, *u = nullptr;
  PA_TRY(pa_vec_create(x->plan, &r));
  PA_TRY(pa_vec_create(x->plan, &cv));
  PA_TRY(pa_vec_create(x->plan, &u));
  double *d_hist = nullptr, *d_rho = nullptr, *d_one = nullptr;  // ||r||^2 per iteration, rho per iteration
  PA_CUDA(cudaMalloc((void **)&d_hist, (size_t)(maxiter + 2) * sizeof(double)));
  PA_CUDA(cudaMalloc((void **)&d_rho, (size_t)(maxiter + 2) * sizeof(double)));
  d_one = c->d_scal + S_RHO0;
  const double one = 1.0;
  PA_CUDA(cudaMemcpyAsync(d_one, &one, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  double *d_uc = c->d_scal + S_UC;
  double *d_rho_cur = c->d_scal + S_RHO1, *d_rho_prev = c->d_scal + S_NRM0, *d_nrm2 = c->d_scal + S_NRM2;
  int *d_it = nullptr;
  PA_CUDA(cudaMalloc((void **)&d_it, sizeof(int)));
  PA_CUDA(cudaMemsetAsync(d_it, 0, sizeof(int), c->stream));
  int rc = PA_OK;
  std::vector<double> hist((size_t)maxiter + 1, 0.0);
  int iters = 0, converged = 0;
  double nrm0 = 0.0, nrm = 0.0;
  auto body = [&]() -> int {
    // cg_iterator! (ref_cg.jl:76-97): u .= 0 ; r = b ; c = A*x ; r .-= c ; residual = norm(r)
    PA_TRY(pa_vec_fill(u, 0.0));
    PA_TRY(pa_vec_copy(r, b));
    PA_TRY(pa_spmv(A, x, cv, 1.0, 0.0, PA_SPMV_DEFAULT));
    PA_TRY(pa_waxpby_dev(r, coef_imm(1.0), r, coef_imm(-1.0), cv));
    PA_TRY(pa_reduce_dev_to(r, nullptr, 1, d_hist));
    double h0 = 0.0;
    PA_CUDA(cudaMemcpyAsync(&h0, d_hist, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PA_CUDA(cudaStreamSynchronize(c->stream));
    nrm0 = nrm = sqrt(h0);
    hist[0] = nrm0;
    PA_CUDA(cudaMemcpyAsync(d_rho_cur, d_hist, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    PA_CUDA(cudaMemcpyAsync(d_rho_prev, &one, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (nrm0 == 0.0) {  // exact initial guess (the reference would iterate on NaNs here)
      converged = 1;
      return PA_OK;
    }
    for (int it = 0; it < maxiter; ++it) {
      if (nrm / nrm0 <= tol) {
        converged = 1;
        break;
      }
      const double *rho, *rho_prev;
      if (mg) {
        PA_TRY(pa_mg_apply(mg, cv, r));                       // ldiv!(c, Pl, r): one V-cycle
        PA_TRY(pa_reduce_dev_to(cv, r, 0, d_rho + it + 1));   // rho = dot(c,r)
        rho = d_rho + it + 1;
        rho_prev = it ? d_rho + it : d_one;
        PA_TRY(pa_waxpby_dev(u, coef_imm(1.0), cv, coef_ratio(rho, rho_prev, 1.0), u));  // u .= c .+ beta.*u
        PA_TRY(pa_spmv_dot(A, u, cv, 1.0, 0.0, PA_SPMV_DEFAULT, u, d_uc));              // c = A*u ; uc = dot(u,c)
        PA_TRY(cg_update(x, u, r, cv, rho, d_uc, d_hist + it + 1));                     // x += alpha u ; r -= alpha c ; ||r||^2
      } else if (ref_ops) {
        PA_TRY(pa_vec_copy(cv, r));                       // ldiv!(c, Identity, r)
        PA_TRY(pa_reduce_dev_to(cv, r, 0, d_rho + it + 1));  // rho = dot(c,r)
        rho = d_rho + it + 1;
        rho_prev = it ? d_rho + it : d_one;
        PA_TRY(pa_waxpby_dev(u, coef_imm(1.0), cv, coef_ratio(rho, rho_prev, 1.0), u));  // u .= c .+ beta.*u
        PA_TRY(pa_spmv(A, u, cv, 1.0, 0.0, PA_SPMV_DEFAULT));  // mul_no_lat!
        PA_TRY(pa_reduce_dev_to(u, cv, 0, d_uc));              // uc = dot(u,c)
        PA_TRY(pa_waxpby_dev(x, coef_imm(1.0), x, coef_ratio(rho, d_uc, 1.0), u));    // x .+= alpha.*u
        PA_TRY(pa_waxpby_dev(r, coef_imm(1.0), r, coef_ratio(rho, d_uc, -1.0), cv));  // r .-= alpha.*c
        PA_TRY(pa_reduce_dev_to(r, nullptr, 1, d_hist + it + 1));                     // norm(r)
      } else {
        // fused schedule; rho_cur = ||r||^2 = dot(c,r) with c == r, rho_prev from the previous iteration (1 at start)
        auto fused_iteration = [&]() -> int {
          PA_TRY(pa_waxpby_dev(u, coef_imm(1.0), r, coef_ratio(d_rho_cur, d_rho_prev, 1.0), u));
          PA_TRY(pa_spmv_dot(A, u, cv, 1.0, 0.0, PA_SPMV_SKIP_GHOST_REFRESH, u, d_uc));  // c = A*u and u.c in one pass
          PA_TRY(cg_update(x, u, r, cv, d_rho_cur, d_uc, d_nrm2));
          k_cg_commit<<<1, 1, 0, c->stream>>>(d_rho_cur, d_rho_prev, d_nrm2, d_hist, d_it);
          c->launches++;
          PA_CUDA(cudaGetLastError());
          return PA_OK;
        };
        const bool want_graph = tol <= 0.0 && it == 1 && maxiter >= 4 && pa_knob(c, "cg_graph", 1) != 0;
        if (want_graph) {
          // iteration 0 ran eagerly (warm caches, lazy allocations); capture iteration 1 once and replay it for the rest
          const int64_t l0 = c->launches;
          cudaGraph_t graph = nullptr;
          cudaGraphExec_t exec = nullptr;
          bool ok = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
          int rcap = ok ? fused_iteration() : PA_ECUDA;
          if (ok) ok = cudaStreamEndCapture(c->stream, &graph) == cudaSuccess && rcap == PA_OK && graph;
          if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
          if (ok) {
            const int64_t per_iter = c->launches - l0;
            for (int j = it; j < maxiter && ok; ++j) ok = cudaGraphLaunch(exec, c->stream) == cudaSuccess;
            c->launches = l0 + per_iter * (maxiter - it);
          }
          if (exec) cudaGraphExecDestroy(exec);
          if (graph) cudaGraphDestroy(graph);
          if (ok) {
            iters = maxiter;
            break;
          }
          cudaGetLastError();  // capture unsupported here: fall through to the eager loop
          PA_CHECK(rcap == PA_OK || rcap == PA_ECUDA, rcap, "%s", pa_last_error());
          c->launches = l0;
          c->pending_done.clear();
          PA_CUDA(cudaStreamSynchronize(c->stream));
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
    if (tol <= 0.0 && iters > 0) {
      PA_CUDA(cudaMemcpyAsync(hist.data() + 1, d_hist + 1, (size_t)iters * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      PA_CUDA(cudaStreamSynchronize(c->stream));
      for (int i = 1; i <= iters; ++i) hist[i] = sqrt(hist[i]);
      nrm = hist[iters];
    }
    if (!converged && nrm / nrm0 <= tol) converged = 1;
    return pa_check_device_error(c);
  };
  rc = body();
  cudaStreamSynchronize(c->stream);
  cudaFree(d_hist);
  cudaFree(d_rho);
  cudaFree(d_it);
  pa_vec_destroy(u);
  pa_vec_destroy(cv);
  pa_vec_destroy(r);
  if (rc != PA_OK) return rc;
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
