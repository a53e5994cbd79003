/* hpcg_cg.c — the C ABI of libpa_b200.so driven from plain C (no Python, no torch): build the HPCG 27-point operator of an
 * n^3 grid on one GPU with the on-device generator, run ref_cg!(x, A, b; maxiter, Pl = Identity) and print what the
 * reference's HPCG driver prints (HPCG/src/ref_cg.jl:119-134: residual history, iterations).
 *
 *   gcc -O2 -Iinclude examples/hpcg_cg.c -Lpartitionedarrays.jl_b200/lib -lpa_b200 -Wl,-rpath,$PWD/partitionedarrays.jl_b200/lib -o hpcg_cg
 *   ./hpcg_cg [n=64] [maxiter=50]
 *
 * Exit code 0 = ran (residual reduced; the exact solution of the HPCG system is the vector of ones), 2 = no CUDA device (the
 * library has no CPU fallback), 1 = any other error. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pa_b200.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc_ = (call);                                                            \
    if (rc_ != PA_OK) {                                                          \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, pa_last_error());     \
      return strstr(pa_last_error(), "no CUDA device") ? 2 : 1;                  \
    }                                                                            \
  } while (0)

int main(int argc, char **argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 64;
  const int32_t maxiter = argc > 2 ? atoi(argv[2]) : 50;
  const int64_t rows = n * n * n;
  const int32_t part_id = 1;
  pa_ctx *ctx = NULL;
  /* one part = the whole grid; arena for x, b and the four CG work vectors */
  CHECK(pa_ctx_create(1, 1, &part_id, 0, (uint64_t)(8 * rows * 8 + (64 << 20)), NULL, &ctx));

  /* PRange of n^3 own ids, no ghosts, no neighbours (uniform_partition on a single part) */
  pa_plan *plan = NULL;
  CHECK(pa_plan_create(ctx, &plan));
  CHECK(pa_plan_set_part(plan, 0, rows, rows, NULL, NULL, 0, NULL, NULL, NULL, NULL, 0, NULL, NULL, NULL, NULL));
  CHECK(pa_plan_commit(plan, 0));

  /* A, b = build_p_matrix(...) (HPCG/src/sparse_matrix.jl:105-122): diag 26, off-diagonal -1, b = 27 - nnz_row */
  pa_mat *A = NULL;
  pa_vec *b = NULL, *x = NULL;
  CHECK(pa_vec_create(plan, &b));
  CHECK(pa_vec_create(plan, &x));
  CHECK(pa_mat_create(plan, plan, &A));
  const int64_t gn[3] = {n, n, n}, lo[3] = {0, 0, 0}, hi[3] = {n, n, n};
  CHECK(pa_mat_set_stencil(A, 0, 27, gn, lo, hi, 0, NULL, NULL, b));
  CHECK(pa_mat_commit(A));
  int64_t nnz = 0;
  CHECK(pa_mat_nnz(A, 0, &nnz));

  CHECK(pa_vec_fill(x, 0.0));
  double *hist = (double *)calloc((size_t)maxiter + 1, sizeof(double));
  pa_cg_result res;
  CHECK(pa_cg(A, x, b, maxiter, 0.0, 0, &res, hist));
  CHECK(pa_ctx_sync(ctx));

  double *xs = (double *)malloc((size_t)rows * sizeof(double));
  CHECK(pa_vec_download(x, 0, xs, rows));
  double err = 0.0;
  for (int64_t i = 0; i < rows; ++i) err = fmax(err, fabs(xs[i] - 1.0));
  int64_t launches = 0;
  CHECK(pa_ctx_launch_count(ctx, &launches));
  printf("HPCG 27-pt %lld^3: %lld rows, %lld nnz; %d CG iterations, ||r||/||r0|| = %.6e (first %.6e), max |x - 1| = %.3e, %lld kernel launches\n",
         (long long)n, (long long)rows, (long long)nnz, res.iters, res.residual / res.residual0, hist[0], err, (long long)launches);

  free(xs);
  free(hist);
  CHECK(pa_mat_destroy(A));
  CHECK(pa_vec_destroy(x));
  CHECK(pa_vec_destroy(b));
  CHECK(pa_plan_destroy(plan));
  CHECK(pa_ctx_destroy(ctx));
  /* CG on an SPD operator: the residual must have gone down and the iterate must be finite */
  return (res.iters == maxiter || res.converged) && isfinite(err) && res.residual < res.residual0 ? 0 : 1;
}
