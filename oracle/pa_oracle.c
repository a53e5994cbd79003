/* CPU ORACLE (C twin) — test infrastructure, NOT product code.
 *
 * Plain-C restatement of the reference's hot loops, in the reference's own arithmetic
 * order (sequential per row, separate multiply and add: build with
 * -O2 -fno-fast-math -ffp-contract=off).  Citations are path:line in /root/reference.
 *
 *   spmv_csr!            src/sparse_utils.jl:649-669
 *   5-arg mul!(b,A,x,1,1) call sites src/p_sparse_matrix.jl:2088,2101 (SparseMatricesCSR 0.6)
 *   pack / unpack        src/p_vector.jl:595-609   exchange copy src/primitives.jl:1020-1042
 *   dot / norm           src/p_vector.jl:1189-1206 (per-part partial, sum over parts in order)
 *   CG iteration         HPCG/src/ref_cg.jl:40-71, cg_iterator! :76-97, mul_no_lat! hpcg_utils.jl:6-17
 *   generators           src/gallery.jl:12-86 (7-pt), HPCG/src/sparse_matrix.jl:27-80 (27-pt)
 *
 * Used by tests/ as the checker and by bench.py as the cpu_baseline / --impl reference arm
 * (one part per host thread, the analogue of one MPI rank per core).  Parity status: see the
 * header of oracle/pa_oracle.py (pinned against the reference's golden vectors).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int pa_oracle_abi(void) { return 1; }

int pa_oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* y = A*x (y0 == NULL) or y = y0 + A*x with terms added one by one.  0-based arrays. */
void pa_oracle_spmv_csr(int64_t m, const int64_t *rowptr, const int32_t *colval, const double *nzval,
                        const double *x, const double *y0, double *y) {
  for (int64_t row = 0; row < m; ++row) {
    double bi = y0 ? y0[row] : 0.0;
    for (int64_t p = rowptr[row]; p < rowptr[row + 1]; ++p) {
      double aij = nzval[p];
      double xj = x[colval[p]];
      bi += aij * xj;
    }
    y[row] = bi;
  }
}

double pa_oracle_dot(int64_t n, const double *a, const double *b) {
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* ---------------- stencil generators (box partition, unsplit CSR, local column ids) ------------ */

static int64_t ghost_lookup(int64_t gid, int64_t ng, const int64_t *sorted_gid, const int32_t *lid_of_sorted) {
  int64_t lo = 0, hi = ng - 1;
  while (lo <= hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sorted_gid[mid] == gid) return lid_of_sorted[mid];
    if (sorted_gid[mid] < gid) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

/* kind = 7 or 27.  gn = global dims, lo/hi = own box (0-based, hi exclusive).
 * ghosts: ng sorted 0-based gids + their 0-based ghost ids.  Columns: own id (column-major in
 * the box) for owned points, n_own + ghost id otherwise; sorted ascending within the row.
 * Pass rowptr != NULL, colval == NULL for the counting pass (fills rowptr[0..n_own]).
 * bvec (nullable): 27-pt -> 27 - rowlen (HPCG rhs); 7-pt -> alpha * (#missing neighbours) = (A*1)_i. */
int64_t pa_oracle_stencil_csr(int kind, const int64_t *gn, const int64_t *lo, const int64_t *hi, int64_t ng,
                              const int64_t *sorted_gid, const int32_t *lid_of_sorted, int64_t *rowptr,
                              int32_t *colval, double *nzval, double *bvec) {
  const int64_t bx = hi[0] - lo[0], by = hi[1] - lo[1], bz = hi[2] - lo[2];
  const int64_t n_own = bx * by * bz;
  const double alpha = (double)(gn[0] + 1) * (double)(gn[1] + 1) * (double)(gn[2] + 1);
  const double diag = kind == 7 ? alpha * 2 * 3 : 26.0;
  const double off = kind == 7 ? -alpha : -1.0;
  int64_t nnz = 0;
  for (int64_t iz = 0; iz < bz; ++iz)
    for (int64_t iy = 0; iy < by; ++iy)
      for (int64_t ix = 0; ix < bx; ++ix) {
        const int64_t row = ix + bx * (iy + by * iz);
        const int64_t gx = lo[0] + ix, gy = lo[1] + iy, gz = lo[2] + iz;
        int32_t oc[27], gc[27];
        double ov[27], gv[27];
        int no = 0, ngc = 0;
        for (int sz = -1; sz <= 1; ++sz)
          for (int sy = -1; sy <= 1; ++sy)
            for (int sx = -1; sx <= 1; ++sx) {
              if (kind == 7 && (abs(sx) + abs(sy) + abs(sz)) > 1) continue;
              const int64_t cx = gx + sx, cy = gy + sy, cz = gz + sz;
              if (cx < 0 || cx >= gn[0] || cy < 0 || cy >= gn[1] || cz < 0 || cz >= gn[2]) continue;
              const double v = (sx == 0 && sy == 0 && sz == 0) ? diag : off;
              if (cx >= lo[0] && cx < hi[0] && cy >= lo[1] && cy < hi[1] && cz >= lo[2] && cz < hi[2]) {
                oc[no] = (int32_t)((cx - lo[0]) + bx * ((cy - lo[1]) + by * (cz - lo[2])));
                ov[no++] = v;
              } else {
                const int64_t gid = cx + gn[0] * (cy + gn[1] * cz);
                const int64_t g = ghost_lookup(gid, ng, sorted_gid, lid_of_sorted);
                if (g < 0) return -1;
                int k = ngc++;
                while (k > 0 && gc[k - 1] > (int32_t)(n_own + g)) { gc[k] = gc[k - 1]; gv[k] = gv[k - 1]; --k; }
                gc[k] = (int32_t)(n_own + g);
                gv[k] = v;
              }
            }
        if (colval) {
          int64_t p = rowptr[row];
          for (int k = 0; k < no; ++k) { colval[p] = oc[k]; nzval[p++] = ov[k]; }
          for (int k = 0; k < ngc; ++k) { colval[p] = gc[k]; nzval[p++] = gv[k]; }
          if (bvec) bvec[row] = kind == 7 ? alpha * (double)(7 - (no + ngc)) : 27.0 - (double)(no + ngc);
        } else {
          rowptr[row + 1] = no + ngc;
        }
        nnz += no + ngc;
      }
  if (!colval) {
    rowptr[0] = 0;
    for (int64_t r = 0; r < n_own; ++r) rowptr[r + 1] += rowptr[r];
  }
  return nnz;
}

/* ---------------- multi-part CG (one part per OpenMP thread) ---------------------------------- */

typedef struct {
  int64_t n_own, n_local;          /* own entries are local[0:n_own] (block partitions) */
  const int64_t *rowptr;           /* unsplit CSR n_own x n_local, 0-based */
  const int32_t *colval;
  const double *nzval;
  /* consistent! plan = reversed assembly cache (src/p_vector.jl:427-437,748): */
  int32_t n_snd;                   /* neighbours I send own values to   (assembly neighbors_rcv) */
  const int32_t *nbr_snd;          /* 0-based part ids */
  const int32_t *snd_ptrs;         /* 0-based, len n_snd+1 */
  const int32_t *snd_lids;         /* 0-based local ids to pack         (assembly local_indices_rcv) */
  int32_t n_rcv;                   /* neighbours I receive ghosts from  (assembly neighbors_snd) */
  const int32_t *nbr_rcv;
  const int32_t *rcv_ptrs;
  const int32_t *rcv_lids;         /* 0-based local ids to unpack into  (assembly local_indices_snd) */
  double *buf_snd, *buf_rcv;       /* JaggedArray data buffers */
  const double *b;                 /* n_local */
  double *x, *r, *c, *u;           /* n_local each */
} pa_oracle_part;

/* pack values[lid] of vector v into buf_snd (src/p_vector.jl:595-599) */
static void pack_vec(pa_oracle_part *P, const double *v) {
  const int64_t n = P->snd_ptrs[P->n_snd];
  for (int64_t p = 0; p < n; ++p) P->buf_snd[p] = v[P->snd_lids[p]];
}

/* exchange: copy each neighbour's matching send segment into my receive buffer
 * (src/primitives.jl:1020-1042) */
static void exchange_into(pa_oracle_part *parts, int me) {
  pa_oracle_part *R = &parts[me];
  for (int i = 0; i < R->n_rcv; ++i) {
    pa_oracle_part *S = &parts[R->nbr_rcv[i]];
    int j = 0;
    while (S->nbr_snd[j] != me) ++j;
    const int64_t len = R->rcv_ptrs[i + 1] - R->rcv_ptrs[i];
    memcpy(R->buf_rcv + R->rcv_ptrs[i], S->buf_snd + S->snd_ptrs[j], (size_t)len * sizeof(double));
  }
}

/* unpack with insert(a,b)=b (src/p_vector.jl:605-609,755) */
static void unpack_vec(pa_oracle_part *P, double *v) {
  const int64_t n = P->rcv_ptrs[P->n_rcv];
  for (int64_t p = 0; p < n; ++p) v[P->rcv_lids[p]] = P->buf_rcv[p];
}

typedef struct { double t_total, t_spmv, t_dot, t_waxpby, t_exch; } pa_oracle_times;

static double now_s(void) {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  return 0.0;
#endif
}

/* ref_cg! with Pl=Identity (HPCG/src/ref_cg.jl).  hist has maxiter+1 slots (residual0 first).
 * Returns the number of iterations done.  x must hold the initial guess. */
int pa_oracle_cg(int nparts, pa_oracle_part *parts, int maxiter, double tol, double *hist, pa_oracle_times *tm) {
  double *partial = (double *)calloc((size_t)nparts, sizeof(double));
  double rho = 1.0, residual0 = 0.0, residual = 0.0;
  int iters = 0, stop = 0;
  pa_oracle_times T = {0, 0, 0, 0, 0};
#pragma omp parallel num_threads(nparts)
  {
#ifdef _OPENMP
    const int p = omp_get_thread_num();
#else
    const int p = 0;
#endif
    {
      pa_oracle_part *P = &parts[p];
      /* cg_iterator!: u .= 0 ; r = b ; c = A*x ; r .-= c */
      for (int64_t i = 0; i < P->n_local; ++i) { P->u[i] = 0.0; P->r[i] = P->b[i]; }
      pack_vec(P, P->x);
    }
#pragma omp barrier
    exchange_into(parts, p);
#pragma omp barrier
    {
      pa_oracle_part *P = &parts[p];
      unpack_vec(P, P->x);
      pa_oracle_spmv_csr(P->n_own, P->rowptr, P->colval, P->nzval, P->x, NULL, P->c);
      for (int64_t i = 0; i < P->n_local; ++i) P->r[i] -= P->c[i];
      /* norm(own)^2 : BLAS nrm2 then squared in the reference (src/p_vector.jl:1203) */
      double s = 0.0;
      for (int64_t i = 0; i < P->n_own; ++i) s += P->r[i] * P->r[i];
      partial[p] = s;
    }
#pragma omp barrier
#pragma omp single
    {
      double s = 0.0;
      for (int q = 0; q < nparts; ++q) s += partial[q];
      residual0 = residual = sqrt(s);
      hist[0] = residual;
      stop = (maxiter <= 0) || (residual0 == 0.0) || (residual / residual0 <= tol);
    }
    while (!stop) {
      pa_oracle_part *P = &parts[p];
      double t0 = now_s(), t1;
      /* ldiv!(c, Identity, r) ; rho = dot(c,r) */
      memcpy(P->c, P->r, (size_t)P->n_local * sizeof(double));
      partial[p] = pa_oracle_dot(P->n_own, P->c, P->r);
#pragma omp barrier
      double rho_prev = rho, rho_l = 0.0;
      for (int q = 0; q < nparts; ++q) rho_l += partial[q];
#pragma omp barrier
      t1 = now_s();
      if (p == 0) T.t_dot += t1 - t0;
      const double beta = rho_l / rho_prev;
      /* u .= c .+ beta .* u (own and ghost entries, src/p_vector.jl:1271-1276) */
      for (int64_t i = 0; i < P->n_local; ++i) P->u[i] = P->c[i] + beta * P->u[i];
      double t2 = now_s();
      if (p == 0) T.t_waxpby += t2 - t1;
      /* mul_no_lat!: consistent!(u)|>wait ; spmv! */
      pack_vec(P, P->u);
#pragma omp barrier
      exchange_into(parts, p);
      unpack_vec(P, P->u);
      double t3 = now_s();
      if (p == 0) T.t_exch += t3 - t2;
      pa_oracle_spmv_csr(P->n_own, P->rowptr, P->colval, P->nzval, P->u, NULL, P->c);
      double t4 = now_s();
      if (p == 0) T.t_spmv += t4 - t3;
      partial[p] = pa_oracle_dot(P->n_own, P->u, P->c);
#pragma omp barrier
      double uc = 0.0;
      for (int q = 0; q < nparts; ++q) uc += partial[q];
#pragma omp barrier
      double t5 = now_s();
      if (p == 0) T.t_dot += t5 - t4;
      const double alpha = rho_l / uc;
      for (int64_t i = 0; i < P->n_local; ++i) P->x[i] += alpha * P->u[i];
      for (int64_t i = 0; i < P->n_local; ++i) P->r[i] -= alpha * P->c[i];
      double t6 = now_s();
      if (p == 0) T.t_waxpby += t6 - t5;
      double s = 0.0;
      for (int64_t i = 0; i < P->n_own; ++i) s += P->r[i] * P->r[i];
      partial[p] = s;
#pragma omp barrier
      double nr = 0.0;
      for (int q = 0; q < nparts; ++q) nr += partial[q];
      nr = sqrt(nr);
#pragma omp barrier
      double t7 = now_s();
      if (p == 0) { T.t_dot += t7 - t6; T.t_total += t7 - t0; }
#pragma omp single
      {
        rho = rho_l;
        residual = nr;
        iters += 1;
        hist[iters] = residual;
        stop = (iters >= maxiter) || (residual / residual0 <= tol);
      }
      /* implicit barrier after single */
    }
  }
  if (tm) *tm = T;
  free(partial);
  return iters;
}

/* Timed SpMV loop over all parts (consistent! + spmv!), `reps` times; returns seconds. */
double pa_oracle_time_spmv(int nparts, pa_oracle_part *parts, int reps) {
  double t = 0.0;
#pragma omp parallel num_threads(nparts)
  {
#ifdef _OPENMP
    const int p = omp_get_thread_num();
#else
    const int p = 0;
#endif
    pa_oracle_part *P = &parts[p];
#pragma omp barrier
    double t0 = now_s();
    for (int k = 0; k < reps; ++k) {
      pack_vec(P, P->u);
#pragma omp barrier
      exchange_into(parts, p);
      unpack_vec(P, P->u);
      pa_oracle_spmv_csr(P->n_own, P->rowptr, P->colval, P->nzval, P->u, NULL, P->c);
#pragma omp barrier
    }
    if (p == 0) t = now_s() - t0;
  }
  return t;
}

/* ---------------- Gauss-Seidel sweeps of the HPCG multigrid smoother ---------------------------
 * PartitionedSolvers gauss_seidel_sweep! / gauss_seidel_sweep_zero! for SparseMatrixCSR
 * (PartitionedSolvers/src/smoothers.jl:162-176, 248-269): per part, rows in own order (forward 1:n,
 * backward n:-1:1), unsplit CSR n_own x n_local, ghost entries of x fixed during the sweep.
 *   full:  s = b[row]; for p in row: s -= a*x[col]; s += d*x[row]; s = s/d; x[row] = s
 *   zero:  s = b[row]; for p in row with col < row: s -= a*x[col];       s = s/d; x[row] = s   */
void pa_oracle_gs_sweep(int64_t n, const int64_t *rowptr, const int32_t *colval, const double *nzval,
                        const double *diag, const double *b, double *x, int backward, int zero_guess) {
  for (int64_t k = 0; k < n; ++k) {
    const int64_t row = backward ? n - 1 - k : k;
    double s = b[row];
    for (int64_t p = rowptr[row]; p < rowptr[row + 1]; ++p) {
      const int32_t col = colval[p];
      if (zero_guess && col >= row) continue;
      s -= nzval[p] * x[col];
    }
    const double d = diag[row];
    if (!zero_guess) s += d * x[row];
    s = s / d;
    x[row] = s;
  }
}

/* Gauss-Seidel sweep in an arbitrary row order (multi-colour smoother: perm = rows sorted by colour): the same per-row
 * arithmetic as pa_oracle_gs_sweep, rows visited as perm[0..n-1] (backward: perm[n-1..0]).  The zero-guess variant uses only
 * the entries whose column was already updated in this sweep (x starts at zero: the other terms vanish) and skips
 * s += d*x[row], like the lexicographic zero-guess sweep (smoothers.jl:248-269), of which it is the generalisation.
 * done: n_own bytes of scratch. */
void pa_oracle_gs_sweep_perm(int64_t n, const int64_t *rowptr, const int32_t *colval, const double *nzval, const double *diag,
                             const double *b, double *x, const int32_t *perm, unsigned char *done, int backward, int zero_guess) {
  for (int64_t k = 0; k < n; ++k) done[k] = 0;
  for (int64_t k = 0; k < n; ++k) {
    const int64_t row = perm[backward ? n - 1 - k : k];
    double s = b[row];
    for (int64_t p = rowptr[row]; p < rowptr[row + 1]; ++p) {
      const int32_t col = colval[p];
      if (zero_guess && !(col < n && done[col])) continue;
      s -= nzval[p] * x[col];
    }
    const double d = diag[row];
    if (!zero_guess) s += d * x[row];
    s = s / d;
    x[row] = s;
    done[row] = 1;
  }
}
