// On-device COO -> CSR compression and value refresh.
// Reference: sparse_matrix / compresscoo (src/sparse_utils.jl:313-350,392-405: entries with i<1 or j<1 become a stored
// (1,1,0); columns sorted within rows; duplicates combined with + in input order), precompute_nzindex (:434-455) and
// sparse_matrix!(A,V,K) (:457-469: fillstored!(A,0); A_nz[K[p]] += V[p] in input order) as used by psparse / psparse!
// (src/p_sparse_matrix.jl:1196-1203,1291-1305).
//
// Device version: 64-bit keys (row<<32 | col), stable radix sort carrying the input position, segment heads -> unique
// entries, one thread per unique entry adds its duplicates in input order (stable sort keeps it) starting from 0 — the
// same floating-point sequence as the reference.  The sorted permutation and the segment table are kept so that a
// later refresh with new values (same pattern) is a single gather-sum kernel: the reference's K cache.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "pa_internal.h"

__global__ void k_coo_keys(const int64_t *I, const int64_t *J, int64_t n, int64_t nrows, int64_t ncols, uint64_t *key, int32_t *pos,
                           unsigned char *valid, int *err) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = I[p], j = J[p];
    const bool ok = i >= 1 && j >= 1;
    const bool oob = ok && (i > nrows || j > ncols);
    if (oob) *err = 4;  // reported as PA_EINVAL; the entry is parked on key 0 so that no later kernel indexes out of bounds
    key[p] = (ok && !oob) ? (((uint64_t)(i - 1)) << 32) | (uint64_t)(j - 1) : 0ull;  // skipped entries -> stored (1,1,0)
    pos[p] = (int32_t)p;
    valid[p] = ok ? 1 : 0;
  }
}

__global__ void k_coo_heads(const uint64_t *key, int64_t n, int32_t *head) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x)
    head[s] = (s == 0 || key[s] != key[s - 1]) ? 1 : 0;
}

// upos = exclusive scan of head shifted: unique index of sorted entry s is incl[s]-1
__global__ void k_coo_unique(const uint64_t *key, const int32_t *head, const int32_t *incl, int64_t n, int32_t *ucol, int32_t *segstart,
                             int64_t *rowcount) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (int64_t)gridDim.x * blockDim.x)
    if (head[s]) {
      const int32_t u = incl[s] - 1;
      ucol[u] = (int32_t)(key[s] & 0xffffffffull);
      segstart[u] = (int32_t)s;
      atomicAdd((unsigned long long *)(rowcount + (key[s] >> 32)), 1ull);
    }
}

__global__ void k_coo_sum(const double *V, const int32_t *perm, const unsigned char *valid, const int32_t *segstart, int64_t nuniq,
                          int64_t n, double *nz) {
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < nuniq; u += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = segstart[u], b = u + 1 < nuniq ? segstart[u + 1] : n;
    double acc = 0.0;
    for (int64_t s = a; s < b; ++s) {
      const int32_t p = perm[s];
      if (!valid || valid[p]) acc = __dadd_rn(acc, V[p]);  // A_nz[k] += v, input order
    }
    nz[u] = acc;
  }
}

__global__ void k_narrow64(const int64_t *in, int32_t *out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (int32_t)in[i];
}

static void free_coo_cache(MatPart &m) {
  cudaFree(m.d_coo_perm);
  cudaFree(m.d_coo_seg);
  cudaFree(m.d_coo_valid);
  m.d_coo_perm = m.d_coo_seg = nullptr;
  m.d_coo_valid = nullptr;
  m.n_coo = 0;
}

/* Setup-time helper of the assembly stages that stay on the host (per-part compression of disassembled triplets, COO -> CSR of
 * the blocks): perm = the STABLE ascending order of 64-bit keys (row and column packed into one key = the order of
 * sortperm / the reference's compresscoo).  Device radix sort; keys and permutation travel over PCIe (12 bytes per entry)
 * instead of a host merge sort of the same array (22 M entries: seconds). */
__global__ void k_iota32(int32_t *p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (int32_t)i;
}
extern "C" int pa_sort_perm_u64(pa_ctx *c, const uint64_t *keys, int64_t n, int32_t *perm) {
  PA_CHECK(c && (n == 0 || (keys && perm)) && n >= 0 && n < (1ll << 31), PA_EINVAL, "pa_sort_perm_u64: bad arguments");
  if (n == 0) return PA_OK;
  PA_CUDA(cudaSetDevice(c->device));
  unsigned long long *k1 = nullptr, *k2 = nullptr;
  int32_t *p1 = nullptr, *p2 = nullptr;
  void *tmp = nullptr;
  size_t tb = 0;
  cudaError_t e = cudaMalloc((void **)&k1, n * 8);
  if (e == cudaSuccess) e = cudaMalloc((void **)&k2, n * 8);
  if (e == cudaSuccess) e = cudaMalloc((void **)&p1, n * 4);
  if (e == cudaSuccess) e = cudaMalloc((void **)&p2, n * 4);
  if (e == cudaSuccess) {
    cub::DeviceRadixSort::SortPairs(nullptr, tb, k1, k2, p1, p2, (int)n, 0, 64, c->stream);
    e = cudaMalloc(&tmp, tb ? tb : 1);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(k1, keys, n * 8, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    k_iota32<<<148 * 4, 256, 0, c->stream>>>(p1, n);
    e = cub::DeviceRadixSort::SortPairs(tmp, tb, k1, k2, p1, p2, (int)n, 0, 64, c->stream);  // stable
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(perm, p2, n * 4, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(k1); cudaFree(k2); cudaFree(p1); cudaFree(p2); cudaFree(tmp);
  c->launches += 2;
  if (e != cudaSuccess) {
    cudaGetLastError();
    pa_set_error("pa_sort_perm_u64: %s", cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? PA_ENOMEM : PA_ECUDA;
  }
  return PA_OK;
}

/* sparse_matrix(T, I, J, V, m, n; reuse=true) on the device for local part k.
 * I: 1-based OWN row ids, J: 1-based LOCAL column ids (ids < 1 are skipped like the reference), idx_bits 32 or 64.
 * The pattern cache (the reference's K) is kept for pa_mat_update_coo_values. */
extern "C" int pa_mat_set_coo(pa_mat *A, int32_t k, int64_t n, int32_t idx_bits, const void *I, const void *J, const double *V) {
  PA_CHECK(A && !A->committed && k >= 0 && k < A->ctx->nlocal, PA_ESTATE, "pa_mat_set_coo: bad matrix/part");
  PA_CHECK((idx_bits == 32 || idx_bits == 64) && n >= 0 && n < (1ll << 31) && (n == 0 || (I && J && V)), PA_EINVAL, "pa_mat_set_coo: bad arguments");
  pa_ctx *c = A->ctx;
  const PlanPart &rp = A->rows->parts[k], &cp = A->cols->parts[k];
  PA_CHECK(cp.prefix, PA_EINVAL, "pa_mat_set_coo: needs the own-first column layout (use the host path otherwise)");
  PA_CUDA(cudaSetDevice(c->device));
  MatPart &m = A->parts[k];
  cudaFree(m.d_rowptr); cudaFree(m.d_colval); cudaFree(m.d_nzval);
  m.d_rowptr = nullptr; m.d_colval = nullptr; m.d_nzval = nullptr;
  free_coo_cache(m);
  const int64_t nrows = rp.n_own, ncols = cp.n_local;
  if (nrows == 0 || ncols == 0) n = 0;  // m*n == 0: nothing is stored (src/sparse_utils.jl:326-329)
  std::vector<int64_t> hI(n), hJ(n);
  for (int64_t p = 0; p < n; ++p) {
    hI[p] = idx_bits == 64 ? ((const int64_t *)I)[p] : ((const int32_t *)I)[p];
    hJ[p] = idx_bits == 64 ? ((const int64_t *)J)[p] : ((const int32_t *)J)[p];
  }
  const int64_t nn = std::max<int64_t>(n, 1);
  int64_t *dI = nullptr, *dJ = nullptr, *d_rowcount = nullptr, *d_rp64 = nullptr;
  double *dV = nullptr;
  uint64_t *key = nullptr, *key2 = nullptr;
  int32_t *pos = nullptr, *head = nullptr, *incl = nullptr;
  PA_CUDA(cudaMalloc((void **)&dI, nn * 8)); PA_CUDA(cudaMalloc((void **)&dJ, nn * 8)); PA_CUDA(cudaMalloc((void **)&dV, nn * 8));
  PA_CUDA(cudaMalloc((void **)&key, nn * 8)); PA_CUDA(cudaMalloc((void **)&key2, nn * 8));
  PA_CUDA(cudaMalloc((void **)&pos, nn * 4)); PA_CUDA(cudaMalloc((void **)&m.d_coo_perm, nn * 4));
  PA_CUDA(cudaMalloc((void **)&head, nn * 4)); PA_CUDA(cudaMalloc((void **)&incl, nn * 4));
  PA_CUDA(cudaMalloc((void **)&m.d_coo_valid, nn));
  PA_CUDA(cudaMalloc((void **)&d_rowcount, (nrows + 1) * 8)); PA_CUDA(cudaMalloc((void **)&d_rp64, (nrows + 1) * 8));
  PA_CUDA(cudaMemsetAsync(d_rowcount, 0, (nrows + 1) * 8, c->stream));
  int64_t nuniq = 0;
  if (n) {
    PA_CUDA(cudaMemcpyAsync(dI, hI.data(), n * 8, cudaMemcpyHostToDevice, c->stream));
    PA_CUDA(cudaMemcpyAsync(dJ, hJ.data(), n * 8, cudaMemcpyHostToDevice, c->stream));
    PA_CUDA(cudaMemcpyAsync(dV, V, n * 8, cudaMemcpyHostToDevice, c->stream));
    const int g = 148 * 8;
    k_coo_keys<<<g, 256, 0, c->stream>>>(dI, dJ, n, nrows, ncols, key, pos, m.d_coo_valid, c->d_err);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, key, key2, pos, m.d_coo_perm, (int)n, 0, 64, c->stream);
    void *tmp = nullptr;
    PA_CUDA(cudaMalloc(&tmp, tb ? tb : 1));
    PA_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, key, key2, pos, m.d_coo_perm, (int)n, 0, 64, c->stream));  // stable
    k_coo_heads<<<g, 256, 0, c->stream>>>(key2, n, head);
    size_t tb2 = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb2, head, incl, (int)n, c->stream);
    void *tmp2 = nullptr;
    PA_CUDA(cudaMalloc(&tmp2, tb2 ? tb2 : 1));
    PA_CUDA(cub::DeviceScan::InclusiveSum(tmp2, tb2, head, incl, (int)n, c->stream));
    int32_t last = 0;
    PA_CUDA(cudaMemcpyAsync(&last, incl + n - 1, 4, cudaMemcpyDeviceToHost, c->stream));
    PA_CUDA(cudaStreamSynchronize(c->stream));
    nuniq = last;
    PA_CUDA(cudaMalloc((void **)&m.d_colval, (nuniq + PA_MAT_PAD) * 4));
    PA_CUDA(cudaMalloc((void **)&m.d_nzval, (nuniq + PA_MAT_PAD) * 8));
    PA_CUDA(cudaMalloc((void **)&m.d_coo_seg, (nuniq + 1) * 4));
    k_coo_unique<<<g, 256, 0, c->stream>>>(key2, head, incl, n, m.d_colval, m.d_coo_seg, d_rowcount);
    k_coo_sum<<<g, 256, 0, c->stream>>>(dV, m.d_coo_perm, m.d_coo_valid, m.d_coo_seg, nuniq, n, m.d_nzval);
    cudaFree(tmp);  // (stream-ordered frees would be nicer; these are setup-time)
    cudaFree(tmp2);
    c->launches += 6;
  } else {
    PA_CUDA(cudaMalloc((void **)&m.d_colval, 16 * 4));
    PA_CUDA(cudaMalloc((void **)&m.d_nzval, 16 * 8));
  }
  // rowptr = exclusive scan of the per-row unique counts
  size_t tb3 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb3, d_rowcount, d_rp64, (int)(nrows + 1), c->stream);
  void *tmp3 = nullptr;
  PA_CUDA(cudaMalloc(&tmp3, tb3 ? tb3 : 1));
  PA_CUDA(cub::DeviceScan::ExclusiveSum(tmp3, tb3, d_rowcount, d_rp64, (int)(nrows + 1), c->stream));
  PA_CUDA(cudaMalloc(&m.d_rowptr, (nrows + 1) * sizeof(int32_t)));
  k_narrow64<<<148, 256, 0, c->stream>>>(d_rp64, (int32_t *)m.d_rowptr, nrows + 1);
  PA_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(tmp3);
  cudaFree(dI); cudaFree(dJ); cudaFree(dV); cudaFree(key); cudaFree(key2); cudaFree(pos); cudaFree(head); cudaFree(incl);
  cudaFree(d_rowcount); cudaFree(d_rp64);
  int herr = 0;
  PA_CUDA(cudaMemcpy(&herr, c->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (herr == 4) {
    cudaMemset(c->d_err, 0, sizeof(int));
    pa_set_error("pa_mat_set_coo: a row or column id exceeds the local matrix size");
    return PA_EINVAL;
  }
  m.nrows = nrows;
  m.ncols = ncols;
  m.nnz = nuniq;
  m.ptr64 = false;
  m.n_coo = n;
  m.rows_per_cta = nuniq <= 8 * nrows ? 256 : (nuniq <= 16 * nrows ? 128 : (nuniq <= 32 * nrows ? 64 : 32));
  m.tile_nnz.clear();
  m.ghost_scanned = false;
  m.set = true;
  return PA_OK;
}

/* sparse_matrix!(A, V, K): refresh the values of a matrix built by pa_mat_set_coo with a new V of the same COO pattern
 * (psparse!, src/p_sparse_matrix.jl:1291-1305).  May be called on a committed matrix. */
extern "C" int pa_mat_update_coo_values(pa_mat *A, int32_t k, const double *V, int64_t n) {
  PA_CHECK(A && k >= 0 && k < A->ctx->nlocal && A->parts[k].set, PA_ESTATE, "pa_mat_update_coo_values: bad matrix/part");
  MatPart &m = A->parts[k];
  PA_CHECK(m.d_coo_perm && n == m.n_coo && (n == 0 || V), PA_EINVAL, "pa_mat_update_coo_values: matrix was not built from COO or length differs");
  pa_ctx *c = A->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  if (!n) return PA_OK;
  double *dV = nullptr;
  PA_CUDA(cudaMalloc((void **)&dV, n * 8));
  PA_CUDA(cudaMemcpyAsync(dV, V, n * 8, cudaMemcpyHostToDevice, c->stream));
  k_coo_sum<<<148 * 8, 256, 0, c->stream>>>(dV, m.d_coo_perm, m.d_coo_valid, m.d_coo_seg, m.nnz, n, m.d_nzval);
  c->launches++;
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(dV);
  pa_mat_drop_transpose(A);  // the cached transposes copy the values: rebuild them on the next transpose product
  return PA_OK;
}

// ------------------------------------------------------------------ transpose product
// mul!(c, transpose(A), b, alpha, beta) (src/p_sparse_matrix.jl:2144-2162): the ghost entries of c receive
// alpha * A_oh^T b_own, assemble!(c) ships them to their owners, and the own entries get
// beta*c_own + alpha * A_oo^T b_own before the received contributions are added.
// Device: the local matrix is transposed once (stable radix sort by column, so every transposed row lists its entries
// by ascending original row — the summation order of a transposed-CSC/CSR product), the product is then ONE ordinary
// streaming SpMV of A_local^T (rows = own|ghost columns of A) and the ghost -> owner step is pa_vec_assemble.
__global__ void k_row_of_entry(const int32_t *rowptr32, const int64_t *rowptr64, int64_t nrows, int32_t *rowidx, int32_t *pos) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = rowptr32 ? (int64_t)rowptr32[r] : rowptr64[r], b = rowptr32 ? (int64_t)rowptr32[r + 1] : rowptr64[r + 1];
    for (int64_t p = a; p < b; ++p) {
      rowidx[p] = (int32_t)r;
      pos[p] = (int32_t)p;
    }
  }
}
__global__ void k_count_cols(const int32_t *colval, int64_t nnz, int64_t *count) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x)
    atomicAdd((unsigned long long *)(count + colval[p]), 1ull);
}
__global__ void k_gather_transposed(const int32_t *perm, const int32_t *rowidx, const double *nzval, int64_t nnz, int32_t *tcol, double *tval) {
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nnz; s += (int64_t)gridDim.x * blockDim.x) {
    const int32_t p = perm[s];
    tcol[s] = rowidx[p];
    tval[s] = nzval[p];
  }
}

static int build_transpose(pa_mat *A) {
  if (A->T) return PA_OK;
  pa_ctx *c = A->ctx;
  pa_mat *T = new pa_mat();
  T->ctx = c;
  T->rows = A->cols;  // rows of A^T = local columns of A (own | ghost)
  T->cols = A->rows;
  T->parts.resize(c->nlocal);
  for (int k = 0; k < c->nlocal; ++k) {
    const MatPart &m = A->parts[k];
    MatPart &t = T->parts[k];
    const int64_t tn = A->cols->parts[k].n_local, nnz = m.nnz;
    PA_CHECK(nnz < (1ll << 31), PA_EINVAL, "transpose product: local matrix too large (nnz >= 2^31)");
    PA_CHECK(A->cols->parts[k].prefix && A->rows->parts[k].prefix, PA_EINVAL, "transpose product needs own-first layouts");
    t.nrows = tn;
    t.ncols = m.nrows;
    t.nnz = nnz;
    t.ptr64 = false;
    int64_t *d_count = nullptr, *d_tp64 = nullptr;
    PA_CUDA(cudaMalloc((void **)&d_count, (tn + 1) * 8));
    PA_CUDA(cudaMalloc((void **)&d_tp64, (tn + 1) * 8));
    PA_CUDA(cudaMemsetAsync(d_count, 0, (tn + 1) * 8, c->stream));
    PA_CUDA(cudaMalloc((void **)&t.d_colval, (nnz + PA_MAT_PAD) * 4));
    PA_CUDA(cudaMalloc((void **)&t.d_nzval, (nnz + PA_MAT_PAD) * 8));
    PA_CUDA(cudaMalloc(&t.d_rowptr, (tn + 1) * 4));
    const int g = 148 * 8;
    if (nnz) {
      int32_t *rowidx = nullptr, *pos = nullptr, *perm = nullptr, *keys2 = nullptr;
      PA_CUDA(cudaMalloc((void **)&rowidx, nnz * 4)); PA_CUDA(cudaMalloc((void **)&pos, nnz * 4));
      PA_CUDA(cudaMalloc((void **)&perm, nnz * 4)); PA_CUDA(cudaMalloc((void **)&keys2, nnz * 4));
      k_row_of_entry<<<g, 256, 0, c->stream>>>(m.ptr64 ? nullptr : (const int32_t *)m.d_rowptr, m.ptr64 ? (const int64_t *)m.d_rowptr : nullptr,
                                                m.nrows, rowidx, pos);
      k_count_cols<<<g, 256, 0, c->stream>>>(m.d_colval, nnz, d_count);
      int bits = 1;
      while ((1ll << bits) < tn) ++bits;
      size_t tb = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tb, m.d_colval, keys2, pos, perm, (int)nnz, 0, bits, c->stream);
      void *tmp = nullptr;
      PA_CUDA(cudaMalloc(&tmp, tb ? tb : 1));
      PA_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, m.d_colval, keys2, pos, perm, (int)nnz, 0, bits, c->stream));  // stable
      k_gather_transposed<<<g, 256, 0, c->stream>>>(perm, rowidx, m.d_nzval, nnz, t.d_colval, t.d_nzval);
      PA_CUDA(cudaStreamSynchronize(c->stream));
      cudaFree(tmp); cudaFree(rowidx); cudaFree(pos); cudaFree(perm); cudaFree(keys2);
      c->launches += 4;
    }
    size_t tb3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb3, d_count, d_tp64, (int)(tn + 1), c->stream);
    void *tmp3 = nullptr;
    PA_CUDA(cudaMalloc(&tmp3, tb3 ? tb3 : 1));
    PA_CUDA(cub::DeviceScan::ExclusiveSum(tmp3, tb3, d_count, d_tp64, (int)(tn + 1), c->stream));
    k_narrow64<<<148, 256, 0, c->stream>>>(d_tp64, (int32_t *)t.d_rowptr, tn + 1);
    PA_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(tmp3); cudaFree(d_count); cudaFree(d_tp64);
    t.rows_per_cta = nnz <= 8 * tn ? 256 : (nnz <= 16 * tn ? 128 : (nnz <= 32 * tn ? 64 : 32));
    t.set = true;
  }
  T->committed = true;
  A->T = T;
  return PA_OK;
}

__global__ void k_zero_tail(double *v, int64_t from, int64_t to) {
  for (int64_t i = from + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < to; i += (int64_t)gridDim.x * blockDim.x) v[i] = 0.0;
}

/* mul!(c, transpose(A), b, alpha, beta): b on the row partition of A (own entries are read), c on the column partition
 * of A (own entries updated, ghost entries are zero on return, like after assemble!). */
extern "C" int pa_spmv_transpose(pa_mat *A, pa_vec *b, pa_vec *cvec, double alpha, double beta) {
  PA_CHECK(A && b && cvec && A->committed, PA_ESTATE, "pa_spmv_transpose: matrix missing or not committed");
  pa_ctx *c = A->ctx;
  PA_CHECK(b->plan->ctx == c && cvec->plan->ctx == c && b->offset != cvec->offset, PA_EINVAL, "pa_spmv_transpose: bad operands");
  for (int k = 0; k < c->nlocal; ++k) {
    const PlanPart &rp = A->rows->parts[k], &cp = A->cols->parts[k], &bp = b->plan->parts[k], &yp = cvec->plan->parts[k];
    PA_CHECK(bp.n_own == rp.n_own && bp.prefix, PA_EINVAL, "pa_spmv_transpose: b does not match axes(A,1) on part %d", c->part_ids[k] + 1);
    PA_CHECK(yp.n_own == cp.n_own && yp.n_local == cp.n_local && yp.prefix, PA_EINVAL,
             "pa_spmv_transpose: c does not match axes(A,2) (own and ghost ids) on part %d", c->part_ids[k] + 1);
  }
  PA_CHECK(!A->subassembled, PA_EINVAL, "pa_spmv_transpose: needs an assembled matrix");
  PA_CUDA(cudaSetDevice(c->device));
  PA_TRY(build_transpose(A));
  PA_TRY(pa_before_write(c));
  for (int k = 0; k < c->nlocal; ++k) {  // fill!(ghost_values(c), 0)
    const PlanPart &yp = cvec->plan->parts[k];
    if (!yp.n_ghost) continue;
    k_zero_tail<<<(unsigned)std::min<int64_t>((yp.n_ghost + 255) / 256, 1184), 256, 0, c->stream>>>(cvec->d[k], yp.n_own, yp.n_local);
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  // c_local = beta*c_local + alpha * A_local^T b_own (ghost rows start from zero)
  PA_TRY(pa_spmv_local(A->T, b, cvec, alpha, beta, 0, nullptr, nullptr));
  return pa_vec_assemble(cvec);
}


// ------------------------------------------------------------------ local sparse x sparse product (spmm / spmtm / rap)
// D_k = A_k * C_k on the device for every local part (src/p_sparse_matrix.jl:2237-2262: `spmm(partition(A), partition(C))` after
// C = consistent(B, axes(A,2)) brought the rows of B that A's ghost columns refer to).  A_k: rows of part k x local columns of
// axes(A,2); C_k: one row per LOCAL column of A (own rows of B first, then the fetched ghost rows) x local columns of C.
// Expand - sort - compress: every product a_ik * c_kj becomes a triplet (i, j, a*c), generated row by row of A with k in the
// stored (ascending local id) order; a stable radix sort by (i, j) keeps that order inside every output entry, and one thread
// per output entry adds its products in that order — the order of a Gustavson product over the local matrices (and of
// SparseArrays' CSC product: ascending k), with one rounding per product and per addition, no FMA.
__device__ __forceinline__ int64_t rp_at(const void *rp, int is64, int64_t i) {
  return is64 ? reinterpret_cast<const int64_t *>(rp)[i] : (int64_t)reinterpret_cast<const int32_t *>(rp)[i];
}
__global__ void k_spgemm_count(const int32_t *colA, int64_t nnzA, const void *rpC, int c64, int64_t *cnt) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p <= nnzA; p += (int64_t)gridDim.x * blockDim.x)
    cnt[p] = p < nnzA ? rp_at(rpC, c64, colA[p] + 1) - rp_at(rpC, c64, colA[p]) : 0;
}
__global__ void k_spgemm_expand(const int32_t *rowidxA, const int32_t *colA, const double *nzA, int64_t nnzA, const int64_t *off, const void *rpC,
                                int c64, const int32_t *colC, const double *nzC, uint64_t *key, int32_t *pos, double *val) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnzA; p += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t hi = ((uint64_t)(uint32_t)rowidxA[p]) << 32;
    const double a = nzA[p];
    int64_t o = off[p];
    for (int64_t q = rp_at(rpC, c64, colA[p]); q < rp_at(rpC, c64, colA[p] + 1); ++q, ++o) {
      key[o] = hi | (uint64_t)(uint32_t)colC[q];
      val[o] = __dmul_rn(a, nzC[q]);
      pos[o] = (int32_t)o;
    }
  }
}

/* D = A * C, part by part (see above).  D: created on (axes(A,1), axes(C,2)), not committed; C must hold one row per local
 * column of A.  Entries of D are sorted by column inside every row; explicit zeros are kept (like the reference's product). */
extern "C" int pa_mat_spmm_local(pa_mat *A, pa_mat *C, pa_mat *D) {
  PA_CHECK(A && C && D && A->committed && C->committed && !D->committed, PA_ESTATE, "pa_mat_spmm_local: A and C must be committed, D must not");
  pa_ctx *c = A->ctx;
  PA_CHECK(C->ctx == c && D->ctx == c, PA_EINVAL, "pa_mat_spmm_local: matrices live on different backends");
  PA_CUDA(cudaSetDevice(c->device));
  const int g = 148 * 8;
  for (int k = 0; k < c->nlocal; ++k) {
    const MatPart &a = A->parts[k], &cm = C->parts[k];
    MatPart &d = D->parts[k];
    PA_CHECK(cm.nrows == a.ncols, PA_EINVAL, "pa_mat_spmm_local: part %d: C has %lld rows, A has %lld local columns", c->part_ids[k] + 1,
             (long long)cm.nrows, (long long)a.ncols);
    PA_CHECK(D->rows->parts[k].n_own == a.nrows || (D->rows->parts[k].prefix && D->rows->parts[k].n_local == a.nrows), PA_EINVAL,
             "pa_mat_spmm_local: D's row partition does not match A's rows");
    PA_CHECK(D->cols->parts[k].n_local == cm.ncols && D->cols->parts[k].prefix, PA_EINVAL, "pa_mat_spmm_local: D's column partition does not match C's");
    PA_CHECK(a.nnz < (1ll << 31), PA_EINVAL, "pa_mat_spmm_local: A too large");
    cudaFree(d.d_rowptr); cudaFree(d.d_colval); cudaFree(d.d_nzval);
    d = MatPart();
    d.nrows = a.nrows;
    d.ncols = cm.ncols;
    if (a.nrows != D->rows->parts[k].n_own) D->subassembled = true;
    int64_t *d_cnt = nullptr, *d_off = nullptr, *d_rowcount = nullptr, *d_rp64 = nullptr;
    int32_t *rowidx = nullptr, *epos = nullptr;
    const int64_t na = std::max<int64_t>(a.nnz, 1);
    PA_CUDA(cudaMalloc((void **)&d_cnt, (na + 1) * 8));
    PA_CUDA(cudaMalloc((void **)&d_off, (na + 1) * 8));
    PA_CUDA(cudaMalloc((void **)&rowidx, na * 4));
    PA_CUDA(cudaMalloc((void **)&epos, na * 4));
    PA_CUDA(cudaMalloc((void **)&d_rowcount, (a.nrows + 1) * 8));
    PA_CUDA(cudaMalloc((void **)&d_rp64, (a.nrows + 1) * 8));
    PA_CUDA(cudaMemsetAsync(d_rowcount, 0, (a.nrows + 1) * 8, c->stream));
    int64_t T = 0, nuniq = 0;
    if (a.nnz) {
      k_row_of_entry<<<g, 256, 0, c->stream>>>(a.ptr64 ? nullptr : (const int32_t *)a.d_rowptr, a.ptr64 ? (const int64_t *)a.d_rowptr : nullptr, a.nrows, rowidx, epos);
      k_spgemm_count<<<g, 256, 0, c->stream>>>(a.d_colval, a.nnz, cm.d_rowptr, cm.ptr64 ? 1 : 0, d_cnt);
      size_t tb = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tb, d_cnt, d_off, (int)(a.nnz + 1), c->stream);
      void *tmp = nullptr;
      PA_CUDA(cudaMalloc(&tmp, tb ? tb : 1));
      PA_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, d_cnt, d_off, (int)(a.nnz + 1), c->stream));
      PA_CUDA(cudaMemcpyAsync(&T, d_off + a.nnz, 8, cudaMemcpyDeviceToHost, c->stream));
      PA_CUDA(cudaStreamSynchronize(c->stream));
      cudaFree(tmp);
      c->launches += 3;
    }
    PA_CHECK(T < (1ll << 31), PA_EINVAL, "pa_mat_spmm_local: %lld products in part %d: too many for one pass", (long long)T, c->part_ids[k] + 1);
    if (T) {
      uint64_t *key = nullptr, *key2 = nullptr;
      int32_t *pos = nullptr, *perm = nullptr, *head = nullptr, *incl = nullptr, *seg = nullptr;
      double *val = nullptr;
      PA_CUDA(cudaMalloc((void **)&key, T * 8)); PA_CUDA(cudaMalloc((void **)&key2, T * 8)); PA_CUDA(cudaMalloc((void **)&val, T * 8));
      PA_CUDA(cudaMalloc((void **)&pos, T * 4)); PA_CUDA(cudaMalloc((void **)&perm, T * 4));
      PA_CUDA(cudaMalloc((void **)&head, T * 4)); PA_CUDA(cudaMalloc((void **)&incl, T * 4));
      k_spgemm_expand<<<g, 256, 0, c->stream>>>(rowidx, a.d_colval, a.d_nzval, a.nnz, d_off, cm.d_rowptr, cm.ptr64 ? 1 : 0, cm.d_colval, cm.d_nzval, key, pos, val);
      size_t tb = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tb, key, key2, pos, perm, (int)T, 0, 64, c->stream);
      void *tmp = nullptr;
      PA_CUDA(cudaMalloc(&tmp, tb ? tb : 1));
      PA_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, key, key2, pos, perm, (int)T, 0, 64, c->stream));  // stable
      k_coo_heads<<<g, 256, 0, c->stream>>>(key2, T, head);
      size_t tb2 = 0;
      cub::DeviceScan::InclusiveSum(nullptr, tb2, head, incl, (int)T, c->stream);
      void *tmp2 = nullptr;
      PA_CUDA(cudaMalloc(&tmp2, tb2 ? tb2 : 1));
      PA_CUDA(cub::DeviceScan::InclusiveSum(tmp2, tb2, head, incl, (int)T, c->stream));
      int32_t last = 0;
      PA_CUDA(cudaMemcpyAsync(&last, incl + T - 1, 4, cudaMemcpyDeviceToHost, c->stream));
      PA_CUDA(cudaStreamSynchronize(c->stream));
      nuniq = last;
      PA_CUDA(cudaMalloc((void **)&d.d_colval, (nuniq + PA_MAT_PAD) * 4));
      PA_CUDA(cudaMalloc((void **)&d.d_nzval, (nuniq + PA_MAT_PAD) * 8));
      PA_CUDA(cudaMalloc((void **)&seg, (nuniq + 1) * 4));
      k_coo_unique<<<g, 256, 0, c->stream>>>(key2, head, incl, T, d.d_colval, seg, d_rowcount);
      k_coo_sum<<<g, 256, 0, c->stream>>>(val, perm, nullptr, seg, nuniq, T, d.d_nzval);
      PA_CUDA(cudaStreamSynchronize(c->stream));
      cudaFree(tmp); cudaFree(tmp2); cudaFree(key); cudaFree(key2); cudaFree(val); cudaFree(pos); cudaFree(perm); cudaFree(head); cudaFree(incl); cudaFree(seg);
      c->launches += 6;
    } else {
      PA_CUDA(cudaMalloc((void **)&d.d_colval, 16 * 4));
      PA_CUDA(cudaMalloc((void **)&d.d_nzval, 16 * 8));
    }
    size_t tb3 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb3, d_rowcount, d_rp64, (int)(a.nrows + 1), c->stream);
    void *tmp3 = nullptr;
    PA_CUDA(cudaMalloc(&tmp3, tb3 ? tb3 : 1));
    PA_CUDA(cub::DeviceScan::ExclusiveSum(tmp3, tb3, d_rowcount, d_rp64, (int)(a.nrows + 1), c->stream));
    PA_CUDA(cudaMalloc(&d.d_rowptr, (a.nrows + 1) * sizeof(int32_t)));
    k_narrow64<<<148, 256, 0, c->stream>>>(d_rp64, (int32_t *)d.d_rowptr, a.nrows + 1);
    PA_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(tmp3); cudaFree(d_cnt); cudaFree(d_off); cudaFree(rowidx); cudaFree(epos); cudaFree(d_rowcount); cudaFree(d_rp64);
    d.nnz = nuniq;
    d.ptr64 = false;
    d.rows_per_cta = nuniq <= 8 * d.nrows ? 256 : (nuniq <= 16 * d.nrows ? 128 : (nuniq <= 32 * d.nrows ? 64 : 32));
    d.set = true;
  }
  return PA_OK;
}

/* The local transposes of an assembled matrix as a matrix of their own: T_k = (A_k)^T, one row per LOCAL column of A (own, then
 * ghost) x own rows of A — the left factor of spmtm (transpose(A)*B, src/p_sparse_matrix.jl:2276-2290).  T: created on
 * (axes(A,2), axes(A,1)), not committed; its ghost rows make it a sub-assembled matrix. */
extern "C" int pa_mat_transpose_local(pa_mat *A, pa_mat *T) {
  PA_CHECK(A && T && A->committed && !T->committed && !A->subassembled, PA_ESTATE, "pa_mat_transpose_local: A must be committed and assembled, T not committed");
  pa_ctx *c = A->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  PA_TRY(build_transpose(A));
  for (int k = 0; k < c->nlocal; ++k) {
    const MatPart &t = A->T->parts[k];
    MatPart &o = T->parts[k];
    PA_CHECK(T->rows->parts[k].prefix && T->rows->parts[k].n_local == t.nrows && T->cols->parts[k].n_own == t.ncols, PA_EINVAL,
             "pa_mat_transpose_local: T must live on (axes(A,2), axes(A,1))");
    cudaFree(o.d_rowptr); cudaFree(o.d_colval); cudaFree(o.d_nzval);
    o = MatPart();
    o.nrows = t.nrows; o.ncols = T->cols->parts[k].n_local; o.nnz = t.nnz; o.ptr64 = false; o.rows_per_cta = t.rows_per_cta;
    PA_CUDA(cudaMalloc(&o.d_rowptr, (t.nrows + 1) * 4));
    PA_CUDA(cudaMalloc((void **)&o.d_colval, (t.nnz + PA_MAT_PAD) * 4));
    PA_CUDA(cudaMalloc((void **)&o.d_nzval, (t.nnz + PA_MAT_PAD) * 8));
    PA_CUDA(cudaMemcpyAsync(o.d_rowptr, t.d_rowptr, (t.nrows + 1) * 4, cudaMemcpyDeviceToDevice, c->stream));
    if (t.nnz) {
      PA_CUDA(cudaMemcpyAsync(o.d_colval, t.d_colval, t.nnz * 4, cudaMemcpyDeviceToDevice, c->stream));
      PA_CUDA(cudaMemcpyAsync(o.d_nzval, t.d_nzval, t.nnz * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    PA_CUDA(cudaStreamSynchronize(c->stream));
    if (t.nrows != T->rows->parts[k].n_own) T->subassembled = true;
    o.set = true;
  }
  return PA_OK;
}
