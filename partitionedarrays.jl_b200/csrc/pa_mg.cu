// HPCG multigrid preconditioner: symmetric Gauss-Seidel smoother + injection restrict/prolong + V-cycle.
// Reference: PartitionedSolvers/src/smoothers.jl:82-125 (gauss_seidel step), :162-176 (CSR sweep),
// :248-269 (zero-guess sweep); HPCG/src/mg_preconditioner.jl:81-101 (f2c), :202-206 (ldiv!),
// :224-251 (restrict!/prolongate!), :314-328 (pc_solve!).
//
// The reference sweeps the own rows of each part sequentially (1:n, then n:-1:1).  A sequential sweep is a
// dependency DAG: row i needs the NEW values of its lower-numbered neighbours and the OLD values of the others.
// We execute exactly that DAG as a dataflow: one warp per row, rows handed out in wavefront (level) order, each
// lane that needs a NEW value spins on that row's published (value, sweep epoch) pair.  Every row sees
// precisely the inputs of the sequential sweep and performs the same arithmetic in the same order
// (s -= a*x[col] in CSR order, s += d*x[row], s /= d; separate multiply/add), so the result is bit-identical to
// the reference's sweep; no multi-colouring (which would change the iteration) and no per-level barrier.
// Deadlock freedom: rows are processed in level order by a persistent grid whose warps are all resident; a warp
// only ever waits for rows of lower levels, which sit earlier in the order of some resident warp.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <array>

#include "pa_internal.h"

#define GS_THREADS 256
#define GS_SPIN_LIMIT (20000000000LL)  // ~10 s of SM cycles, then give up and flag an error instead of hanging

struct GsPart {
  int64_t n = 0;
  int nlev = 0;
  int32_t *d_rows = nullptr;  // own rows sorted by wavefront level (ascending); every level starts at a multiple
                              // of 4 (-1 = padding) so that the rows a warp takes together never depend on each other
  int64_t npad = 0;           // length of d_rows (multiple of 4)
  int maxlen = 0;             // longest stored row
  ulonglong2 *d_xe = nullptr; // [n] value of the row's last update + the sweep epoch it happened in (gs_publish)
  int epoch = 0;
  bool geom = false;
  int64_t dims[3] = {0, 0, 0}, w[3] = {0, 0, 0};
};

struct pa_gs {
  pa_mat *A = nullptr;
  std::vector<GsPart> parts;
  bool committed = false;
};

struct pa_mg {
  pa_ctx *ctx = nullptr;
  int nlev = 0;
  std::vector<pa_mat *> A;   // [0] coarsest ... [nlev-1] finest (the reference's 1-based levels minus one)
  std::vector<pa_gs *> gs;
  std::vector<pa_vec *> r, x, Axf;
  std::vector<std::vector<std::array<int64_t, 3>>> dims;  // [level][local part] local box dims
};

static int gs_default_lanes(pa_ctx *c, const GsPart &p);

template <typename PtrT>
struct GsArgs {
  const PtrT *rowptr;
  const int32_t *colval;
  const double *nzval;
  const double *b;
  double *x;
  const int32_t *rows;
  ulonglong2 *xe;
  int *err;
  int64_t n, npad;
  int epoch, backward, zero_guess;
};

// Publishing a row: the new value travels WITH its "done in this sweep" mark, so a waiting reader needs one
// round trip and the writer needs no fence between value and flag.  The 64-bit value is split over two 8-byte
// words, each carrying the sweep epoch in its upper half ({lo32, epoch}, {hi32, epoch}); an aligned 8-byte
// store is single-copy atomic, so a reader that sees the current epoch in BOTH words has both halves of the
// new value, however the 16-byte access is split on the way.
__device__ __forceinline__ double gs_wait_value(const ulonglong2 *slot, unsigned epoch, int *err) {
  unsigned long long w0, w1;
  long long t0 = 0;
  for (;;) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(slot) : "memory");
    if ((unsigned)(w0 >> 32) == epoch && (unsigned)(w1 >> 32) == epoch) break;
    if (!t0) {
      t0 = clock64();
    } else if (clock64() - t0 > GS_SPIN_LIMIT) {
      *err = 3;
      break;
    }
  }
  return __longlong_as_double((long long)((w1 << 32) | (w0 & 0xffffffffull)));
}
__device__ __forceinline__ void gs_publish(ulonglong2 *slot, double *x, double s, unsigned epoch) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(s), e = (unsigned long long)epoch << 32;
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(slot), "l"((bits & 0xffffffffull) | e), "l"((bits >> 32) | e) : "memory");
  *x = s;  // the plain vector: read by later rows as an OLD value never, by the next kernels always
}

// The matrix is read once per sweep and is two orders of magnitude larger than x: marked evict-first in L2 so
// that the L2 keeps x and the published values, which neighbouring rows (up to two grid planes apart in the
// wavefront order) come back for.  Without the hint the 512^3 sweep misses L2 on most x gathers.
__device__ __forceinline__ uint64_t gs_stream_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double gs_ld_stream(const double *p, uint64_t pol) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int32_t gs_ld_stream(const int32_t *p, uint64_t pol) {
  int32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}

// any row length: one warp per row, 32 entries per trip
template <typename PtrT>
__global__ void __launch_bounds__(GS_THREADS) k_gs_flow(const GsArgs<PtrT> a) {
  __shared__ double prod[GS_THREADS / 32][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t W = (int64_t)gridDim.x * (GS_THREADS / 32);
  const int64_t w = (int64_t)blockIdx.x * (GS_THREADS / 32) + warp;
  const uint64_t pol = gs_stream_policy();
  for (int64_t pos = w; pos < a.npad; pos += W) {
    const int64_t row = a.backward ? a.rows[a.npad - 1 - pos] : a.rows[pos];
    if (row < 0) continue;
    const int64_t ps = (int64_t)a.rowptr[row], pe = (int64_t)a.rowptr[row + 1];
    double s = 0.0, d = 0.0, xold = 0.0;
    if (lane == 0) {
      s = a.b[row];
      xold = __ldcg(a.x + row);
    }
    for (int64_t p0 = ps; p0 < pe; p0 += 32) {
      const int64_t p = p0 + lane;
      const bool valid = p < pe;
      const int32_t col = valid ? gs_ld_stream(a.colval + p, pol) : -1;
      const double v = valid ? gs_ld_stream(a.nzval + p, pol) : 0.0;
      const bool use = valid && (!a.zero_guess || col < row);
      // NEW value needed: an own row that precedes this one in the sweep order
      const bool fresh = use && col < a.n && (a.backward ? col > row : col < row);
      double xv = 0.0;
      if (fresh) xv = gs_wait_value(a.xe + col, (unsigned)a.epoch, a.err);
      else if (use) xv = __ldcg(a.x + col);  // x changes during the sweep: always through L2
      prod[warp][lane] = __dmul_rn(v, xv);
      const unsigned usemask = __ballot_sync(0xffffffffu, use);
      const unsigned dmask = __ballot_sync(0xffffffffu, valid && col == row);
      if (dmask) d = __shfl_sync(0xffffffffu, v, __ffs(dmask) - 1);
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const double pk = prod[warp][k];
          if ((usemask >> k) & 1u) s = __dsub_rn(s, pk);  // s -= a*x[col], in CSR order
        }
      }
      __syncwarp();
    }
    if (lane == 0) {
      if (!a.zero_guess) s = __dadd_rn(s, __dmul_rn(d, xold));  // s += d*x[row]
      s = __ddiv_rn(s, d);
      gs_publish(a.xe + row, a.x + row, s, (unsigned)a.epoch);
    }
  }
}

// Rows of at most 32 stored entries (every stencil operator): G lanes per row (32/G rows per warp), each lane
// owning 32/G entries, and a three-stage software pipeline over the warp's rows -- the row id of iteration i+2,
// the row extent of iteration i+1 and the entries/b/x_old of iteration i+1 are in flight while iteration i waits
// for its inputs -- so that the per-row critical path is wait -> ordered subtraction -> publish.
// (x_old and the entries can be fetched early: nobody writes x[row] before this row does, and a row's
// coefficients are constant.)  Same arithmetic, same order as k_gs_flow.
template <typename PtrT, int G>
__global__ void __launch_bounds__(GS_THREADS, 6) k_gs_flow_pipe(const GsArgs<PtrT> a) {
  constexpr int NJ = 32 / G;   // entries per lane
  constexpr int RPW = 32 / G;  // rows per warp
  constexpr unsigned GM = G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u);
  // +2: the leaders of a warp's rows read their products with 16-byte loads at the same time; 34 doubles apart
  // they fall into different banks
  __shared__ __align__(16) double prod[GS_THREADS / 32][RPW][34];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / G, hl = lane % G;
  const int64_t W = (int64_t)gridDim.x * (GS_THREADS / 32) * RPW;
  const int64_t slot = ((int64_t)blockIdx.x * (GS_THREADS / 32) + warp) * RPW + sub;
  const int64_t np = a.npad;
  const int64_t niter = (np + W - 1) / W;
  const unsigned epoch = (unsigned)a.epoch;
  const uint64_t pol = gs_stream_policy();
  auto row_at = [&](int64_t pos) -> int32_t { return pos < np ? a.rows[a.backward ? np - 1 - pos : pos] : -1; };
  auto extent = [&](int32_t r, int64_t &ps, int &cnt) {
    ps = 0;
    cnt = 0;
    if (r >= 0) {
      ps = (int64_t)a.rowptr[r];
      cnt = (int)((int64_t)a.rowptr[r + 1] - ps);
    }
  };
  // pipeline registers: A = row id (i+2), B = row id + extent (i+1), C = everything (i)
  int32_t rC = row_at(slot), rB = row_at(slot + W), rA = row_at(slot + 2 * W);
  int64_t psC, psB;
  int cntC, cntB;
  extent(rC, psC, cntC);
  extent(rB, psB, cntB);
  int32_t colC[NJ];
  double vC[NJ], bC = 0.0, xoC = 0.0;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int k = hl + j * G;
    colC[j] = k < cntC ? gs_ld_stream(a.colval + psC + k, pol) : -1;
    vC[j] = k < cntC ? gs_ld_stream(a.nzval + psC + k, pol) : 0.0;
  }
  if (hl == 0 && rC >= 0) {
    bC = a.b[rC];
    xoC = __ldcg(a.x + rC);
  }
  for (int64_t it = 0; it < niter; ++it) {
    // ---- issue the loads of the next stages
    const int32_t rA2 = row_at(slot + (it + 3) * W);
    int64_t psA;
    int cntA;
    extent(rA, psA, cntA);
    int32_t colB[NJ];
    double vB[NJ], bB = 0.0, xoB = 0.0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int k = hl + j * G;
      colB[j] = k < cntB ? gs_ld_stream(a.colval + psB + k, pol) : -1;
      vB[j] = k < cntB ? gs_ld_stream(a.nzval + psB + k, pol) : 0.0;
    }
    if (hl == 0 && rB >= 0) {
      bB = a.b[rB];
      xoB = __ldcg(a.x + rB);
    }
    // ---- the row of this iteration
    const int32_t row = rC;
    bool use[NJ], fresh[NJ];
    double xv[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int32_t col = colC[j];
      use[j] = col >= 0 && (!a.zero_guess || col < row);
      // NEW value needed: an own row that precedes this one in the sweep order
      fresh[j] = use[j] && col < a.n && (a.backward ? col > row : col < row);
      xv[j] = (use[j] && !fresh[j]) ? __ldcg(a.x + col) : 0.0;  // OLD values and ghosts: through L2
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j)
      if (fresh[j]) xv[j] = gs_wait_value(a.xe + colC[j], epoch, a.err);
    double d = 0.0;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      // unused entries contribute +0.0: s - (+0.0) == s bit for bit, so the leader's chain needs no predicates
      prod[warp][sub][hl + j * G] = use[j] ? __dmul_rn(vC[j], xv[j]) : 0.0;
      const unsigned db = (__ballot_sync(0xffffffffu, colC[j] >= 0 && colC[j] == row) >> (sub * G)) & GM;
      const double dj = __shfl_sync(0xffffffffu, vC[j], db ? sub * G + __ffs(db) - 1 : lane);
      if (db) d = dj;
    }
    __syncwarp();
    if (hl == 0 && row >= 0) {
      double s = bC;
      const double2 *pp = reinterpret_cast<const double2 *>(&prod[warp][sub][0]);
#pragma unroll
      for (int k = 0; k < 14; ++k) {  // s -= a*x[col], in CSR order
        const double2 pk = pp[k];
        s = __dsub_rn(__dsub_rn(s, pk.x), pk.y);
      }
      if (cntC > 28) {
        const double2 p14 = pp[14], p15 = pp[15];
        s = __dsub_rn(__dsub_rn(__dsub_rn(__dsub_rn(s, p14.x), p14.y), p15.x), p15.y);
      }
      if (!a.zero_guess) s = __dadd_rn(s, __dmul_rn(d, xoC));  // s += d*x[row]
      s = __ddiv_rn(s, d);
      gs_publish(a.xe + row, a.x + row, s, epoch);
    }
    __syncwarp();
    // ---- rotate the pipeline
    rC = rB; psC = psB; cntC = cntB; bC = bB; xoC = xoB;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      colC[j] = colB[j];
      vC[j] = vB[j];
    }
    rB = rA; psB = psA; cntB = cntA;
    rA = rA2;
  }
}

__global__ void k_level_hist(const int32_t *lev, int64_t n, int *count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) atomicAdd(count + lev[i], 1);
}
__global__ void k_pad_levels(const int32_t *lev_sorted, const int32_t *rows_sorted, const int32_t *shift, int64_t n, int32_t *padded) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    padded[i + shift[lev_sorted[i]]] = rows_sorted[i];
}
template <typename PtrT>
__global__ void k_max_rowlen(const PtrT *rowptr, int64_t n, int *out) {
  int m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, (int)(rowptr[i + 1] - rowptr[i]));
  for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

__global__ void k_levels_box(int32_t *lev, int32_t *rows, int64_t n, int64_t bx, int64_t by, int64_t wx, int64_t wy, int64_t wz) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ix = i % bx, iy = (i / bx) % by, iz = i / (bx * by);
    lev[i] = (int32_t)(wx * ix + wy * iy + wz * iz);
    rows[i] = (int32_t)i;
  }
}
__global__ void k_iota(int32_t *rows, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) rows[i] = (int32_t)i;
}

static void gs_free_part(GsPart &g) {
  cudaFree(g.d_rows);
  cudaFree(g.d_xe);
  g = GsPart();
}

extern "C" int pa_gs_create(pa_mat *A, pa_gs **out) {
  PA_CHECK(A && out && A->committed, PA_ESTATE, "pa_gs_create: matrix missing or not committed");
  pa_gs *g = new pa_gs();
  g->A = A;
  g->parts.resize(A->ctx->nlocal);
  *out = g;
  return PA_OK;
}

/* Geometry hint for stencil operators on a box (local dims, x fastest): wavefront weights w such that every
 * lower-numbered neighbour has a strictly smaller w.(ix,iy,iz): 27-pt (1,2,4), 7-pt (1,1,1). */
extern "C" int pa_gs_set_box(pa_gs *g, int32_t k, int32_t kind, const int64_t *dims) {
  PA_CHECK(g && dims && !g->committed && k >= 0 && k < (int)g->parts.size(), PA_EINVAL, "pa_gs_set_box: bad arguments");
  PA_CHECK(kind == 7 || kind == 27, PA_EINVAL, "pa_gs_set_box: kind must be 7 or 27");
  GsPart &p = g->parts[k];
  PA_CHECK(dims[0] * dims[1] * dims[2] == g->A->parts[k].nrows, PA_EINVAL, "pa_gs_set_box: dims do not match the own rows");
  p.geom = true;
  for (int d = 0; d < 3; ++d) p.dims[d] = dims[d];
  p.w[0] = 1;
  p.w[1] = kind == 27 ? 2 : 1;
  p.w[2] = kind == 27 ? 4 : 1;
  return PA_OK;
}

extern "C" int pa_gs_commit(pa_gs *g) {
  PA_CHECK(g && !g->committed, PA_ESTATE, "pa_gs_commit: missing or already committed");
  pa_ctx *c = g->A->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  for (int k = 0; k < c->nlocal; ++k) {
    GsPart &p = g->parts[k];
    const MatPart &m = g->A->parts[k];
    p.n = m.nrows;
    if (p.n == 0) continue;
    PA_CHECK(p.n < (1ll << 31), PA_EINVAL, "pa_gs_commit: too many rows");
    int32_t *d_lev = nullptr, *d_lev2 = nullptr, *d_rows0 = nullptr, *d_rows1 = nullptr;
    PA_CUDA(cudaMalloc((void **)&d_lev, p.n * sizeof(int32_t)));
    PA_CUDA(cudaMalloc((void **)&d_lev2, p.n * sizeof(int32_t)));
    PA_CUDA(cudaMalloc((void **)&d_rows0, p.n * sizeof(int32_t)));
    PA_CUDA(cudaMalloc((void **)&d_rows1, p.n * sizeof(int32_t)));
    if (p.geom) {
      k_levels_box<<<148 * 8, 256, 0, c->stream>>>(d_lev, d_rows0, p.n, p.dims[0], p.dims[1], p.w[0], p.w[1], p.w[2]);
      p.nlev = (int)(p.w[0] * (p.dims[0] - 1) + p.w[1] * (p.dims[1] - 1) + p.w[2] * (p.dims[2] - 1) + 1);
    } else {
      // generic matrices: level[i] = 1 + max level[j] over own neighbours j < i (host pass over the CSR)
      std::vector<int64_t> rp(p.n + 1);
      std::vector<int32_t> cv(m.nnz);
      PA_TRY(pa_mat_download_csr(g->A, k, rp.data(), cv.data(), nullptr));
      std::vector<int32_t> lev(p.n, 0);
      int mx = 0;
      for (int64_t i = 0; i < p.n; ++i) {
        int l = 0;
        for (int64_t q = rp[i]; q < rp[i + 1]; ++q)
          if (cv[q] < i) l = std::max(l, lev[cv[q]] + 1);
        lev[i] = l;
        mx = std::max(mx, l);
      }
      for (int64_t i = 0; i < p.n; ++i)  // the backward sweep walks the same order in reverse: needs a symmetric pattern
        for (int64_t q = rp[i]; q < rp[i + 1]; ++q)
          PA_CHECK(!(cv[q] > i && cv[q] < p.n && lev[cv[q]] <= lev[i]), PA_EINVAL,
                   "pa_gs_commit: non-symmetric sparsity pattern (row %lld): the wavefront order needs a symmetric pattern", (long long)i);
      p.nlev = mx + 1;
      PA_CUDA(cudaMemcpyAsync(d_lev, lev.data(), p.n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
      k_iota<<<148 * 8, 256, 0, c->stream>>>(d_rows0, p.n);
      PA_CUDA(cudaStreamSynchronize(c->stream));
    }
    size_t tmp_bytes = 0;
    int bits = 1;
    while ((1ll << bits) < p.nlev) ++bits;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_lev, d_lev2, d_rows0, d_rows1, (int)p.n, 0, bits, c->stream);
    void *d_tmp = nullptr;
    PA_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
    PA_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_lev, d_lev2, d_rows0, d_rows1, (int)p.n, 0, bits, c->stream));
    // start every level at a multiple of 4 in the order (see GsPart::d_rows)
    int *d_cnt = nullptr;
    PA_CUDA(cudaMalloc((void **)&d_cnt, (p.nlev + 1) * sizeof(int)));
    PA_CUDA(cudaMemsetAsync(d_cnt, 0, (p.nlev + 1) * sizeof(int), c->stream));
    k_level_hist<<<148 * 8, 256, 0, c->stream>>>(d_lev, p.n, d_cnt);
    if (m.ptr64) k_max_rowlen<int64_t><<<148 * 8, 256, 0, c->stream>>>((const int64_t *)m.d_rowptr, p.n, d_cnt + p.nlev);
    else k_max_rowlen<int32_t><<<148 * 8, 256, 0, c->stream>>>((const int32_t *)m.d_rowptr, p.n, d_cnt + p.nlev);
    std::vector<int> cnt(p.nlev + 1);
    PA_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, (p.nlev + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PA_CUDA(cudaStreamSynchronize(c->stream));
    p.maxlen = cnt[p.nlev];
    std::vector<int32_t> shift(p.nlev);
    int64_t at = 0, padded = 0;
    const int64_t al = 4;
    for (int l = 0; l < p.nlev; ++l) {
      padded = (padded + al - 1) & ~(al - 1);
      PA_CHECK(padded - at < (1ll << 31) && padded + cnt[l] < (1ll << 31), PA_EINVAL, "pa_gs_commit: too many rows");
      shift[l] = (int32_t)(padded - at);
      at += cnt[l];
      padded += cnt[l];
    }
    PA_CHECK(at == p.n, PA_ESTATE, "pa_gs_commit: level histogram does not add up");
    p.npad = (padded + al - 1) & ~(al - 1);
    int32_t *d_shift = nullptr;
    PA_CUDA(cudaMalloc((void **)&d_shift, p.nlev * sizeof(int32_t)));
    PA_CUDA(cudaMemcpyAsync(d_shift, shift.data(), p.nlev * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    PA_CUDA(cudaMalloc((void **)&p.d_rows, p.npad * sizeof(int32_t)));
    PA_CUDA(cudaMemsetAsync(p.d_rows, 0xff, p.npad * sizeof(int32_t), c->stream));
    k_pad_levels<<<148 * 8, 256, 0, c->stream>>>(d_lev2, d_rows1, d_shift, p.n, p.d_rows);
    PA_CUDA(cudaMalloc((void **)&p.d_xe, p.n * sizeof(ulonglong2)));
    PA_CUDA(cudaMemsetAsync(p.d_xe, 0, p.n * sizeof(ulonglong2), c->stream));
    PA_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_tmp); cudaFree(d_lev); cudaFree(d_lev2); cudaFree(d_rows0); cudaFree(d_rows1); cudaFree(d_cnt); cudaFree(d_shift);
    p.epoch = 0;
    c->launches += 5;
  }
  g->committed = true;
  return PA_OK;
}

extern "C" int pa_gs_destroy(pa_gs *g) {
  if (!g) return PA_OK;
  pa_ctx *c = g->A->ctx;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto &p : g->parts) gs_free_part(p);
  delete g;
  return PA_OK;
}

// lanes per row: 16 (two rows per warp) once the levels are wide enough to be throughput bound, 32 where the
// sweep is bound by the level-to-level hop; 0 = the unpipelined warp-per-row kernel (any row length).
// Measured on B200 (27-pt, symmetric sweep): 16.8M rows 11.7 ms (16) vs 13.5 ms (32), 2.1M rows 5.0 vs 3.7 ms.
// (One THREAD per row was tried for the widest levels and is 2.7x slower: 205 ms vs 75 ms at 134M rows.)
static int gs_default_lanes(pa_ctx *c, const GsPart &p) {
  return (int)pa_knob(c, "gs_lanes", p.n >= (4ll << 20) ? 16 : 32);
}

// one sweep over the own rows of every local part (ghost entries of x are inputs only)
static int gs_sweep(pa_gs *g, pa_vec *x, const pa_vec *b, int backward, int zero_guess) {
  pa_ctx *c = g->A->ctx;
  for (int k = 0; k < c->nlocal; ++k) {
    GsPart &p = g->parts[k];
    const MatPart &m = g->A->parts[k];
    if (p.n == 0) continue;
    p.epoch += 1;
    auto launch = [&](auto tag) -> int {
      using PtrT = decltype(tag);
      GsArgs<PtrT> a;
      a.rowptr = (const PtrT *)m.d_rowptr;
      a.colval = m.d_colval;
      a.nzval = m.d_nzval;
      a.b = b->d[k];
      a.x = x->d[k];
      a.rows = p.d_rows;
      a.xe = p.d_xe;
      a.err = c->d_err;
      a.n = p.n;
      a.npad = p.npad;
      a.epoch = p.epoch;
      a.backward = backward;
      a.zero_guess = zero_guess;
      int lanes = gs_default_lanes(c, p);
      if (p.maxlen > 32) lanes = 0;
      void (*kern)(const GsArgs<PtrT>) = lanes == 8 ? k_gs_flow_pipe<PtrT, 8> : lanes == 16 ? k_gs_flow_pipe<PtrT, 16>
                                         : lanes == 32 ? k_gs_flow_pipe<PtrT, 32> : k_gs_flow<PtrT>;
      const int rpw = lanes == 8 ? 4 : lanes == 16 ? 2 : 1;
      int ctas_per_sm = 0;
      PA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, GS_THREADS, 0));
      if (ctas_per_sm < 1) ctas_per_sm = 1;
      int nsm = 148;
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
      // all CTAs of the grid must be co-resident: a waiting warp only waits for rows held by running warps
      const int64_t per_cta = (int64_t)(GS_THREADS / 32) * rpw;
      const int64_t want = (p.npad + per_cta - 1) / per_cta;
      const int64_t grid = std::min<int64_t>(want, (int64_t)nsm * ctas_per_sm);
      kern<<<(unsigned)grid, GS_THREADS, 0, c->stream>>>(a);
      return PA_OK;
    };
    if (m.ptr64) PA_TRY(launch((int64_t)0)); else PA_TRY(launch((int32_t)0));
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}

/* smooth!(x, gauss_seidel state, b; zero_guess): one symmetric iteration (smoothers.jl:98-125):
 * consistent!(x) unless zero_guess, forward sweep (zero-guess variant when zero_guess), backward sweep. */
extern "C" int pa_gs_smooth(pa_gs *g, pa_vec *x, const pa_vec *b, int32_t zero_guess) {
  PA_CHECK(g && x && b && g->committed, PA_ESTATE, "pa_gs_smooth: smoother missing or not committed");
  pa_ctx *c = g->A->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  for (int k = 0; k < c->nlocal; ++k) {
    const PlanPart &xp = x->plan->parts[k], &cp = g->A->cols->parts[k], &bp = b->plan->parts[k];
    PA_CHECK(xp.n_local == cp.n_local && xp.n_own == cp.n_own && xp.prefix && bp.n_own == xp.n_own && bp.prefix, PA_EINVAL,
             "pa_gs_smooth: x/b do not match the (own-first) column partition of A on part %d", c->part_ids[k] + 1);
  }
  if (!zero_guess) PA_TRY(pa_vec_consistent(x));
  PA_TRY(pa_before_write(c));
  PA_TRY(gs_sweep(g, x, b, 0, zero_guess ? 1 : 0));
  PA_TRY(gs_sweep(g, x, b, 1, 0));
  return PA_OK;
}

// ------------------------------------------------------------------ restrict / prolongate (injection, f2c)
__global__ void k_restrict(double *rc, const double *bf, const double *axf, int64_t nc, int64_t cx, int64_t cy, int64_t fx, int64_t fy) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ix = i % cx, iy = (i / cx) % cy, iz = i / (cx * cy);
    const int64_t f = 2 * ix + fx * (2 * iy + fy * 2 * iz);
    rc[i] = __dsub_rn(bf[f], axf[f]);  // r_c[i] = r_f[v] - Axf[v]
  }
}
__global__ void k_prolong(double *xf, const double *xc, int64_t nc, int64_t cx, int64_t cy, int64_t fx, int64_t fy) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ix = i % cx, iy = (i / cx) % cy, iz = i / (cx * cy);
    const int64_t f = 2 * ix + fx * (2 * iy + fy * 2 * iz);
    xf[f] = __dadd_rn(xf[f], xc[i]);  // x_f[v] += x_c[i]
  }
}

/* Mg_preconditioner (mg_preconditioner.jl:44-63): levels[0] = coarsest ... levels[n-1] = finest.
 * dims: nlevels x nlocal x 3 local box dims (x fastest); each level halves the one above. */
extern "C" int pa_mg_create(int32_t nlevels, pa_mat **A, pa_gs **gs, const int64_t *dims, pa_mg **out) {
  PA_CHECK(nlevels >= 1 && A && gs && dims && out, PA_EINVAL, "pa_mg_create: bad arguments");
  pa_mg *M = new pa_mg();
  M->ctx = A[0]->ctx;
  M->nlev = nlevels;
  const int nl = M->ctx->nlocal;
  for (int l = 0; l < nlevels; ++l) {
    PA_CHECK(A[l] && gs[l] && A[l]->committed && gs[l]->committed && gs[l]->A == A[l] && A[l]->ctx == M->ctx, PA_ESTATE,
             "pa_mg_create: level %d not ready", l);
    M->A.push_back(A[l]);
    M->gs.push_back(gs[l]);
    std::vector<std::array<int64_t, 3>> d(nl);
    for (int k = 0; k < nl; ++k) {
      for (int q = 0; q < 3; ++q) d[k][q] = dims[((size_t)l * nl + k) * 3 + q];
      PA_CHECK(d[k][0] * d[k][1] * d[k][2] == A[l]->parts[k].nrows, PA_EINVAL, "pa_mg_create: dims of level %d do not match", l);
      if (l > 0)
        for (int q = 0; q < 3; ++q)
          PA_CHECK(M->dims[l - 1][k][q] * 2 == d[k][q], PA_EINVAL, "pa_mg_create: level %d is not twice level %d", l, l - 1);
    }
    M->dims.push_back(d);
  }
  M->r.assign(nlevels, nullptr);
  M->x.assign(nlevels, nullptr);
  M->Axf.assign(nlevels, nullptr);
  for (int l = 0; l < nlevels; ++l) {
    if (l < nlevels - 1) {
      PA_TRY(pa_vec_create(A[l]->cols, &M->r[l]));
      PA_TRY(pa_vec_create(A[l]->cols, &M->x[l]));
      PA_TRY(pa_vec_fill(M->r[l], 0.0));
      PA_TRY(pa_vec_fill(M->x[l], 0.0));
    }
    if (l > 0) {
      PA_TRY(pa_vec_create(A[l]->cols, &M->Axf[l]));
      PA_TRY(pa_vec_fill(M->Axf[l], 0.0));
    }
  }
  *out = M;
  return PA_OK;
}

extern "C" int pa_mg_destroy(pa_mg *M) {
  if (!M) return PA_OK;
  for (int l = M->nlev - 1; l >= 0; --l) {  // reverse creation order (symmetric heap discipline)
    pa_vec_destroy(M->Axf[l]);
    pa_vec_destroy(M->x[l]);
    pa_vec_destroy(M->r[l]);
  }
  delete M;
  return PA_OK;
}

static int small_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

// pc_solve!(x, s, b, l; zero_guess) — mg_preconditioner.jl:314-328 (l is 0-based here)
static int pc_solve(pa_mg *M, pa_vec *x, const pa_vec *b, int l, int zero_guess) {
  pa_ctx *c = M->ctx;
  PA_TRY(pa_gs_smooth(M->gs[l], x, b, zero_guess));  // bottom solve / pre-smoother
  if (l == 0) return PA_OK;
  PA_TRY(pa_spmv(M->A[l], x, M->Axf[l], 1.0, 0.0, PA_SPMV_DEFAULT));  // mul_no_lat!(Axf, A, x)
  PA_TRY(pa_before_write(c));
  for (int k = 0; k < c->nlocal; ++k) {
    const auto &dc = M->dims[l - 1][k], &df = M->dims[l][k];
    const int64_t nc = dc[0] * dc[1] * dc[2];
    if (!nc) continue;
    k_restrict<<<small_grid(nc), 256, 0, c->stream>>>(M->r[l - 1]->d[k], b->d[k], M->Axf[l]->d[k], nc, dc[0], dc[1], df[0], df[1]);
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  PA_TRY(pa_vec_fill(M->x[l - 1], 0.0));
  PA_TRY(pc_solve(M, M->x[l - 1], M->r[l - 1], l - 1, 1));
  PA_TRY(pa_before_write(c));
  for (int k = 0; k < c->nlocal; ++k) {
    const auto &dc = M->dims[l - 1][k], &df = M->dims[l][k];
    const int64_t nc = dc[0] * dc[1] * dc[2];
    if (!nc) continue;
    k_prolong<<<small_grid(nc), 256, 0, c->stream>>>(x->d[k], M->x[l - 1]->d[k], nc, dc[0], dc[1], df[0], df[1]);
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  return pa_gs_smooth(M->gs[l], x, b, 0);  // post-smoother
}

/* ldiv!(x, P::Mg_preconditioner, b) — mg_preconditioner.jl:202-206 */
extern "C" int pa_mg_apply(pa_mg *M, pa_vec *x, const pa_vec *b) {
  PA_CHECK(M && x && b, PA_EINVAL, "pa_mg_apply: null argument");
  PA_CUDA(cudaSetDevice(M->ctx->device));
  PA_TRY(pa_vec_fill(x, 0.0));
  return pc_solve(M, x, b, M->nlev - 1, 1);
}
