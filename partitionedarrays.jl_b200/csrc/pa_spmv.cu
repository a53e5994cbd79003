// PSparseMatrix storage, the CSR SpMV kernel (own rows x local columns, ghost columns read from
// the owner's HBM over NVLink inside the kernel), and on-device generators of the benchmark operators.
// Reference: mul! src/p_sparse_matrix.jl:2090-2142, spmv_csr! src/sparse_utils.jl:649-669,
// mul_no_lat! HPCG/src/hpcg_utils.jl:6-17, generators src/gallery.jl:12-86, HPCG/src/sparse_matrix.jl:27-80.
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <map>
#include <vector>

#include "pa_device.cuh"
#include "pa_internal.h"

#define SPMV_BLOCK 256
// table word of an entry whose column is a ghost: ghost ids follow no pattern (they number the part's halo), so the kernel reads
// that one column from colval — the rows on a part face then share a pattern (and the fixed-length path) like all the others
#define PA_PAT_GHOST ((int32_t)0x80000000)

// ------------------------------------------------------------------ the SpMV kernel
// CSR-stream: a CTA owns ROWS consecutive rows.  All 256 threads stream the CTA's contiguous slice of
// nzval/colval with coalesced loads (U = TILE/256 independent 8-byte + 4-byte loads in flight per
// thread), gather x (L1/L2-resident for banded operators; ghost columns -> NVLink peer load from the
// owner's arena), and park the products a_ij*x_j in shared memory.  Then one thread per row adds its
// row's products from shared memory *sequentially in column order*: exactly the reference loop
// `bi += aij*xj` (separate multiply and add, no FMA), so the result is bit-identical to spmv_csr!.
// Row lengths 7 and 27 are odd -> the per-row walk through shared memory is bank-conflict free.
// HBM traffic = the matrix stream once + x + y + rowptr: the kernel is bandwidth bound by design.
template <typename PtrT>
struct SpmvArgs {
  int64_t nrows;
  const PtrT *rowptr;
  const int32_t *colval;
  const double *nzval;
  const double *x;
  double *y;
  const int32_t *rowmap;  // y index of row i (nullptr: i)
  double alpha, beta;
  int64_t n_own_cols;     // fused: columns >= n_own_cols are ghosts
  const int32_t *gslot, *grlid;
  PeerPtrs peers;
  // fused dot epilogue (CG: u.c with c = A*u): sum_i y_i * dotw_i, one partial per CTA
  const double *dotw;
  double *dot_part;
  // fused consistent! (MODE 4): the kernel itself pulls the ghost values from the owners' HBM into x's ghost slots
  double *xw;                                  // x, writable
  const int32_t *cons_lid, *cons_slot, *cons_rlid;  // the consistent! table of x's plan (ghost slot <- owner slot, lid)
  int64_t n_cons;
  unsigned long long *arrive;                  // CTAs whose share of the gather is stored, summed over all launches
  // folded dot epilogue: the last CTA adds the per-CTA partials in a fixed order and pushes the part's value to the peers
  int fold;
  unsigned *dot_ticket;
  RedPush push;
  const unsigned char *tile_ghost;  // MODE 4: tile t holds a row with a ghost column (computed once per matrix and tile size)
  // PAT: the column stream is replaced by one byte per row (see build_patterns)
  const unsigned char *pat;
  const int32_t *ptab;
  int npat, pat_w;
};

// The matrix stream is read exactly once: keep it out of L1 and mark it evict-first in L2 so the
// 126 MB L2 is left to x (each x entry is reused by up to 7/27 rows, two grid planes apart).
__device__ __forceinline__ uint64_t stream_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double ld_stream_f64(const double *p, uint64_t pol) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ int32_t ld_stream_s32(const int32_t *p, uint64_t pol) {
  int32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}

template <typename PtrT, int ROWS, int TILE, bool FUSED>
__global__ void __launch_bounds__(SPMV_BLOCK) k_spmv_stream(const SpmvArgs<PtrT> a) {
  constexpr int U = TILE / SPMV_BLOCK;
  __shared__ double prod[TILE];
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * ROWS;
  const int64_t r1 = min(r0 + (int64_t)ROWS, a.nrows);
  const int64_t row = r0 + tid;
  const bool has_row = tid < ROWS && row < r1;
  int64_t rs = 0, re = 0;
  if (has_row) {
    rs = (int64_t)a.rowptr[row];
    re = (int64_t)a.rowptr[row + 1];
  }
  const int64_t s = (int64_t)a.rowptr[r0], e = (int64_t)a.rowptr[r1];
  double acc = 0.0;
  const uint64_t pol = stream_policy();
  for (int64_t base = s; base < e; base += TILE) {
    double v[U];
    int32_t c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = base + tid + u * SPMV_BLOCK;
      const bool ok = p < e;
      v[u] = ok ? ld_stream_f64(a.nzval + p, pol) : 0.0;
      c[u] = ok ? ld_stream_s32(a.colval + p, pol) : -1;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      double xv = 0.0;
      if (c[u] >= 0) {
        if (FUSED && c[u] >= a.n_own_cols) {
          const int64_t g = c[u] - a.n_own_cols;
          xv = __ldcg(a.peers.p[a.gslot[g]] + a.grlid[g]);
        } else {
          xv = __ldg(a.x + c[u]);
        }
      }
      prod[tid + u * SPMV_BLOCK] = __dmul_rn(v[u], xv);
    }
    __syncthreads();
    if (has_row) {
      const int64_t lo = max(rs, base), hi = min(re, base + (int64_t)TILE);
      for (int64_t p = lo; p < hi; ++p) acc = __dadd_rn(acc, prod[p - base]);
    }
    __syncthreads();
  }
  if (has_row) {
    const int64_t yi = a.rowmap ? (int64_t)a.rowmap[row] : row;
    if (a.alpha == 1.0 && a.beta == 0.0) {
      a.y[yi] = acc;
    } else {
      // mul!(y,A,x,alpha,beta): y = alpha*(A*x) + beta*y
      const double by = a.beta == 0.0 ? 0.0 : __dmul_rn(a.beta, a.y[yi]);
      a.y[yi] = __dadd_rn(__dmul_rn(a.alpha, acc), by);
    }
  }
}

template <typename PtrT, bool FUSED>
static void launch_spmv_t(const SpmvArgs<PtrT> &a, int rows, cudaStream_t st) {
  const int64_t grid = (a.nrows + rows - 1) / rows;
  switch (rows) {
    case 256: k_spmv_stream<PtrT, 256, 2048, FUSED><<<(unsigned)grid, SPMV_BLOCK, 0, st>>>(a); break;
    case 128: k_spmv_stream<PtrT, 128, 2048, FUSED><<<(unsigned)grid, SPMV_BLOCK, 0, st>>>(a); break;
    case 64: k_spmv_stream<PtrT, 64, 2048, FUSED><<<(unsigned)grid, SPMV_BLOCK, 0, st>>>(a); break;
    default: k_spmv_stream<PtrT, 32, 2048, FUSED><<<(unsigned)grid, SPMV_BLOCK, 0, st>>>(a); break;
  }
}

// ------------------------------------------------------------------ the TMA-pipelined SpMV kernel
// Persistent CTAs (a multiple of the 148 SMs), warp specialised:
//  * one producer lane streams each tile's contiguous nzval/colval slice HBM -> shared memory with
//    cp.async.bulk (the TMA engine; SASS UBLKCP) into a ring of STAGES buffers, completion on an mbarrier,
//    L2 evict-first (the matrix is read exactly once);
//  * ROWS consumer threads, one per row: walk the row's entries in shared memory in column order, gather x
//    (consecutive rows of a banded operator hit consecutive x entries -> coalesced; ghost columns are
//    NVLink peer loads from the owner's arena) and accumulate `bi += aij*xj` sequentially with separate
//    multiply and add — the exact arithmetic of the reference's spmv_csr!.
// The LSU/L1 only sees the x gather and the conflict-free shared-memory reads; the matrix stream never
// passes through registers.
struct TmaCfg {
  int rows, cap, stages, batch;
  int64_t ntiles;
  int rpt = 1;  // rows per consumer thread (k_spmv_pat)
  int fl = 0;   // length of the most frequent row pattern when the fixed-length path is compiled for it (7, 27), else 0
  bool use_pat_kernel = false;
};

__device__ __forceinline__ void keep_live8(const double *x) {
  asm volatile("" ::"d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]), "d"(x[4]), "d"(x[5]), "d"(x[6]), "d"(x[7]));
}
__device__ __forceinline__ void keep_live16(const double *x) {
  asm volatile("" ::"d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]), "d"(x[4]), "d"(x[5]), "d"(x[6]), "d"(x[7]), "d"(x[8]), "d"(x[9]),
               "d"(x[10]), "d"(x[11]), "d"(x[12]), "d"(x[13]), "d"(x[14]), "d"(x[15]));
}
template <int N>
__device__ __forceinline__ void keep_live(const double (&x)[N]) {
  if (N == 8) keep_live8(x);
  if (N >= 16) keep_live16(x);
  if (N >= 32) keep_live16(x + 16);
}

// MODE 4: wait until every CTA of this launch has stored its share of the ghost values.  ONE thread per CTA polls the global
// counter (relaxed loads, then one acquire fence; 150 000 threads polling one L2 line would starve the very atomics they wait
// for) and the consumers meet at a named barrier — they reach a tile together anyway.
__device__ __forceinline__ void spmv_wait_gather(const unsigned long long *arrive, unsigned long long target, int tid, int rows) {
  if (tid == 0) {
    unsigned long long got;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(got) : "l"(arrive) : "memory");
      if (got >= target || clock64() - t0 > 20000000000LL) break;  // (~10 s: never hang the GPU)
      __nanosleep(100);
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  asm volatile("bar.sync 2, %0;" ::"r"(rows) : "memory");
}

// minBlocksPerSM is stated explicitly: with maxThreads alone ptxas squeezes the kernel into 32 registers
// (full-occupancy target) by sinking every load next to its use, which serialises the x gathers.
// MODE 0: every column is a local x entry (1 part, or ghosts already refreshed)
// MODE 1: ghost columns are loaded from the owner's arena inside this kernel (inline NVLink loads)
// MODE 2: own block only (A_oo * x_own); the ghost block is added afterwards by k_spmv_ghost_rows
// MODE 4: consistent!(x) fused into this kernel: while the first matrix tiles are in flight (TMA), the producer warp
//         of every CTA pulls its share of the ghost values from the owners' HBM (NVLink peer loads) into x's ghost
//         slots, fences and counts itself in; rows that touch a ghost column wait (once per thread) until all CTAs
//         are counted in and read the slots through L2.  Own-block products never wait: on a banded operator only
//         the first tiles of the persistent grid can meet the gather still in flight.
// fused dot epilogue of the TMA kernels (consumers only: the producer warp has left): fixed-order block reduction, one partial
// per CTA; folded variant: the last CTA adds the partials in a fixed order and posts the part's value to every part of the job
template <typename PtrT>
__device__ __forceinline__ void spmv_dot_epilogue(const SpmvArgs<PtrT> &a, double dsum, const int tid, const int ROWS) {
  if (a.dotw) {  // consumers only (the producer warp has left): fixed-order block reduction, one partial per CTA
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    if ((tid & 31) == 0) red[tid >> 5] = dsum;
    asm volatile("bar.sync 1, %0;" ::"r"(ROWS) : "memory");
    __shared__ bool last_cta;
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < (ROWS >> 5); ++w) t += red[w];
      a.dot_part[blockIdx.x] = t;
      if (a.fold) {
        __threadfence();
        last_cta = atomicInc(a.dot_ticket, gridDim.x - 1) == gridDim.x - 1;  // wraps to 0: self resetting
      }
    }
    if (a.fold) {
      // the last CTA to finish adds all partials (thread t: partials t, t+ROWS, ... then a fixed tree: deterministic) and
      // posts the part's value to every part of the job: no k_sum_parts launch, no all-reduce launch
      asm volatile("bar.sync 1, %0;" ::"r"(ROWS) : "memory");
      if (last_cta) {
        __threadfence();
        double s = 0.0;
        for (int i = tid; i < (int)gridDim.x; i += ROWS) s += __ldcg(a.dot_part + i);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((tid & 31) == 0) red[tid >> 5] = s;
        asm volatile("bar.sync 1, %0;" ::"r"(ROWS) : "memory");
        if (tid < 32) {
          double t = 0.0;
          for (int w = 0; w < (ROWS >> 5); ++w) t += red[w];
          pa_red_post(a.push, t, tid);
        }
      }
    }
  }
}

// PAT (MODE 0 only): rows of a structured operator repeat a handful of column patterns (column - row for every entry:
//      27 box positions of a stencil).  The kernel then streams ONE BYTE per row — the pattern id — instead of four bytes per
//      entry; the patterns live in shared memory (every interior thread of a warp reads the same word: a broadcast).  Rows
//      whose pattern is not in the table (ghost columns, rare boundary classes) carry id 255 and read colval from global
//      memory.  Matrix stream 12 -> 8 bytes per entry + 1 per row; same products, same order, same bits.
template <typename PtrT, int MODE, int BATCH, bool PAT = false>
__global__ void __launch_bounds__(288, (BATCH >= 32 ? 1 : (BATCH >= 16 ? 2 : (MODE == 1 ? 3 : 4)))) k_spmv_tma(const SpmvArgs<PtrT> a, const TmaCfg cfg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int ROWS = cfg.rows, CAP = cfg.cap, S = cfg.stages;
  // layout: val[S][CAP+2] | col[S][CAP+8] (PAT: pattern ids [S][ROWS+16 bytes]) | p0[S] (int64) | full[S] | empty[S] | PAT: table
  double *val_s = reinterpret_cast<double *>(smem_raw);
  int32_t *col_s = reinterpret_cast<int32_t *>(val_s + (size_t)S * (CAP + 2));
  int64_t *p0_s = reinterpret_cast<int64_t *>(col_s + (size_t)S * (PAT ? (ROWS + 16) / 4 : (CAP + 8)));
  uint64_t *full = reinterpret_cast<uint64_t *>(p0_s + S);
  uint64_t *empty = full + S;
  int32_t *ptab_s = reinterpret_cast<int32_t *>(empty + S);
  const int tid = threadIdx.x;
  __shared__ unsigned long long gather_target;
  if (PAT)
    for (int i = tid; i < a.npat * a.pat_w; i += blockDim.x) ptab_s[i] = a.ptab[i];
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, ROWS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (MODE == 4) {
      // arrivals of earlier launches are complete (kernel boundary) and this CTA has not arrived yet, so fewer than
      // gridDim.x arrivals of THIS launch can be in the count: the launch ends at the next multiple of gridDim.x
      unsigned long long a0;
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a0) : "l"(a.arrive) : "memory");
      gather_target = (a0 / gridDim.x + 1ull) * gridDim.x;
    }
  }
  __syncthreads();
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int64_t nloc = first < cfg.ntiles ? (cfg.ntiles - first + stride - 1) / stride : 0;
  if (tid >= ROWS) {
    // ---------------- producer warp: one elected lane drives the TMA ring
    auto issue = [&](int64_t j, uint64_t pol) {
      const int s = (int)(j % S);
      if (j >= S) mbar_wait(empty + s, (uint32_t)(((j / S) - 1) & 1));
      const int64_t t = first + j * stride;
      const int64_t r0 = t * ROWS, r1 = min(r0 + (int64_t)ROWS, a.nrows);
      const int64_t p0 = (int64_t)a.rowptr[r0], p1 = (int64_t)a.rowptr[r1];
      p0_s[s] = p0;
      const int64_t pv = p0 & ~(int64_t)1, pc = p0 & ~(int64_t)3;
      const uint32_t bv = (uint32_t)(((p1 - pv + 1) & ~(int64_t)1) * 8), bc = (uint32_t)(((p1 - pc + 3) & ~(int64_t)3) * 4);
      const bool any = p1 > p0;
      if (PAT) {
        const uint32_t bp = (uint32_t)(((r1 - r0) + 15) & ~(int64_t)15);  // pattern ids of the tile's rows (r0 is a multiple of 32)
        mbar_expect_tx(full + s, (any ? bv : 0u) + bp);
        if (any) tma_load_1d(val_s + (size_t)s * (CAP + 2), a.nzval + pv, bv, full + s, pol);
        tma_load_1d(reinterpret_cast<unsigned char *>(col_s) + (size_t)s * (ROWS + 16), a.pat + r0, bp, full + s, pol);
      } else {
        mbar_expect_tx(full + s, any ? bv + bc : 0u);
        if (any) {
          tma_load_1d(val_s + (size_t)s * (CAP + 2), a.nzval + pv, bv, full + s, pol);
          tma_load_1d(col_s + (size_t)s * (CAP + 8), a.colval + pc, bc, full + s, pol);
        }
      }
    };
    const uint64_t pol = stream_policy();
    int64_t j0 = 0;
    if (MODE == 4) {
      // fill the ring first (asynchronous), then the whole warp does this CTA's share of consistent!(x)
      const int64_t nfill = min((int64_t)S, nloc);
      if (tid == ROWS)
        for (; j0 < nfill; ++j0) issue(j0, pol);
      __syncwarp();
      const int lane = tid - ROWS;
      const int64_t gstride = (int64_t)gridDim.x * 32;
      for (int64_t q = (int64_t)blockIdx.x * 32 + lane; q < a.n_cons; q += 4 * gstride) {  // 4 NVLink loads in flight per lane
        double g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t qu = q + u * gstride;
          g[u] = qu < a.n_cons ? __ldcg(a.peers.p[a.cons_slot[qu]] + a.cons_rlid[qu]) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t qu = q + u * gstride;
          if (qu < a.n_cons) a.xw[a.cons_lid[qu]] = g[u];
        }
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(a.arrive, 1ull);
      j0 = nfill;  // (only lane 0 advanced its copy)
    }
    if (tid == ROWS)
      for (int64_t j = j0; j < nloc; ++j) issue(j, pol);
    return;
  }
  // ---------------- consumers: one thread per row
  double dsum = 0.0;
  bool ghosts_ready = false;  // MODE 4: this thread has seen the gather complete
  int64_t rs_n = 0, re_n = 0;
  if (nloc > 0) {
    const int64_t row = first * ROWS + tid;
    if (row < a.nrows) {
      rs_n = (int64_t)a.rowptr[row];
      re_n = (int64_t)a.rowptr[row + 1];
    }
  }
  for (int64_t j = 0; j < nloc; ++j) {
    const int s = (int)(j % S);
    const int64_t t = first + j * stride;
    const int64_t row = t * ROWS + tid;
    const int64_t rs = rs_n, re = re_n;
    if (j + 1 < nloc) {  // prefetch the next tile's row pointers while this tile is processed
      const int64_t rown = (t + stride) * ROWS + tid;
      rs_n = re_n = 0;
      if (rown < a.nrows) {
        rs_n = (int64_t)a.rowptr[rown];
        re_n = (int64_t)a.rowptr[rown + 1];
      }
    }
    // MODE 4: only tiles that hold a ghost column wait for the gather (once per thread), and only they read x through the
    // coherent path; every other tile runs the instruction stream of MODE 0
    const bool gt = MODE == 4 && a.tile_ghost[t] != 0;
    if (MODE == 4 && gt && !ghosts_ready) {
      spmv_wait_gather(a.arrive, gather_target, tid, ROWS);
      ghosts_ready = true;
    }
    mbar_wait(full + s, (uint32_t)((j / S) & 1));
    if (row < a.nrows) {
      const int64_t p0 = p0_s[s];
      const double *vs = val_s + (size_t)s * (CAP + 2) + (rs - (p0 & ~(int64_t)1));
      const int32_t *cs = col_s + (size_t)s * (CAP + 8) + (rs - (p0 & ~(int64_t)3));
      const int len = (int)(re - rs);
      double acc = 0.0;
      // PAT: pattern id of this row; 255 = columns from global memory
      const unsigned pid = PAT ? (unsigned)reinterpret_cast<const unsigned char *>(col_s)[(size_t)s * (ROWS + 16) + tid] : 0u;
      const bool esc = PAT && pid == 255u;
      const int32_t *tab = ptab_s + (esc ? 0u : pid) * (unsigned)a.pat_w;
      const int32_t *gc = a.colval + rs;
      // BATCH independent (col,val) shared-memory reads and x gathers are issued before the dependent,
      // strictly in-order accumulation; indices past the row end are clamped to the last entry (a
      // redundant, cached load) instead of predicating the loads.
      for (int k0 = 0; k0 < len; k0 += BATCH) {
        double v[BATCH], xv[BATCH];
        int32_t c[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
          const int kk = min(k0 + u, len - 1);
          if (PAT) {
            const int32_t tw = tab[esc ? 0 : kk];
            c[u] = (int32_t)row + tw;
            if (esc || tw == PA_PAT_GHOST) c[u] = __ldg(gc + kk);
          } else {
            c[u] = cs[kk];
          }
          v[u] = vs[kk];
        }
        bool ghost = false;
        if (MODE == 1) {
#pragma unroll
          for (int u = 0; u < BATCH; ++u) ghost |= c[u] >= a.n_own_cols;
        }
        if (MODE == 2) {  // own block only: ghost columns are skipped here and added, in order, by k_spmv_ghost_rows
          const int32_t last_own = (int32_t)a.n_own_cols - 1;
#pragma unroll
          for (int u = 0; u < BATCH; ++u) xv[u] = __ldg(a.x + min(c[u], last_own));
        } else if (MODE == 4) {
          // every tile reads x with plain (weak) loads through the writable alias — never the read-only (.nc) path: the ghost
          // slots are written by this very kernel.  Tiles that hold a ghost column have waited above (acquire fence: the SM's
          // L1 is invalidated after the last gather store), the others never touch a ghost slot.  No branch per batch: the
          // instruction stream is that of MODE 0 (a per-batch test cost 9.6 % more instructions = 8.6 % more time, ncu).
          // (a non-volatile asm without memory clobber: the compiler schedules it like __ldg — with the C++ load through the
          // writable alias it gave up the 128-bit shared-memory reads of the row, 2688 instead of 1600 SASS instructions; the
          // load cannot move above the wait: its address comes from the tile, which is read behind the mbarrier wait)
#pragma unroll
          for (int u = 0; u < BATCH; ++u) asm("ld.global.f64 %0, [%1];" : "=d"(xv[u]) : "l"(a.x + c[u]));
        } else if (MODE == 0 || !ghost) {  // straight-line: all BATCH gathers are in flight together
#pragma unroll
          for (int u = 0; u < BATCH; ++u) xv[u] = __ldg(a.x + c[u]);
        } else {  // boundary rows: ghost columns come from the owner's HBM over NVLink
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            if (c[u] >= a.n_own_cols) {
              const int64_t g = c[u] - a.n_own_cols;
              xv[u] = __ldcg(a.peers.p[a.gslot[g]] + a.grlid[g]);
            } else {
              xv[u] = __ldg(a.x + c[u]);
            }
          }
        }
        // ptxas otherwise sinks every load next to its use to hit a 32-register target, serialising the
        // gathers; pinning all BATCH values live here keeps the loads back to back (memory-level parallelism).
        keep_live(xv);
        keep_live(v);
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {  // branch-free, strictly in column order
          const double t = __dadd_rn(acc, __dmul_rn(v[u], xv[u]));
          const bool take = (k0 + u < len) && (MODE != 2 || c[u] < a.n_own_cols);
          acc = take ? t : acc;
        }
      }
      mbar_arrive(empty + s);
      const int64_t yi = a.rowmap ? (int64_t)a.rowmap[row] : row;
      if (a.alpha == 1.0 && a.beta == 0.0) {
        a.y[yi] = acc;
        if (a.dotw) dsum = __dadd_rn(dsum, __dmul_rn(acc, __ldg(a.dotw + yi)));  // no FMA anywhere in this kernel (SASS-checked)
      } else {
        const double by = a.beta == 0.0 ? 0.0 : __dmul_rn(a.beta, a.y[yi]);
        a.y[yi] = __dadd_rn(__dmul_rn(a.alpha, acc), by);
      }
    } else {
      mbar_arrive(empty + s);
    }
  }
  spmv_dot_epilogue(a, dsum, tid, ROWS);
}

// ------------------------------------------------------------------ pattern kernel for SHORT rows: RPT rows per thread
// With the column stream gone, the 7-pt product is no longer bound by HBM but by how many gathers a consumer thread keeps in
// flight: one row of 7 entries per tile and thread.  Here a tile holds RPT x ROWS rows and every consumer thread owns RPT of
// them (tid, tid + ROWS, ...): their entries are read, their x values gathered and their sums advanced side by side —
// RPT x BATCH loads in flight per thread, the same products in the same order per row.  MODE 0 only (columns are local).
// FL > 0: rows of exactly FL entries whose pattern is in the table take a FIXED-LENGTH path when the whole warp agrees (the
// interior of a stencil operator): no clamped indices, no predicated accumulation, shared-memory offsets are immediates —
// ~8 instead of ~14 instructions per entry (ncu: the generic loop issues 300-370 warp instructions per 7-entry row and keeps
// the issue slots 56-66 % busy, which is what bounds the kernel once the column stream is gone).
template <typename PtrT, int BATCH, int RPT, int MINB, int FL>
__global__ void __launch_bounds__(288, MINB) k_spmv_pat(const SpmvArgs<PtrT> a, const TmaCfg cfg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int ROWS = cfg.rows, CAP = cfg.cap, S = cfg.stages, TR = ROWS * RPT;
  // layout: val[S][CAP+2] | pattern ids [S][TR+16 bytes] | p0[S] (int64) | full[S] | empty[S] | table
  double *val_s = reinterpret_cast<double *>(smem_raw);
  unsigned char *pat_s = reinterpret_cast<unsigned char *>(val_s + (size_t)S * (CAP + 2));
  int64_t *p0_s = reinterpret_cast<int64_t *>(pat_s + (size_t)S * (TR + 16));
  uint64_t *full = reinterpret_cast<uint64_t *>(p0_s + S);
  uint64_t *empty = full + S;
  int32_t *ptab_s = reinterpret_cast<int32_t *>(empty + S);
  unsigned char *pg_s = reinterpret_cast<unsigned char *>(ptab_s + a.npat * a.pat_w);  // [256] pattern reads colval (ghost marker / none)
  // EARLY (long rows): the pattern id is ALSO read from global memory one tile ahead, like the row pointers, so that rows which
  // will read columns from colval (ghost marker, no pattern) can bring those lines into L1 before the tile arrives: the read
  // sits in front of an x gather and one late row holds up its whole tile.  (27-pt, two parts: 2.96 -> 2.77 ms.  For 7-entry
  // rows the extra byte load and test per row cost more than they save: 1.68 -> 2.03 ms, so short rows do without.)
  constexpr bool EARLY = FL >= 16 || BATCH >= 16;
  const int tid = threadIdx.x;
  for (int i = tid; i < a.npat * a.pat_w; i += blockDim.x) ptab_s[i] = a.ptab[i];
  if (EARLY)
    for (int p = tid; p < 256; p += blockDim.x) {
      bool g = p >= a.npat;
      if (!g)
        for (int k2 = 0; k2 < a.pat_w; ++k2) g |= a.ptab[p * a.pat_w + k2] == PA_PAT_GHOST;
      pg_s[p] = g ? 1 : 0;
    }
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, ROWS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int64_t nloc = first < cfg.ntiles ? (cfg.ntiles - first + stride - 1) / stride : 0;
  if (tid >= ROWS) {
    if (tid == ROWS) {  // producer: one elected lane drives the TMA ring
      const uint64_t pol = stream_policy();
      int s = 0;
      uint32_t ph = 1;
      for (int64_t j = 0; j < nloc; ++j, ++s) {
        if (s == S) { s = 0; ph ^= 1u; }
        if (j >= S) mbar_wait(empty + s, ph);
        const int64_t t = first + j * stride;
        const int64_t r0 = t * TR, r1 = min(r0 + (int64_t)TR, a.nrows);
        const int64_t p0 = (int64_t)a.rowptr[r0], p1 = (int64_t)a.rowptr[r1];
        p0_s[s] = p0;
        const int64_t pv = p0 & ~(int64_t)1;
        const uint32_t bv = (uint32_t)(((p1 - pv + 1) & ~(int64_t)1) * 8), bp = (uint32_t)(((r1 - r0) + 15) & ~(int64_t)15);
        const bool any = p1 > p0;
        mbar_expect_tx(full + s, (any ? bv : 0u) + bp);
        if (any) tma_load_1d(val_s + (size_t)s * (CAP + 2), a.nzval + pv, bv, full + s, pol);
        tma_load_1d(pat_s + (size_t)s * (TR + 16), a.pat + r0, bp, full + s, pol);
      }
    }
    return;
  }
  double dsum = 0.0;
  int64_t rs_n[RPT], re_n[RPT];
  unsigned pid_n[RPT];
#pragma unroll
  for (int q = 0; q < RPT; ++q) {
    rs_n[q] = re_n[q] = 0;
    pid_n[q] = 255u;
    const int64_t row = first * TR + q * ROWS + tid;
    if (nloc > 0 && row < a.nrows) {
      rs_n[q] = (int64_t)a.rowptr[row];
      re_n[q] = (int64_t)a.rowptr[row + 1];
      if (EARLY) pid_n[q] = (unsigned)__ldg(a.pat + row);
    }
  }
  int s = 0;
  uint32_t ph = 0;
  for (int64_t j = 0; j < nloc; ++j, ++s) {
    if (s == S) { s = 0; ph ^= 1u; }
    const int64_t t = first + j * stride;
    int64_t rs[RPT];
    int len[RPT];
    int32_t row[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      rs[q] = rs_n[q];
      len[q] = (int)(re_n[q] - rs_n[q]);
      row[q] = (int32_t)(t * TR + q * ROWS + tid);
      if (EARLY && len[q] > 0 && pg_s[pid_n[q]]) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(a.colval + rs[q]));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(a.colval + rs[q] + len[q] - 1));
      }
      rs_n[q] = re_n[q] = 0;
      pid_n[q] = 255u;
      const int64_t rown = (t + stride) * TR + q * ROWS + tid;  // the next tile's row pointers while this tile is processed
      if (j + 1 < nloc && rown < a.nrows) {
        rs_n[q] = (int64_t)a.rowptr[rown];
        re_n[q] = (int64_t)a.rowptr[rown + 1];
        if (EARLY) pid_n[q] = (unsigned)__ldg(a.pat + rown);
      }
    }
    mbar_wait(full + s, ph);
    const int64_t p0 = p0_s[s];
    const double *vs[RPT];
    const int32_t *tab[RPT], *gc[RPT];
    bool esc[RPT];
    double acc[RPT];
    int maxlen = 0;
    bool fixed = FL > 0;
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      vs[q] = val_s + (size_t)s * (CAP + 2) + (len[q] > 0 ? rs[q] - (p0 & ~(int64_t)1) : 0);
      // (rows past the end of the matrix: the TMA copy stops at the last row, what lies behind it in the stage is stale)
      const unsigned pid = (int64_t)row[q] < a.nrows ? (unsigned)pat_s[(size_t)s * (TR + 16) + q * ROWS + tid] : 255u;
      esc[q] = pid == 255u;
      tab[q] = ptab_s + (esc[q] ? 0u : pid) * (unsigned)a.pat_w;
      gc[q] = a.colval + rs[q];
      acc[q] = 0.0;
      maxlen = max(maxlen, len[q]);  // (rows past the end of the matrix have length 0)
      fixed = fixed && len[q] == FL && !esc[q];
    }
    if (FL > 0 && __all_sync(0xffffffffu, fixed)) {
      // ---- fixed-length path: FB entries of every row in flight, constant offsets
      constexpr int FB = FL > 0 ? (FL <= 9 ? FL : 9) : 1;
#pragma unroll
      for (int k0 = 0; k0 < FL; k0 += FB) {
        double v[RPT][FB], xv[RPT][FB];
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
          for (int u = 0; u < FB; ++u)
            if (k0 + u < FL) {
              v[q][u] = vs[q][k0 + u];
              const int32_t tw = tab[q][k0 + u];
              int32_t c = row[q] + tw;
              if (tw == PA_PAT_GHOST) c = __ldg(gc[q] + (k0 + u));  // a ghost column (rows on a part face): one read of colval
              xv[q][u] = __ldg(a.x + c);
            }
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
          for (int u = 0; u < FB; ++u)
            if (k0 + u < FL) acc[q] = __dadd_rn(acc[q], __dmul_rn(v[q][u], xv[q][u]));  // strictly in column order
      }
    } else {
      for (int k0 = 0; k0 < maxlen; k0 += BATCH) {
        double v[RPT][BATCH], xv[RPT][BATCH];
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            const int kk = max(min(k0 + u, len[q] - 1), 0);  // past the row end: a redundant, cached read of its last entry
            const int32_t tw = tab[q][esc[q] ? 0 : kk];
            int32_t c = row[q] + tw;
            if (esc[q] || tw == PA_PAT_GHOST) c = __ldg(gc[q] + kk);
            v[q][u] = vs[q][kk];
            xv[q][u] = __ldg(a.x + (len[q] > 0 ? c : 0));
          }
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
          keep_live(xv[q]);
          keep_live(v[q]);
        }
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {  // branch-free, strictly in column order
            const double t2 = __dadd_rn(acc[q], __dmul_rn(v[q][u], xv[q][u]));
            acc[q] = (k0 + u < len[q]) ? t2 : acc[q];
          }
      }
    }
    mbar_arrive(empty + s);
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
      if ((int64_t)row[q] < a.nrows) {
        const int64_t yi = a.rowmap ? (int64_t)a.rowmap[row[q]] : (int64_t)row[q];
        if (a.alpha == 1.0 && a.beta == 0.0) {
          a.y[yi] = acc[q];
          if (a.dotw) dsum = __dadd_rn(dsum, __dmul_rn(acc[q], __ldg(a.dotw + yi)));
        } else {
          const double by = a.beta == 0.0 ? 0.0 : __dmul_rn(a.beta, a.y[yi]);
          a.y[yi] = __dadd_rn(__dmul_rn(a.alpha, acc[q]), by);
        }
      }
    }
  }
  spmv_dot_epilogue(a, dsum, tid, ROWS);
}

__global__ void k_tile_ghost_flags(const int32_t *grows, int64_t n, int rows, unsigned char *flag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) flag[grows[i] / rows] = 1;
}

// largest nnz of any ROWS-row tile (decides whether a tile fits one TMA stage)
template <typename PtrT>
__global__ void k_max_tile_nnz(const PtrT *rowptr, int64_t nrows, int rows, unsigned long long *out) {
  const int64_t ntiles = (nrows + rows - 1) / rows;
  unsigned long long m = 0;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r0 = t * rows, r1 = min(r0 + (int64_t)rows, nrows);
    m = max(m, (unsigned long long)(rowptr[r1] - rowptr[r0]));
  }
  atomicMax(out, m);
}

static int max_tile_nnz(pa_ctx *c, MatPart &m, int rows, int64_t *out) {
  auto it = m.tile_nnz.find(rows);
  if (it == m.tile_nnz.end()) {
    unsigned long long *d = nullptr, h = 0;
    PA_CUDA(cudaMalloc((void **)&d, sizeof(h)));
    PA_CUDA(cudaMemsetAsync(d, 0, sizeof(h), c->stream));
    if (m.ptr64)
      k_max_tile_nnz<int64_t><<<148 * 4, 256, 0, c->stream>>>((const int64_t *)m.d_rowptr, m.nrows, rows, d);
    else
      k_max_tile_nnz<int32_t><<<148 * 4, 256, 0, c->stream>>>((const int32_t *)m.d_rowptr, m.nrows, rows, d);
    PA_CUDA(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    PA_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d);
    it = m.tile_nnz.emplace(rows, (int64_t)h).first;
  }
  *out = it->second;
  return PA_OK;
}

template <typename PtrT>
static int launch_spmv_tma(pa_ctx *c, MatPart &m, const SpmvArgs<PtrT> &a, int mode, const TmaCfg &cfg, int ctas_per_sm, int64_t *grid_out) {
  const bool pat = mode == 0 && a.pat != nullptr;
  if (pat && cfg.use_pat_kernel) {
    const size_t smem2 = (size_t)cfg.stages * ((cfg.cap + 2) * 8 + (cfg.rows * cfg.rpt + 16) + 8 + 16) + (size_t)a.npat * a.pat_w * 4 + 256 + 128;
    void (*k2)(const SpmvArgs<PtrT>, const TmaCfg) = nullptr;
    // (rows per thread, fixed length): 7-entry rows 2 x 7 or 1 x 7, 27-entry rows 1 x 27; everything else the generic loop
    if (cfg.fl == 7) k2 = cfg.rpt == 2 ? k_spmv_pat<PtrT, 8, 2, 2, 7> : k_spmv_pat<PtrT, 8, 1, 3, 7>;
    else if (cfg.fl == 27) k2 = k_spmv_pat<PtrT, 16, 1, 2, 27>;
    else k2 = cfg.rpt == 2 ? k_spmv_pat<PtrT, 8, 2, 2, 0> : (cfg.batch >= 16 ? k_spmv_pat<PtrT, 16, 1, 2, 0> : k_spmv_pat<PtrT, 8, 1, 3, 0>);
    PA_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    if (ctas_per_sm <= 0) {
      PA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k2, cfg.rows + 32, smem2));
      if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    int nsm2 = 148;
    cudaDeviceGetAttribute(&nsm2, cudaDevAttrMultiProcessorCount, c->device);
    const int64_t grid2 = std::min<int64_t>(cfg.ntiles, (int64_t)nsm2 * ctas_per_sm);
    PA_CHECK(!a.dotw || grid2 <= PA_DOT_PARTS, PA_ESTATE, "dot epilogue: grid larger than the partial buffer");
    k2<<<(unsigned)grid2, cfg.rows + 32, smem2, c->stream>>>(a, cfg);
    *grid_out = grid2;
    return PA_OK;
  }
  const size_t smem = pat ? (size_t)cfg.stages * ((cfg.cap + 2) * 8 + (cfg.rows + 16) + 8 + 16) + (size_t)a.npat * a.pat_w * 4 + 128
                          : (size_t)cfg.stages * ((cfg.cap + 2) * 8 + (cfg.cap + 8) * 4 + 8 + 16) + 128;
  const int batch = cfg.batch;
  auto kern = pat ? (batch >= 32 ? k_spmv_tma<PtrT, 0, 32, true> : batch >= 16 ? k_spmv_tma<PtrT, 0, 16, true> : k_spmv_tma<PtrT, 0, 8, true>)
            : mode == 1 ? (batch >= 32 ? k_spmv_tma<PtrT, 1, 32> : batch >= 16 ? k_spmv_tma<PtrT, 1, 16> : k_spmv_tma<PtrT, 1, 8>)
            : mode == 4 ? (batch >= 32 ? k_spmv_tma<PtrT, 4, 32> : batch >= 16 ? k_spmv_tma<PtrT, 4, 16> : k_spmv_tma<PtrT, 4, 8>)
            : mode == 2 ? (batch >= 32 ? k_spmv_tma<PtrT, 2, 32> : batch >= 16 ? k_spmv_tma<PtrT, 2, 16> : k_spmv_tma<PtrT, 2, 8>)
                        : (batch >= 32 ? k_spmv_tma<PtrT, 0, 32> : batch >= 16 ? k_spmv_tma<PtrT, 0, 16> : k_spmv_tma<PtrT, 0, 8>);
  PA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ctas_per_sm <= 0) {  // persistent grid = SMs x resident CTAs (registers, shared memory and threads permitting)
    PA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, cfg.rows + 32, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
  int64_t grid = std::min<int64_t>(cfg.ntiles, (int64_t)nsm * ctas_per_sm);
  PA_CHECK(!a.dotw || grid <= PA_DOT_PARTS, PA_ESTATE, "dot epilogue: grid larger than the partial buffer");
  if (mode == 4 && m.arrive_grid != grid) {  // the arrival counter advances by `grid` per launch: restart it when the grid changes
    PA_CUDA(cudaMemsetAsync(m.d_arrive, 0, sizeof(unsigned long long), c->stream));
    m.arrive_grid = grid;
  }
  kern<<<(unsigned)grid, cfg.rows + 32, smem, c->stream>>>(a, cfg);
  *grid_out = grid;
  return PA_OK;
}

// ------------------------------------------------------------------ ghost block: c_own += A_oh * x_ghost
// Rows that reference ghost columns are listed once per matrix (k_ghost_scan).  Ghost columns are the tail of
// a row (own columns first: sorted CSR with own ids < ghost ids, or the own_own|own_ghost merge), so adding
// the tail entries one by one to the stored own-block result reproduces the sequential row sum bit for bit —
// the same split the reference uses (spmv! on own_own, then muladd! on own_ghost, src/p_sparse_matrix.jl:2099-2101).
template <typename PtrT>
__global__ void k_ghost_scan(const PtrT *rowptr, const int32_t *colval, int64_t nrows, int64_t n_own_cols, unsigned char *flag, int *unordered) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += (int64_t)gridDim.x * blockDim.x) {
    bool seen_ghost = false, bad = false;
    for (int64_t p = (int64_t)rowptr[row]; p < (int64_t)rowptr[row + 1]; ++p) {
      const bool g = colval[p] >= n_own_cols;
      bad |= seen_ghost && !g;
      seen_ghost |= g;
    }
    flag[row] = seen_ghost ? 1 : 0;
    if (bad) *unordered = 1;
  }
}

__global__ void k_compact_rows(const unsigned char *flag, int64_t nrows, int32_t *out, unsigned long long *count) {
  // order of the list is irrelevant (one thread per listed row, rows are independent)
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += (int64_t)gridDim.x * blockDim.x)
    if (flag[row]) out[atomicAdd(count, 1ull)] = (int32_t)row;
}

template <typename PtrT>
__global__ void k_spmv_ghost_rows(const SpmvArgs<PtrT> a, const int32_t *grows, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = grows[i];
    const int64_t yi = a.rowmap ? (int64_t)a.rowmap[row] : row;
    double acc = a.y[yi];
    for (int64_t p = (int64_t)a.rowptr[row]; p < (int64_t)a.rowptr[row + 1]; ++p) {
      const int32_t col = a.colval[p];
      if (col >= a.n_own_cols) acc = __dadd_rn(acc, __dmul_rn(a.nzval[p], a.x[col]));  // x ghost slot: refreshed by consistent!
    }
    a.y[yi] = acc;
  }
}

static int ghost_scan(pa_ctx *c, MatPart &m, int64_t n_own_cols) {
  if (m.ghost_scanned) return PA_OK;
  m.ghost_scanned = true;
  m.n_grows = 0;
  m.ghost_tail_ok = true;
  if (m.nrows == 0 || m.nnz == 0) return PA_OK;
  unsigned char *d_flag = nullptr;
  unsigned long long *d_count = nullptr, h_count = 0;
  int *d_bad = nullptr, h_bad = 0;
  PA_CUDA(cudaMalloc((void **)&d_flag, m.nrows));
  PA_CUDA(cudaMalloc((void **)&d_count, sizeof(h_count)));
  PA_CUDA(cudaMalloc((void **)&d_bad, sizeof(int)));
  PA_CUDA(cudaMemsetAsync(d_count, 0, sizeof(h_count), c->stream));
  PA_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), c->stream));
  if (m.ptr64)
    k_ghost_scan<int64_t><<<148 * 8, 256, 0, c->stream>>>((const int64_t *)m.d_rowptr, m.d_colval, m.nrows, n_own_cols, d_flag, d_bad);
  else
    k_ghost_scan<int32_t><<<148 * 8, 256, 0, c->stream>>>((const int32_t *)m.d_rowptr, m.d_colval, m.nrows, n_own_cols, d_flag, d_bad);
  // upper bound of listed rows is unknown before counting: count first with a dry compaction into a scratch of nrows
  int32_t *d_tmp = nullptr;
  PA_CUDA(cudaMalloc((void **)&d_tmp, m.nrows * sizeof(int32_t)));
  k_compact_rows<<<148 * 8, 256, 0, c->stream>>>(d_flag, m.nrows, d_tmp, d_count);
  PA_CUDA(cudaMemcpyAsync(&h_count, d_count, sizeof(h_count), cudaMemcpyDeviceToHost, c->stream));
  PA_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  m.n_grows = (int64_t)h_count;
  m.ghost_tail_ok = h_bad == 0;
  if (m.n_grows) {
    PA_CUDA(cudaMalloc((void **)&m.d_grows, m.n_grows * sizeof(int32_t)));
    PA_CUDA(cudaMemcpyAsync(m.d_grows, d_tmp, m.n_grows * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
    PA_CUDA(cudaStreamSynchronize(c->stream));
  }
  cudaFree(d_tmp);
  cudaFree(d_flag);
  cudaFree(d_count);
  cudaFree(d_bad);
  c->launches += 2;
  return PA_OK;
}

// mode 0: all columns local; 1: inline NVLink loads; 2: own block only; 3: ghost block only (boundary rows)
// sum of the per-CTA partials of the fused dot epilogue, fixed order
__global__ void k_sum_parts(const double *part, int n, double *out) {
  __shared__ double sm[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    *out = t;
  }
}

// dotw/d_out (nullable): fused epilogue *d_out = sum_parts dot(y_own, dotw_own), only with mode 0/1 on the TMA kernel
// ------------------------------------------------------------------ row patterns (column stream compression)
// pattern of a row = (length, column - row of every entry).  A structured operator has a handful (a stencil on a box: one per
// box position, 27); rows with ghost columns or rare boundary classes keep reading colval.  Built lazily for the TMA kernel:
// the candidates come from a sample of rows (host side: a dictionary of at most 254 tuples), then ONE pass over the matrix
// assigns every row the id of the tuple it matches EXACTLY, or 255.
#define PA_PAT_ESC 255
#define PA_PAT_MAXW 32
template <typename PtrT>
__global__ void k_pat_sample(const PtrT *rowptr, const int32_t *colval, int64_t nrows, int64_t n_own_cols, int64_t nsample, int32_t *out /* [nsample][1 + MAXW] */) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nsample; i += (int64_t)gridDim.x * blockDim.x) {
    // evenly spread rows with a pseudo-random offset inside every stride, plus the rows next to both ends
    const int64_t stride = nrows / nsample > 0 ? nrows / nsample : 1;
    int64_t row = i * stride + (int64_t)((unsigned long long)(i * 2654435761ull) % (unsigned long long)stride);
    if (i < 512) row = i;
    else if (i < 1024) row = nrows - 1 - (i - 512);
    if (row >= nrows) row = nrows - 1;
    if (row < 0) row = 0;
    const int64_t p0 = (int64_t)rowptr[row], len = (int64_t)rowptr[row + 1] - p0;
    int32_t *o = out + i * (1 + PA_PAT_MAXW);
    o[0] = len <= PA_PAT_MAXW ? (int32_t)len : -1;
    for (int k = 0; k < PA_PAT_MAXW; ++k) {
      int32_t dlt = 0;
      if (k < len && len <= PA_PAT_MAXW) {
        const int32_t col = colval[p0 + k];
        dlt = col >= n_own_cols ? PA_PAT_GHOST : col - (int32_t)row;  // a ghost column: "read it from colval"
      }
      o[1 + k] = dlt;
    }
  }
}
__device__ __forceinline__ unsigned long long pat_hash(int len, const int32_t *d) {
  unsigned long long h = 0x9E3779B97F4A7C15ull ^ (unsigned long long)len;
  for (int k = 0; k < len; ++k) {
    h ^= (unsigned long long)(unsigned)d[k] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
  }
  return h;
}
template <typename PtrT>
__global__ void k_pat_assign(const PtrT *rowptr, const int32_t *colval, int64_t nrows, int64_t n_own_cols, const int32_t *ptab, const int32_t *plen, const unsigned long long *phash,
                             int npat, int pat_w, unsigned char *pat, unsigned long long *n_esc) {
  extern __shared__ unsigned long long sh_hash[];  // [npat] then lengths
  int32_t *sh_len = reinterpret_cast<int32_t *>(sh_hash + npat);
  for (int i = threadIdx.x; i < npat; i += blockDim.x) {
    sh_hash[i] = phash[i];
    sh_len[i] = plen[i];
  }
  __syncthreads();
  unsigned long long esc = 0;
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p0 = (int64_t)rowptr[row];
    const int len = (int)((int64_t)rowptr[row + 1] - p0);
    int id = PA_PAT_ESC;
    if (len <= pat_w) {
      int32_t d[PA_PAT_MAXW];
      for (int k = 0; k < len; ++k) {
        const int32_t col = colval[p0 + k];
        d[k] = col >= n_own_cols ? PA_PAT_GHOST : col - (int32_t)row;
      }
      const unsigned long long h = pat_hash(len, d);
      for (int q = 0; q < npat && id == PA_PAT_ESC; ++q) {
        if (sh_hash[q] != h || sh_len[q] != len) continue;
        bool same = true;
        for (int k = 0; k < len; ++k) same &= ptab[q * pat_w + k] == d[k];  // exact: a hash collision must not change a column
        if (same) id = q;
      }
    }
    pat[row] = (unsigned char)id;
    esc += id == PA_PAT_ESC;
  }
  if (esc) atomicAdd(n_esc, esc);
}

static unsigned long long pat_hash_host(int len, const int32_t *d) {
  unsigned long long h = 0x9E3779B97F4A7C15ull ^ (unsigned long long)len;
  for (int k = 0; k < len; ++k) {
    h ^= (unsigned long long)(unsigned)d[k] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
  }
  return h;
}

static int build_patterns(pa_ctx *c, MatPart &m, int64_t n_own_cols) {
  m.pat_state = -1;
  if (m.nrows < pa_knob(c, "spmv_pattern_min_rows", 4096) || m.nnz == 0) return PA_OK;  // nothing to gain on small parts
  const int64_t nsample = std::min<int64_t>(m.nrows, 65536);
  int32_t *d_s = nullptr;
  PA_CUDA(cudaMalloc((void **)&d_s, nsample * (1 + PA_PAT_MAXW) * sizeof(int32_t)));
  if (m.ptr64) k_pat_sample<int64_t><<<256, 256, 0, c->stream>>>((const int64_t *)m.d_rowptr, m.d_colval, m.nrows, n_own_cols, nsample, d_s);
  else k_pat_sample<int32_t><<<256, 256, 0, c->stream>>>((const int32_t *)m.d_rowptr, m.d_colval, m.nrows, n_own_cols, nsample, d_s);
  std::vector<int32_t> hs((size_t)nsample * (1 + PA_PAT_MAXW));
  PA_CUDA(cudaMemcpyAsync(hs.data(), d_s, hs.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_s);
  c->launches++;
  // dictionary of the sampled tuples, most frequent first
  std::map<std::vector<int32_t>, int64_t> freq;
  int wmax = 0;
  for (int64_t i = 0; i < nsample; ++i) {
    const int32_t *o = hs.data() + i * (1 + PA_PAT_MAXW);
    if (o[0] < 0) continue;
    std::vector<int32_t> key(o, o + 1 + o[0]);
    freq[key]++;
  }
  // tuples seen only once or twice in a large sample are rows of their own kind (ghost columns on a part face): keeping them
  // would only grow the table (shared memory per CTA, occupancy) — they read colval instead
  const int64_t min_count = std::max<int64_t>(1, nsample / 8192);
  std::vector<std::pair<int64_t, std::vector<int32_t>>> order;
  for (auto &kv : freq)
    if (kv.second >= min_count) order.emplace_back(-kv.second, kv.first);
  std::sort(order.begin(), order.end());
  if (order.empty()) return PA_OK;
  if (order.size() > 254) order.resize(254);
  for (auto &e : order) wmax = std::max(wmax, e.second[0]);
  while (order.size() > 1 && order.size() * (size_t)std::max(wmax, 1) * 4 > 12 * 1024) order.pop_back();  // the table lives in shared memory
  wmax = 1;
  for (auto &e : order) wmax = std::max(wmax, e.second[0]);
  const int npat = (int)order.size();
  std::vector<int32_t> tab((size_t)npat * wmax, 0), len(npat);
  std::vector<unsigned long long> hash(npat);
  for (int q = 0; q < npat; ++q) {
    len[q] = order[q].second[0];
    for (int k = 0; k < len[q]; ++k) tab[(size_t)q * wmax + k] = order[q].second[1 + k];
    hash[q] = pat_hash_host(len[q], tab.data() + (size_t)q * wmax);
  }
  int32_t *d_len = nullptr;
  unsigned long long *d_hash = nullptr, *d_esc = nullptr, h_esc = 0;
  PA_CUDA(cudaMalloc((void **)&m.d_ptab, tab.size() * sizeof(int32_t)));
  PA_CUDA(cudaMalloc((void **)&m.d_pat, (size_t)m.nrows + 512));
  PA_CUDA(cudaMalloc((void **)&d_len, npat * sizeof(int32_t)));
  PA_CUDA(cudaMalloc((void **)&d_hash, npat * sizeof(unsigned long long)));
  PA_CUDA(cudaMalloc((void **)&d_esc, sizeof(unsigned long long)));
  PA_CUDA(cudaMemcpyAsync(m.d_ptab, tab.data(), tab.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  PA_CUDA(cudaMemcpyAsync(d_len, len.data(), npat * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  PA_CUDA(cudaMemcpyAsync(d_hash, hash.data(), npat * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
  PA_CUDA(cudaMemsetAsync(d_esc, 0, sizeof(unsigned long long), c->stream));
  PA_CUDA(cudaMemsetAsync(m.d_pat, PA_PAT_ESC, (size_t)m.nrows + 512, c->stream));
  const size_t sh = (size_t)npat * (sizeof(unsigned long long) + sizeof(int32_t));
  if (m.ptr64)
    k_pat_assign<int64_t><<<148 * 8, 256, sh, c->stream>>>((const int64_t *)m.d_rowptr, m.d_colval, m.nrows, n_own_cols, m.d_ptab, d_len, d_hash, npat, wmax, m.d_pat, d_esc);
  else
    k_pat_assign<int32_t><<<148 * 8, 256, sh, c->stream>>>((const int32_t *)m.d_rowptr, m.d_colval, m.nrows, n_own_cols, m.d_ptab, d_len, d_hash, npat, wmax, m.d_pat, d_esc);
  PA_CUDA(cudaMemcpyAsync(&h_esc, d_esc, sizeof(h_esc), cudaMemcpyDeviceToHost, c->stream));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_len); cudaFree(d_hash); cudaFree(d_esc);
  c->launches++;
  // rows without a pattern pay a global read of their columns: worth it only when they are few
  if ((double)h_esc > 0.08 * (double)m.nrows) {
    cudaFree(m.d_pat); cudaFree(m.d_ptab);
    m.d_pat = nullptr; m.d_ptab = nullptr;
    return PA_OK;
  }
  m.npat = npat;
  m.pat_w = wmax;
  m.pat_len0 = len[0];
  m.pat_state = 1;
  return PA_OK;
}

int pa_spmv_local(pa_mat *A, pa_vec *x, pa_vec *y, double alpha, double beta, int mode, const pa_vec *dotw, double *d_out, int fold) {
  pa_ctx *c = A->ctx;
  PA_CHECK(!fold || (dotw && pa_fold_ok(c)), PA_ESTATE, "folded dot epilogue needs one local part with mapped peers");
  for (int k = 0; k < c->nlocal; ++k) {
    MatPart &m = A->parts[k];
    if (m.nrows == 0) continue;
    const PlanPart &cp = A->cols->parts[k];
    const PlanPart &rp = A->rows->parts[k];
    int kmode = mode;
    const PlanPart &xp = x->plan->parts[k];
    if ((mode == 1 || mode == 2 || mode == 4) && (cp.n_ghost == 0 || (mode == 4 && xp.n_cons == 0))) kmode = 0;
    if (mode == 3 && m.n_grows == 0) continue;
    int rows = (int)pa_knob(c, "spmv_rows", 0);
    if (rows != 256 && rows != 128 && rows != 64 && rows != 32) rows = m.rows_per_cta;
    // TMA pipeline configuration (knobs allow sweeping on the GPU without recompiling)
    TmaCfg cfg;
    // spmv_patterns: 1 (default) = wherever the rows repeat patterns (build_patterns decides per part), 0 = never
    const int64_t pat_knob = pa_knob(c, "spmv_patterns", 1);
    const bool long_rows = m.nnz > 12 * m.nrows;
    bool want_pat = kmode == 0 && pat_knob != 0 && m.tma_ok && pa_knob(c, "spmv_kernel", 3) == 3;
    if (want_pat && m.pat_state == 0) PA_TRY(build_patterns(c, m, cp.prefix ? cp.n_own : cp.n_local));
    want_pat = want_pat && m.pat_state == 1;
    // measured best (profiles/): 7-pt 256 rows per tile, 27-pt 64 with the column stream / 128 with row patterns
    // with row patterns (k_spmv_pat, fixed-length path): 27-pt 96 rows x 2 stages (4.60 ms; column stream 7.40), 7-pt 2 x 128 rows per
    // tile x 3 stages (1.72 ms; column stream 1.96)
    cfg.rows = (int)pa_knob(c, "tma_rows", long_rows ? (want_pat ? 96 : 64) : (want_pat ? 128 : 256));
    cfg.rpt = want_pat ? (int)pa_knob(c, "spmv_pattern_rpt", long_rows ? 1 : 2) : 1;  // rows per consumer thread of the pattern kernel
    if (cfg.rpt != 2) cfg.rpt = 1;
    // spmv_pattern_kernel 1 (default) = k_spmv_pat (fixed-length path for the dominant pattern), 0 = k_spmv_tma<PAT> (generic loop)
    cfg.use_pat_kernel = want_pat && pa_knob(c, "spmv_pattern_kernel", 1) != 0;
    if (!cfg.use_pat_kernel) cfg.rpt = 1;
    cfg.fl = (cfg.use_pat_kernel && pa_knob(c, "spmv_pattern_fixed", 1) != 0 && (m.pat_len0 == 7 || m.pat_len0 == 27)) ? m.pat_len0 : 0;
    if (cfg.fl == 27) cfg.rpt = 1;
    cfg.stages = (int)pa_knob(c, "tma_stages", (want_pat && !long_rows) ? 3 : 2);
    cfg.batch = (int)pa_knob(c, "tma_batch", m.nnz > 8 * m.nrows ? 16 : 8);
    int ctas = (int)pa_knob(c, "tma_ctas", 0);
    bool use_tma = mode != 3 && m.tma_ok && pa_knob(c, "spmv_kernel", 3) == 3 && cfg.rows >= 32 && cfg.rows <= 256 && cfg.rows % 32 == 0 && cfg.stages >= 2;
    if (use_tma) {
      int64_t mt = 0;
      PA_TRY(max_tile_nnz(c, m, cfg.rows * cfg.rpt, &mt));
      cfg.cap = (int)((std::max<int64_t>(mt, 64) + 63) / 64 * 64);
      cfg.ntiles = (m.nrows + (int64_t)cfg.rows * cfg.rpt - 1) / ((int64_t)cfg.rows * cfg.rpt);
      const size_t smem = (size_t)cfg.stages * ((cfg.cap + 2) * 8 + (want_pat ? (cfg.rows * cfg.rpt + 16) : (cfg.cap + 8) * 4) + 8 + 16) + 128 + (want_pat ? 16 * 1024 : 0);
      if (smem > 200 * 1024) use_tma = false;  // a tile does not fit: irregular rows -> chunked kernel
    }
    PA_CHECK(use_tma || (kmode != 2 && kmode != 4), PA_ESTATE, "own-block / fused-exchange modes need the TMA kernel");
    const bool use_pat = want_pat && use_tma && kmode == 0;
    if (kmode == 4 && !m.d_arrive) PA_CUDA(cudaMalloc((void **)&m.d_arrive, sizeof(unsigned long long)));
    const unsigned char *tile_flags = nullptr;
    if (kmode == 4) {
      auto it = m.tile_ghost.find(cfg.rows);
      if (it == m.tile_ghost.end()) {
        unsigned char *f = nullptr;
        PA_CUDA(cudaMalloc((void **)&f, (size_t)cfg.ntiles));
        PA_CUDA(cudaMemsetAsync(f, 0, (size_t)cfg.ntiles, c->stream));
        if (m.n_grows) k_tile_ghost_flags<<<148 * 4, 256, 0, c->stream>>>(m.d_grows, m.n_grows, cfg.rows, f);
        c->launches++;
        it = m.tile_ghost.emplace(cfg.rows, f).first;
      }
      tile_flags = it->second;
    }
    auto fill = [&](auto &a) {
      a.nrows = m.nrows;
      a.colval = m.d_colval;
      a.nzval = m.d_nzval;
      a.x = x->d[k];
      a.y = y->d[k];
      a.rowmap = rp.prefix ? nullptr : rp.d_own_to_local;
      a.alpha = alpha;
      a.beta = beta;
      a.n_own_cols = cp.n_own;
      a.gslot = cp.d_gslot_by_gid;
      a.grlid = cp.d_grlid_by_gid;
      a.peers = pa_peer_ptrs(x, k);
      a.dotw = dotw ? dotw->d[k] : nullptr;
      a.dot_part = m.d_dotpart;
      a.xw = x->d[k];
      a.cons_lid = xp.d_ghost_lid;
      a.cons_slot = xp.d_ghost_slot;
      a.cons_rlid = xp.d_ghost_rlid;
      a.n_cons = xp.n_cons;
      a.arrive = m.d_arrive;
      a.fold = fold;
      a.dot_ticket = m.d_dot_ticket;
      if (fold) a.push = pa_red_push(c);
      a.tile_ghost = tile_flags;
      a.pat = use_pat ? m.d_pat : nullptr;
      a.ptab = m.d_ptab;
      a.npat = m.npat;
      a.pat_w = m.pat_w;
    };
    if (dotw) {
      PA_CHECK(use_tma && mode != 2 && mode != 3 && rp.prefix && alpha == 1.0 && beta == 0.0, PA_ESTATE, "dot epilogue unavailable for this configuration");
      if (!m.d_dotpart) PA_CUDA(cudaMalloc((void **)&m.d_dotpart, PA_DOT_PARTS * sizeof(double)));
      if (!m.d_dot_ticket) {
        PA_CUDA(cudaMalloc((void **)&m.d_dot_ticket, sizeof(unsigned)));
        PA_CUDA(cudaMemsetAsync(m.d_dot_ticket, 0, sizeof(unsigned), c->stream));
      }
      PA_CHECK(!fold || use_tma, PA_ESTATE, "folded dot epilogue needs the TMA kernel");
    }
    int64_t grid_used = 0;
    auto go = [&](auto &a, auto tag) -> int {
      using PtrT = decltype(tag);
      if (mode == 3) {
        const int64_t g = std::min<int64_t>((m.n_grows + 255) / 256, 148 * 8);
        k_spmv_ghost_rows<PtrT><<<(unsigned)g, 256, 0, c->stream>>>(a, m.d_grows, m.n_grows);
      } else if (use_tma) {
        PA_TRY(launch_spmv_tma<PtrT>(c, m, a, kmode, cfg, ctas, &grid_used));
      } else if (kmode == 1) {
        launch_spmv_t<PtrT, true>(a, rows, c->stream);
      } else {
        launch_spmv_t<PtrT, false>(a, rows, c->stream);
      }
      return PA_OK;
    };
    if (m.ptr64) {
      SpmvArgs<int64_t> a;
      a.rowptr = (const int64_t *)m.d_rowptr;
      fill(a);
      PA_TRY(go(a, (int64_t)0));
    } else {
      SpmvArgs<int32_t> a;
      a.rowptr = (const int32_t *)m.d_rowptr;
      fill(a);
      PA_TRY(go(a, (int32_t)0));
    }
    c->launches++;
    if (dotw && !fold) {
      k_sum_parts<<<1, 256, 0, c->stream>>>(m.d_dotpart, (int)grid_used, c->nlocal == 1 ? d_out : c->d_partial + k);
      c->launches++;
    }
  }
  PA_CUDA(cudaGetLastError());
  if (dotw && !fold) PA_TRY(pa_reduce_finish(c, d_out));
  return PA_OK;
}

// does every local part run the TMA kernel? (regular rows: a tile fits a stage; own-first layout)
static bool tma_usable(pa_mat *A, pa_vec *x) {
  pa_ctx *c = A->ctx;
  if (pa_knob(c, "spmv_kernel", 3) != 3) return false;
  for (int k = 0; k < c->nlocal; ++k) {
    MatPart &m = A->parts[k];
    if (!x->plan->parts[k].prefix || !A->rows->parts[k].prefix) return false;
    if (m.nrows == 0 || !m.tma_ok) return false;
    int rows = (int)pa_knob(c, "tma_rows", m.nnz > 12 * m.nrows ? 64 : 256);
    int64_t mt = 0;
    if (max_tile_nnz(c, m, rows, &mt) != PA_OK) return false;
    const int64_t cap = (std::max<int64_t>(mt, 64) + 63) / 64 * 64;
    if ((size_t)pa_knob(c, "tma_stages", 2) * ((cap + 2) * 8 + (cap + 8) * 4 + 24) + 128 > 200 * 1024) return false;
  }
  return true;
}
// can pa_spmv_dot fuse the dot into the SpMV for this matrix?
static bool dot_fusable(pa_mat *A, pa_vec *x) { return !pa_knob(A->ctx, "no_dot_fusion", 0) && tma_usable(A, x); }

static int check_mul_args(pa_mat *A, pa_vec *x, pa_vec *y) {
  PA_CHECK(A && x && y, PA_EINVAL, "pa_spmv: null argument");
  PA_CHECK(A->committed, PA_ESTATE, "pa_spmv: matrix not committed");
  PA_CHECK(x != y && x->offset != y->offset, PA_EINVAL, "pa_spmv: x and y must not alias");
  PA_CHECK(x->plan->ctx == A->ctx && y->plan->ctx == A->ctx, PA_EINVAL, "pa_spmv: operands live on different backends");
  for (int k = 0; k < A->ctx->nlocal; ++k) {
    // @boundscheck matching_own_indices / matching_ghost_indices (src/p_sparse_matrix.jl:2091-2093)
    const PlanPart &cp = A->cols->parts[k], &xp = x->plan->parts[k], &rp = A->rows->parts[k], &yp = y->plan->parts[k];
    PA_CHECK(cp.n_own == xp.n_own && cp.n_local == xp.n_local, PA_EINVAL,
             "pa_spmv: x does not match axes(A,2) on part %d (own %lld vs %lld, local %lld vs %lld)", A->ctx->part_ids[k] + 1,
             (long long)xp.n_own, (long long)cp.n_own, (long long)xp.n_local, (long long)cp.n_local);
    PA_CHECK(rp.n_own == yp.n_own && (rp.prefix == yp.prefix), PA_EINVAL, "pa_spmv: y does not match axes(A,1) on part %d", A->ctx->part_ids[k] + 1);
    PA_CHECK(yp.n_local >= rp.n_own, PA_EINVAL, "pa_spmv: y too short");
    if (A->subassembled)  // matching_ghost_indices(axes(a,1), axes(c,1)) (src/p_sparse_matrix.jl:2094-2097,2109-2111)
      PA_CHECK(y->plan == A->rows || (yp.n_local == rp.n_local && yp.signature == rp.signature), PA_EINVAL,
               "pa_spmv: a sub-assembled matrix needs c on the row partition of A, ghost rows included (part %d)", A->ctx->part_ids[k] + 1);
  }
  return PA_OK;
}

int pa_reduce_dev_to(const pa_vec *x, const pa_vec *y, int mode, double *d_out);

// mul! with an optional fused epilogue *d_out = dot(dotw, y) (device resident; CG's u.c)
bool pa_spmv_dot_foldable(pa_mat *A, pa_vec *x) { return pa_fold_ok(A->ctx) && dot_fusable(A, x); }

int pa_spmv_dot(pa_mat *A, pa_vec *x, pa_vec *y, double alpha, double beta, uint32_t flags, const pa_vec *dotw, double *d_out, int fold) {
  PA_TRY(check_mul_args(A, x, y));
  pa_ctx *c = A->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  PA_TRY(pa_before_write(c));
  // Strategy (all of them read ghost values straight from the owner's HBM over NVLink; none packs or sends):
  //  overlap (default): consistent!(x) runs on the side stream (peer-load gather into x's ghost slots) WHILE the
  //      own-block kernel streams A_oo*x_own; then the ghost-block kernel adds A_oh*x_ghost — the latency hiding of
  //      the reference's mul! (t=consistent!(b); spmv! own_own; wait(t); muladd! own_ghost), bit-identical row sums.
  //  inline (PA_SPMV_INLINE_PEER_LOADS): one kernel, ghost columns dereference the peer arena inside the SpMV.
  //  explicit (PA_SPMV_EXPLICIT_EXCHANGE): consistent!(x) first, then one purely local SpMV (HPCG mul_no_lat!).
  bool any_ghost = false, prefix = true, tail_ok = true, tma_ok = true;
  for (int k = 0; k < c->nlocal; ++k) {
    any_ghost |= x->plan->parts[k].n_ghost > 0;
    prefix &= x->plan->parts[k].prefix;
    tail_ok &= A->parts[k].ghost_tail_ok;
    tma_ok &= A->parts[k].tma_ok;
  }
  // default = the fastest measured on B200 (profiles/r01_multigpu_strategies.md): peer-load gather, then one local SpMV
  if (A->subassembled) {
    // !a.assembled (src/p_sparse_matrix.jl:2109-2142): own AND ghost rows are multiplied (c_local = beta*c_local + alpha*A_local*b_local,
    // own-block entries of every row first), then assemble!(c) adds the ghost-row results to their owners and zeroes the ghosts
    PA_CHECK(!fold, PA_ESTATE, "pa_spmv_dot: no fused dot for sub-assembled matrices");
    PA_TRY(pa_vec_consistent(x));
    PA_TRY(pa_before_write(c));
    PA_TRY(pa_spmv_local(A, x, y, alpha, beta, 0, nullptr, nullptr));
    PA_TRY(pa_vec_assemble(y));
    if (dotw) PA_TRY(pa_reduce_dev_to(dotw, y, 0, d_out));
    return PA_OK;
  }
  int strategy = (flags & PA_SPMV_INLINE_PEER_LOADS) ? 1 : ((flags & PA_SPMV_OVERLAP) ? 2 : ((flags & PA_SPMV_FUSED_EXCHANGE) ? 3 : 0));
  const int64_t forced = pa_knob(c, "spmv_strategy", -1);
  if (forced >= 0 && forced <= 3) strategy = (int)forced;
  if (strategy == 3 && !tma_usable(A, x)) strategy = 0;  // irregular rows / permuted layouts: explicit gather first
  if (!prefix) strategy = 0;                                         // permuted layouts: plain local kernel after consistent!
  if (strategy == 2 && (!tail_ok || !tma_ok || alpha != 1.0 || beta != 0.0 || pa_knob(c, "spmv_kernel", 3) != 3)) strategy = 0;
  if (!any_ghost) strategy = 0;
  const pa_vec *want_dot = dotw;
  if (dotw && (strategy == 2 || alpha != 1.0 || beta != 0.0 || !dot_fusable(A, x))) dotw = nullptr;
  // the exchange plan of x is the one that knows where the ghosts live (it equals the column plan)
  pa_plan *xp = x->plan;
  PA_CHECK(!fold || (dotw && (strategy == 0 || strategy == 3)), PA_ESTATE, "pa_spmv_dot: folded dot epilogue unavailable for this configuration");
  if (strategy == 0 && pa_fold_ok(c)) {
    // one local part: signal + wait + gather + "done" are ONE kernel; the neighbours may overwrite their vectors as soon
    // as the gather has finished (the local SpMV only reads this part's HBM)
    if (any_ghost || c->nparts > 1) PA_TRY(pa_consistent_sync(x));
    PA_TRY(pa_spmv_local(A, x, y, alpha, beta, 0, dotw, d_out, fold));
    if (want_dot && !dotw) PA_TRY(pa_reduce_dev_to(want_dot, y, 0, d_out));
    return PA_OK;
  }
  PA_TRY(pa_collective_begin(xp));
  if (strategy == 0) {
    PA_TRY(pa_launch_consistent(x));
    PA_TRY(pa_spmv_local(A, x, y, alpha, beta, 0, dotw, d_out, fold));
  } else if (strategy == 3) {
    PA_TRY(pa_spmv_local(A, x, y, alpha, beta, 4, dotw, d_out, fold));  // consistent!(x) happens inside the SpMV kernel
  } else if (strategy == 1) {
    PA_TRY(pa_spmv_local(A, x, y, alpha, beta, 1, dotw, d_out));
    if (!(flags & PA_SPMV_SKIP_GHOST_REFRESH)) PA_TRY(pa_launch_consistent(x));
  } else {
    PA_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    PA_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    cudaStream_t main_stream = c->stream;
    c->stream = c->side;  // consistent!(x) on the side stream: NVLink gather overlapped with the own-block product
    int rc = pa_launch_consistent(x);
    c->stream = main_stream;
    PA_TRY(rc);
    PA_CUDA(cudaEventRecord(c->ev_join, c->side));
    PA_TRY(pa_spmv_local(A, x, y, alpha, beta, 2, nullptr, nullptr));
    PA_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    PA_TRY(pa_spmv_local(A, x, y, alpha, beta, 3, nullptr, nullptr));
  }
  PA_TRY(pa_collective_end(xp));
  if (want_dot && !dotw) PA_TRY(pa_reduce_dev_to(want_dot, y, 0, d_out));  // unfused fallback: separate dot pass
  return PA_OK;
}

extern "C" int pa_spmv(pa_mat *A, pa_vec *x, pa_vec *y, double alpha, double beta, uint32_t flags) {
  return pa_spmv_dot(A, x, y, alpha, beta, flags, nullptr, nullptr);
}

// ------------------------------------------------------------------ matrix objects
extern "C" int pa_mat_create(pa_plan *rows, pa_plan *cols, pa_mat **out) {
  PA_CHECK(rows && cols && out && rows->committed && cols->committed, PA_ESTATE, "pa_mat_create: plans missing or not committed");
  PA_CHECK(rows->ctx == cols->ctx, PA_EINVAL, "pa_mat_create: row and column plans live on different backends");
  pa_mat *A = new pa_mat();
  A->uid = rows->ctx->next_uid++;
  A->ctx = rows->ctx;
  A->rows = rows;
  A->cols = cols;
  A->parts.resize(A->ctx->nlocal);
  *out = A;
  return PA_OK;
}

static void free_part(MatPart &m) {
  cudaFree(m.d_coo_perm);
  cudaFree(m.d_coo_seg);
  cudaFree(m.d_coo_valid);
  for (auto &kv : m.tile_ghost) cudaFree(kv.second);
  cudaFree(m.d_dotpart);
  cudaFree(m.d_dot_ticket);
  cudaFree(m.d_arrive);
  cudaFree(m.d_pat);
  cudaFree(m.d_ptab);
  cudaFree(m.d_grows);
  cudaFree(m.d_rowptr);
  cudaFree(m.d_colval);
  cudaFree(m.d_nzval);
  m = MatPart();
}

extern "C" int pa_mat_destroy(pa_mat *A) {
  if (!A) return PA_OK;
  cudaSetDevice(A->ctx->device);
  cudaStreamSynchronize(A->ctx->stream);
  pa_cg_drop_work(A->ctx, A, nullptr, false);
  for (auto &m : A->parts) free_part(m);
  pa_mat_drop_transpose(A);
  delete A;
  return PA_OK;
}

// The cached local transposes hold a COPY of the values: every value refresh of A invalidates them.
void pa_mat_drop_transpose(pa_mat *A) {
  if (!A->T) return;
  cudaStreamSynchronize(A->ctx->stream);
  for (auto &m : A->T->parts) free_part(m);
  delete A->T;
  A->T = nullptr;
}

static int64_t rd(const void *p, int bits, int64_t i) { return bits == 64 ? ((const int64_t *)p)[i] : (int64_t)((const int32_t *)p)[i]; }

static int choose_rows(int64_t nrows, int64_t nnz) {
  double avg = nrows ? (double)nnz / (double)nrows : 0.0;
  return avg <= 8.0 ? 256 : (avg <= 16.0 ? 128 : (avg <= 32.0 ? 64 : 32));
}

static int upload_csr(pa_ctx *c, MatPart &m, int64_t nrows, int64_t ncols, const std::vector<int64_t> &rp,
                      const std::vector<int32_t> &cv, const std::vector<double> &nz) {
  free_part(m);
  m.nrows = nrows;
  m.ncols = ncols;
  m.nnz = rp[nrows];
  m.ptr64 = m.nnz >= (1ll << 31);
  m.rows_per_cta = choose_rows(nrows, m.nnz);
  if (m.ptr64) {
    PA_CUDA(cudaMalloc(&m.d_rowptr, (nrows + 1) * sizeof(int64_t)));
    PA_CUDA(cudaMemcpyAsync(m.d_rowptr, rp.data(), (nrows + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  } else {
    std::vector<int32_t> r32(rp.begin(), rp.end());
    PA_CUDA(cudaMalloc(&m.d_rowptr, (nrows + 1) * sizeof(int32_t)));
    PA_CUDA(cudaMemcpyAsync(m.d_rowptr, r32.data(), (nrows + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    PA_CUDA(cudaStreamSynchronize(c->stream));
  }
  if (m.nnz) {
    PA_CUDA(cudaMalloc((void **)&m.d_colval, (m.nnz + PA_MAT_PAD) * sizeof(int32_t)));
    PA_CUDA(cudaMalloc((void **)&m.d_nzval, (m.nnz + PA_MAT_PAD) * sizeof(double)));
    PA_CUDA(cudaMemcpyAsync(m.d_colval, cv.data(), m.nnz * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    PA_CUDA(cudaMemcpyAsync(m.d_nzval, nz.data(), m.nnz * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  PA_CUDA(cudaStreamSynchronize(c->stream));
  m.set = true;
  return PA_OK;
}

static int mat_part_args(pa_mat *A, int32_t k, int64_t nrows, const char *who) {
  PA_CHECK(A && !A->committed, PA_ESTATE, "%s: matrix missing or already committed", who);
  PA_CHECK(k >= 0 && k < A->ctx->nlocal, PA_EINVAL, "%s: local part %d out of range", who, k);
  // own rows only (assembled matrices: the ghost-row blocks are empty, src/p_sparse_matrix.jl:1704-1705) or ALL local rows of an
  // own-first row partition (sub-assembled matrices, psparse(...; assemble=false): blocks ghost_own / ghost_ghost stored too)
  const PlanPart &rpp = A->rows->parts[k];
  PA_CHECK(nrows == rpp.n_own || (rpp.prefix && nrows == rpp.n_local), PA_EINVAL,
           "%s: %lld rows given, the row partition owns %lld (%lld local)", who, (long long)nrows, (long long)rpp.n_own, (long long)rpp.n_local);
  if (nrows != rpp.n_own) A->subassembled = true;
  PA_CUDA(cudaSetDevice(A->ctx->device));
  return PA_OK;
}

extern "C" int pa_mat_set_csr(pa_mat *A, int32_t k, int64_t nrows, int64_t ncols, int32_t index_base, int32_t ptr_bits,
                              int32_t col_bits, const void *rowptr, const void *colval, const double *nzval) {
  PA_TRY(mat_part_args(A, k, nrows, "pa_mat_set_csr"));
  PA_CHECK((ptr_bits == 32 || ptr_bits == 64) && (col_bits == 32 || col_bits == 64) && (index_base == 0 || index_base == 1) && rowptr,
           PA_EINVAL, "pa_mat_set_csr: bad index description");
  const PlanPart &cp = A->cols->parts[k];
  PA_CHECK(ncols == cp.n_local, PA_EINVAL, "pa_mat_set_csr: %lld columns given, the column partition has %lld local ids",
           (long long)ncols, (long long)cp.n_local);
  std::vector<int64_t> rp(nrows + 1);
  for (int64_t i = 0; i <= nrows; ++i) rp[i] = rd(rowptr, ptr_bits, i) - index_base;
  PA_CHECK(rp[0] == 0, PA_EINVAL, "pa_mat_set_csr: rowptr does not start at the index base");
  for (int64_t i = 0; i < nrows; ++i) PA_CHECK(rp[i + 1] >= rp[i], PA_EINVAL, "pa_mat_set_csr: rowptr not monotone at row %lld", (long long)i);
  int64_t nnz = rp[nrows];
  PA_CHECK(nnz == 0 || (colval && nzval), PA_EINVAL, "pa_mat_set_csr: null colval/nzval");
  std::vector<int32_t> cv(nnz);
  for (int64_t p = 0; p < nnz; ++p) {
    int64_t cidx = rd(colval, col_bits, p) - index_base;
    PA_CHECK(cidx >= 0 && cidx < ncols, PA_EINVAL, "pa_mat_set_csr: column id %lld out of range at entry %lld", (long long)(cidx + index_base), (long long)p);
    cv[p] = (int32_t)cidx;
  }
  std::vector<double> nz(nzval, nzval + nnz);
  return upload_csr(A->ctx, A->parts[k], nrows, ncols, rp, cv, nz);
}

extern "C" int pa_mat_set_csr_split(pa_mat *A, int32_t k, int64_t nrows, int32_t index_base, int32_t ptr_bits, int32_t col_bits,
                                    const void *rowptr_oo, const void *colval_oo, const double *nzval_oo,
                                    const void *rowptr_oh, const void *colval_oh, const double *nzval_oh) {
  PA_TRY(mat_part_args(A, k, nrows, "pa_mat_set_csr_split"));
  PA_CHECK((ptr_bits == 32 || ptr_bits == 64) && (col_bits == 32 || col_bits == 64) && (index_base == 0 || index_base == 1) && rowptr_oo && rowptr_oh,
           PA_EINVAL, "pa_mat_set_csr_split: bad index description");
  const PlanPart &cp = A->cols->parts[k];
  std::vector<int64_t> rp(nrows + 1, 0);
  for (int64_t i = 0; i < nrows; ++i) {
    int64_t a = rd(rowptr_oo, ptr_bits, i + 1) - rd(rowptr_oo, ptr_bits, i), b = rd(rowptr_oh, ptr_bits, i + 1) - rd(rowptr_oh, ptr_bits, i);
    PA_CHECK(a >= 0 && b >= 0, PA_EINVAL, "pa_mat_set_csr_split: rowptr not monotone at row %lld", (long long)i);
    rp[i + 1] = rp[i] + a + b;
  }
  int64_t nnz = rp[nrows];
  std::vector<int32_t> cv(nnz);
  std::vector<double> nz(nnz);
  for (int64_t i = 0; i < nrows; ++i) {
    int64_t q = rp[i];
    for (int64_t p = rd(rowptr_oo, ptr_bits, i) - index_base; p < rd(rowptr_oo, ptr_bits, i + 1) - index_base; ++p, ++q) {
      int64_t o = rd(colval_oo, col_bits, p) - index_base;
      PA_CHECK(o >= 0 && o < cp.n_own, PA_EINVAL, "pa_mat_set_csr_split: own column id out of range");
      cv[q] = cp.prefix ? (int32_t)o : cp.own_to_local[o];
      nz[q] = nzval_oo[p];
    }
    for (int64_t p = rd(rowptr_oh, ptr_bits, i) - index_base; p < rd(rowptr_oh, ptr_bits, i + 1) - index_base; ++p, ++q) {
      int64_t g = rd(colval_oh, col_bits, p) - index_base;
      PA_CHECK(g >= 0 && g < cp.n_ghost, PA_EINVAL, "pa_mat_set_csr_split: ghost column id out of range");
      cv[q] = cp.prefix ? (int32_t)(cp.n_own + g) : cp.ghost_to_local[g];
      nz[q] = nzval_oh[p];
    }
  }
  return upload_csr(A->ctx, A->parts[k], nrows, cp.n_local, rp, cv, nz);
}

// CSC -> CSR on the host (setup time).  Columns are visited in ascending order, so every CSR row lists its
// entries by ascending column: the row sum then adds the terms in the order spmv_csc! scatters them
// (fill!(b,0); for col ascending: b[row] += a*x[col], src/sparse_utils.jl:671-690) — bit-identical results.
static int csc_to_csr(int64_t nrows, int64_t ncols, int index_base, int ptr_bits, int idx_bits, const void *colptr, const void *rowval,
                      const double *nzval, int64_t keep_rows, std::vector<int64_t> &rp, std::vector<int32_t> &cv, std::vector<double> &nz) {
  const int64_t nnz = rd(colptr, ptr_bits, ncols) - index_base;
  PA_CHECK(rd(colptr, ptr_bits, 0) == index_base && nnz >= 0, PA_EINVAL, "CSC colptr does not start at the index base");
  rp.assign(keep_rows + 1, 0);
  for (int64_t p = 0; p < nnz; ++p) {
    const int64_t r = rd(rowval, idx_bits, p) - index_base;
    PA_CHECK(r >= 0 && r < nrows, PA_EINVAL, "CSC row id out of range");
    PA_CHECK(r < keep_rows, PA_EINVAL, "CSC matrix has entries in ghost rows: only assembled matrices are supported");
    rp[r + 1]++;
  }
  for (int64_t r = 0; r < keep_rows; ++r) rp[r + 1] += rp[r];
  cv.resize(nnz);
  nz.resize(nnz);
  std::vector<int64_t> cur(rp.begin(), rp.end() - 1);
  for (int64_t c = 0; c < ncols; ++c) {
    const int64_t a = rd(colptr, ptr_bits, c) - index_base, b = rd(colptr, ptr_bits, c + 1) - index_base;
    PA_CHECK(b >= a, PA_EINVAL, "CSC colptr not monotone");
    for (int64_t p = a; p < b; ++p) {
      const int64_t r = rd(rowval, idx_bits, p) - index_base;
      cv[cur[r]] = (int32_t)c;
      nz[cur[r]++] = nzval[p];
    }
  }
  return PA_OK;
}

/* SparseMatrixCSC local matrices (the reference's default storage, src/p_sparse_matrix.jl:1132-1135).
 * Unsplit: n_local_rows x n_local_cols with own rows first and empty ghost rows. */
extern "C" int pa_mat_set_csc(pa_mat *A, int32_t k, int64_t nrows, int64_t ncols, int32_t index_base, int32_t ptr_bits, int32_t idx_bits,
                              const void *colptr, const void *rowval, const double *nzval) {
  PA_CHECK(A && !A->committed && k >= 0 && k < A->ctx->nlocal, PA_ESTATE, "pa_mat_set_csc: bad matrix/part");
  PA_CHECK((ptr_bits == 32 || ptr_bits == 64) && (idx_bits == 32 || idx_bits == 64) && (index_base == 0 || index_base == 1) && colptr, PA_EINVAL,
           "pa_mat_set_csc: bad index description");
  const PlanPart &rpart = A->rows->parts[k];
  PA_CHECK(rpart.prefix && nrows >= rpart.n_own, PA_EINVAL, "pa_mat_set_csc: needs own rows first");
  std::vector<int64_t> rp;
  std::vector<int32_t> cv;
  std::vector<double> nz;
  PA_TRY(csc_to_csr(nrows, ncols, index_base, ptr_bits, idx_bits, colptr, rowval, nzval, rpart.n_own, rp, cv, nz));
  return pa_mat_set_csr(A, k, rpart.n_own, ncols, 0, 64, 32, rp.data(), cv.data(), nz.data());
}

/* Split format with CSC blocks: own_own (n_own_rows x n_own_cols) and own_ghost (n_own_rows x n_ghost_cols). */
extern "C" int pa_mat_set_csc_split(pa_mat *A, int32_t k, int64_t nrows, int32_t index_base, int32_t ptr_bits, int32_t idx_bits,
                                    const void *colptr_oo, const void *rowval_oo, const double *nzval_oo, const void *colptr_oh,
                                    const void *rowval_oh, const double *nzval_oh) {
  PA_CHECK(A && !A->committed && k >= 0 && k < A->ctx->nlocal, PA_ESTATE, "pa_mat_set_csc_split: bad matrix/part");
  PA_CHECK((ptr_bits == 32 || ptr_bits == 64) && (idx_bits == 32 || idx_bits == 64) && (index_base == 0 || index_base == 1) && colptr_oo && colptr_oh,
           PA_EINVAL, "pa_mat_set_csc_split: bad index description");
  const PlanPart &cp = A->cols->parts[k];
  std::vector<int64_t> rp1, rp2;
  std::vector<int32_t> cv1, cv2;
  std::vector<double> nz1, nz2;
  PA_TRY(csc_to_csr(nrows, cp.n_own, index_base, ptr_bits, idx_bits, colptr_oo, rowval_oo, nzval_oo, nrows, rp1, cv1, nz1));
  PA_TRY(csc_to_csr(nrows, cp.n_ghost, index_base, ptr_bits, idx_bits, colptr_oh, rowval_oh, nzval_oh, nrows, rp2, cv2, nz2));
  return pa_mat_set_csr_split(A, k, nrows, 0, 64, 32, rp1.data(), cv1.data(), nz1.data(), rp2.data(), cv2.data(), nz2.data());
}

extern "C" int pa_mat_commit(pa_mat *A) {
  PA_CHECK(A && !A->committed, PA_ESTATE, "pa_mat_commit: matrix missing or already committed");
  for (int k = 0; k < A->ctx->nlocal; ++k) PA_CHECK(A->parts[k].set, PA_ESTATE, "pa_mat_commit: local part %d not set", k);
  PA_CUDA(cudaSetDevice(A->ctx->device));
  for (int k = 0; k < A->ctx->nlocal; ++k) {
    const PlanPart &cp = A->cols->parts[k];
    MatPart &m = A->parts[k];
    m.tma_ok = true;
    if (cp.prefix && cp.n_ghost > 0) PA_TRY(ghost_scan(A->ctx, m, cp.n_own));
  }
  A->committed = true;
  return PA_OK;
}

extern "C" int pa_mat_nnz(const pa_mat *A, int32_t k, int64_t *out) {
  PA_CHECK(A && out && k >= 0 && k < A->ctx->nlocal && A->parts[k].set, PA_EINVAL, "pa_mat_nnz: bad arguments");
  *out = A->parts[k].nnz;
  return PA_OK;
}

extern "C" int pa_mat_nrows(const pa_mat *A, int32_t k, int64_t *out) {
  PA_CHECK(A && out && k >= 0 && k < A->ctx->nlocal && A->parts[k].set, PA_EINVAL, "pa_mat_nrows: bad arguments");
  *out = A->parts[k].nrows;
  return PA_OK;
}

__global__ void k_narrow(const int64_t *in, int32_t *out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (int32_t)in[i];
}

extern "C" int pa_mat_download_csr(const pa_mat *A, int32_t k, int64_t *rowptr, int32_t *colval, double *nzval) {
  PA_CHECK(A && k >= 0 && k < A->ctx->nlocal && A->parts[k].set, PA_EINVAL, "pa_mat_download_csr: bad arguments");
  pa_ctx *c = A->ctx;
  const MatPart &m = A->parts[k];
  PA_CUDA(cudaSetDevice(c->device));
  if (rowptr) {
    if (m.ptr64) {
      PA_CUDA(cudaMemcpyAsync(rowptr, m.d_rowptr, (m.nrows + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    } else {
      std::vector<int32_t> tmp(m.nrows + 1);
      PA_CUDA(cudaMemcpyAsync(tmp.data(), m.d_rowptr, (m.nrows + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
      PA_CUDA(cudaStreamSynchronize(c->stream));
      for (int64_t i = 0; i <= m.nrows; ++i) rowptr[i] = tmp[i];
    }
  }
  if (colval && m.nnz) PA_CUDA(cudaMemcpyAsync(colval, m.d_colval, m.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  if (nzval && m.nnz) PA_CUDA(cudaMemcpyAsync(nzval, m.d_nzval, m.nnz * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  return PA_OK;
}

__global__ void k_fill_f64(double *v, int64_t n, double a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = a;
}

extern "C" int pa_mat_fill_stored(pa_mat *A, double a) {
  PA_CHECK(A, PA_EINVAL, "pa_mat_fill_stored: null matrix");
  pa_ctx *c = A->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  for (auto &m : A->parts)
    if (m.set && m.nnz) {
      k_fill_f64<<<148 * 8, 256, 0, c->stream>>>(m.d_nzval, m.nnz, a);
      c->launches++;
    }
  PA_CUDA(cudaGetLastError());
  pa_mat_drop_transpose(A);  // mul!(c, transpose(A), b) must see the new values
  return PA_OK;
}

// ------------------------------------------------------------------ on-device stencil generators
struct StencilGeom {
  int kind;
  int64_t gn[3], lo[3], hi[3], b[3];
  int64_t n_own, ng;
  const int64_t *sorted_gid;
  const int32_t *gid_of_sorted;
  double diag, off, alpha;
};

__device__ __forceinline__ int span(int64_t g, int64_t n) { return 3 - (g == 0) - (g == n - 1); }

__global__ void k_stencil_count(StencilGeom s, int64_t *cnt) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row <= s.n_own; row += (int64_t)gridDim.x * blockDim.x) {
    if (row == s.n_own) {
      cnt[row] = 0;
      continue;
    }
    const int64_t ix = row % s.b[0], iy = (row / s.b[0]) % s.b[1], iz = row / (s.b[0] * s.b[1]);
    const int64_t gx = s.lo[0] + ix, gy = s.lo[1] + iy, gz = s.lo[2] + iz;
    const int cx = span(gx, s.gn[0]), cy = span(gy, s.gn[1]), cz = span(gz, s.gn[2]);
    cnt[row] = s.kind == 27 ? cx * cy * cz : 1 + (cx - 1) + (cy - 1) + (cz - 1);
  }
}

template <typename PtrT>
__global__ void k_stencil_fill(StencilGeom s, const PtrT *rowptr, int32_t *colval, double *nzval, double *rhs, int *err) {
  for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < s.n_own; row += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ix = row % s.b[0], iy = (row / s.b[0]) % s.b[1], iz = row / (s.b[0] * s.b[1]);
    const int64_t gx = s.lo[0] + ix, gy = s.lo[1] + iy, gz = s.lo[2] + iz;
    int64_t p = (int64_t)rowptr[row];
    const int64_t p0 = p;
    int32_t gc[27];
    double gv[27];
    int ngc = 0;
    for (int sz = -1; sz <= 1; ++sz)
      for (int sy = -1; sy <= 1; ++sy)
        for (int sx = -1; sx <= 1; ++sx) {
          if (s.kind == 7 && (abs(sx) + abs(sy) + abs(sz)) > 1) continue;
          const int64_t cx = gx + sx, cy = gy + sy, cz = gz + sz;
          if (cx < 0 || cx >= s.gn[0] || cy < 0 || cy >= s.gn[1] || cz < 0 || cz >= s.gn[2]) continue;
          const double v = (sx == 0 && sy == 0 && sz == 0) ? s.diag : s.off;
          if (cx >= s.lo[0] && cx < s.hi[0] && cy >= s.lo[1] && cy < s.hi[1] && cz >= s.lo[2] && cz < s.hi[2]) {
            colval[p] = (int32_t)((cx - s.lo[0]) + s.b[0] * ((cy - s.lo[1]) + s.b[1] * (cz - s.lo[2])));
            nzval[p++] = v;
          } else {
            const int64_t gid = cx + s.gn[0] * (cy + s.gn[1] * cz);
            int64_t l = 0, h = s.ng - 1, g = -1;
            while (l <= h) {
              const int64_t mid = (l + h) >> 1;
              const int64_t t = s.sorted_gid[mid];
              if (t == gid) { g = s.gid_of_sorted[mid]; break; }
              if (t < gid) l = mid + 1; else h = mid - 1;
            }
            if (g < 0) { *err = 2; g = 0; }
            const int32_t col = (int32_t)(s.n_own + g);
            int q = ngc++;
            while (q > 0 && gc[q - 1] > col) { gc[q] = gc[q - 1]; gv[q] = gv[q - 1]; --q; }
            gc[q] = col;
            gv[q] = v;
          }
        }
    for (int q = 0; q < ngc; ++q) { colval[p] = gc[q]; nzval[p++] = gv[q]; }
    if (rhs) rhs[row] = s.kind == 27 ? 27.0 - (double)(p - p0) : s.alpha * (double)(7 - (p - p0));
  }
}

extern "C" int pa_mat_set_stencil(pa_mat *A, int32_t k, int32_t kind, const int64_t *gn, const int64_t *lo, const int64_t *hi,
                                  int64_t ng, const int64_t *ghost_gid_sorted, const int32_t *ghost_id_of_sorted, pa_vec *rhs) {
  PA_CHECK(gn && lo && hi && (kind == 7 || kind == 27), PA_EINVAL, "pa_mat_set_stencil: bad arguments");
  StencilGeom s;
  s.kind = kind;
  s.n_own = 1;
  for (int d = 0; d < 3; ++d) {
    s.gn[d] = gn[d]; s.lo[d] = lo[d]; s.hi[d] = hi[d]; s.b[d] = hi[d] - lo[d];
    PA_CHECK(lo[d] >= 0 && hi[d] > lo[d] && hi[d] <= gn[d], PA_EINVAL, "pa_mat_set_stencil: bad box");
    s.n_own *= s.b[d];
  }
  PA_TRY(mat_part_args(A, k, s.n_own, "pa_mat_set_stencil"));
  pa_ctx *c = A->ctx;
  const PlanPart &cp = A->cols->parts[k];
  PA_CHECK(cp.prefix && cp.n_own == s.n_own && cp.n_ghost == ng, PA_EINVAL,
           "pa_mat_set_stencil: column partition does not match the box (own %lld/%lld, ghost %lld/%lld)", (long long)cp.n_own,
           (long long)s.n_own, (long long)cp.n_ghost, (long long)ng);
  PA_CHECK(ng == 0 || (ghost_gid_sorted && ghost_id_of_sorted), PA_EINVAL, "pa_mat_set_stencil: null ghost tables");
  if (rhs) PA_CHECK(rhs->plan->parts[k].n_own == s.n_own && rhs->plan->parts[k].prefix, PA_EINVAL, "pa_mat_set_stencil: rhs does not match");
  s.ng = ng;
  s.alpha = (double)(gn[0] + 1) * (double)(gn[1] + 1) * (double)(gn[2] + 1);  // prod(n_i+1), src/gallery.jl:36
  s.diag = kind == 7 ? s.alpha * 2 * 3 : 26.0;
  s.off = kind == 7 ? -s.alpha : -1.0;
  int64_t *d_sg = nullptr;
  int32_t *d_sl = nullptr;
  if (ng) {
    PA_CUDA(cudaMalloc((void **)&d_sg, ng * sizeof(int64_t)));
    PA_CUDA(cudaMalloc((void **)&d_sl, ng * sizeof(int32_t)));
    PA_CUDA(cudaMemcpyAsync(d_sg, ghost_gid_sorted, ng * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    PA_CUDA(cudaMemcpyAsync(d_sl, ghost_id_of_sorted, ng * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  }
  s.sorted_gid = d_sg;
  s.gid_of_sorted = d_sl;
  MatPart &m = A->parts[k];
  free_part(m);
  const int64_t n = s.n_own;
  int64_t *d_cnt = nullptr, *d_rp64 = nullptr;
  PA_CUDA(cudaMalloc((void **)&d_cnt, (n + 1) * sizeof(int64_t)));
  PA_CUDA(cudaMalloc((void **)&d_rp64, (n + 1) * sizeof(int64_t)));
  const int grid = 148 * 8;
  k_stencil_count<<<grid, 256, 0, c->stream>>>(s, d_cnt);
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_cnt, d_rp64, n + 1, c->stream);
  void *d_tmp = nullptr;
  PA_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
  PA_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_cnt, d_rp64, n + 1, c->stream));
  int64_t nnz = 0;
  PA_CUDA(cudaMemcpyAsync(&nnz, d_rp64 + n, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_tmp);
  cudaFree(d_cnt);
  m.nrows = n;
  m.ncols = cp.n_local;
  m.nnz = nnz;
  m.ptr64 = nnz >= (1ll << 31);
  m.rows_per_cta = choose_rows(n, nnz);
  PA_CUDA(cudaMalloc((void **)&m.d_colval, (nnz + PA_MAT_PAD) * sizeof(int32_t)));
  PA_CUDA(cudaMalloc((void **)&m.d_nzval, (nnz + PA_MAT_PAD) * sizeof(double)));
  double *d_rhs = nullptr;
  if (rhs) {
    PA_TRY(pa_before_write(c));
    PA_CUDA(cudaMemsetAsync(rhs->d[k], 0, rhs->plan->parts[k].n_local * sizeof(double), c->stream));
    d_rhs = rhs->d[k];
  }
  if (m.ptr64) {
    m.d_rowptr = d_rp64;
    k_stencil_fill<int64_t><<<grid, 256, 0, c->stream>>>(s, d_rp64, m.d_colval, m.d_nzval, d_rhs, c->d_err);
  } else {
    PA_CUDA(cudaMalloc(&m.d_rowptr, (n + 1) * sizeof(int32_t)));
    k_narrow<<<grid, 256, 0, c->stream>>>(d_rp64, (int32_t *)m.d_rowptr, n + 1);
    k_stencil_fill<int32_t><<<grid, 256, 0, c->stream>>>(s, (const int32_t *)m.d_rowptr, m.d_colval, m.d_nzval, d_rhs, c->d_err);
    c->launches++;
  }
  c->launches += 3;
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaStreamSynchronize(c->stream));
  if (!m.ptr64) cudaFree(d_rp64);
  cudaFree(d_sg);
  cudaFree(d_sl);
  int herr = 0;
  PA_CUDA(cudaMemcpy(&herr, c->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (herr == 2) {
    cudaMemset(c->d_err, 0, sizeof(int));
    pa_set_error("pa_mat_set_stencil: a stencil neighbour outside the own box is missing from the ghost table");
    return PA_EINVAL;
  }
  m.set = true;
  return PA_OK;
}
