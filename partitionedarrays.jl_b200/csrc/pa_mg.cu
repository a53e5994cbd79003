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
#include <map>
#include <cstdio>
#include <array>

#include "pa_device.cuh"
#include "pa_internal.h"

#define GS_THREADS 256
#ifndef GS_MINB8
#define GS_MINB8 4  // resident CTAs per SM of the 8-lanes-per-row kernel: 64 registers (4 bytes of spill); 3 CTAs = 76 registers
#endif
#ifndef GS_COLOR_MINB
#define GS_COLOR_MINB 4  // resident CTAs per SM of the multi-colour kernel (64 registers)
#endif
#define GS_SPIN_LIMIT (20000000000LL)  // ~10 s of SM cycles, then give up and flag an error instead of hanging

struct GsPart {
  int64_t n = 0;
  int nlev = 0;
  int32_t *d_rows = nullptr;  // own rows sorted by wavefront level (ascending); every level starts at a multiple
                              // of 8 (-1 = padding) so that the rows a warp takes together never depend on each other
  int64_t npad = 0;           // length of d_rows (multiple of 8)
  int maxlen = 0;             // longest stored row
  ulonglong2 *d_xe = nullptr; // [n] value of the row's last update + the sweep epoch it happened in (gs_publish)
  int epoch = 0;
  // batch kernel (k_gs_level): the same order with every level padded to a multiple of GSB_ROWS (one CTA batch),
  // level of every batch, batches per level, per-level completion counters
  int32_t *d_rows_b = nullptr;
  int64_t npad_b = 0;
  int32_t *d_batch_lev = nullptr;
  uint32_t *d_lev_nb = nullptr;
  unsigned long long *d_done = nullptr;
  unsigned long long sweeps = 0;  // batch sweeps run so far (the counters are never reset)
  bool geom = false;
  int kind = 0;  // 7 or 27 when the geometry hint was given
  int64_t dims[3] = {0, 0, 0}, w[3] = {0, 0, 0};
  // sliced-ELL copies of the matrix in sweep order (see GsOrder): [0] wavefront (bit-exact), [1] multi-colour
  struct GsOrder *ord[2] = {nullptr, nullptr};
};

// The smoother's matrix in the order the sweep visits it.  Rows are sorted by level (wavefront level of the sequential
// sweep's dependency DAG, or colour), every level padded to a multiple of 32 rows (row id -1), and stored as SELL-32:
// slice g holds rows [32g, 32g+32) with entry k of the 32 rows adjacent, cols[(g*W + k)*32 + lane] — so a warp that
// takes one row per lane reads every entry slot with ONE coalesced 256-byte (values) / 128-byte (columns) load and the
// matrix stream is perfectly sequential in HBM (no sector shared with rows of other levels, no row-pointer reads).
// A column word carries, besides the local column id (< 2^29): bit 30 = "updated earlier in the FORWARD sweep" (an own
// column of a lower level), bit 29 = own column; -1 = empty slot.  The flags make the kernel independent of how the
// order was obtained: lexicographic wavefront levels reproduce the reference's sequential sweep bit for bit, colours
// give the multi-colour smoother, with the same per-row arithmetic (s -= a*x[col] in CSR order, s += d*x[row], s /= d).
struct GsOrder {
  int nlev = 0, W = 0;
  int64_t npad = 0;            // padded rows (multiple of 32)
  int32_t *d_rows = nullptr;   // [npad]
  int32_t *d_cols = nullptr;   // [npad * W]
  double *d_vals = nullptr;    // [npad * W]
  std::vector<int64_t> lev_group;  // first slice of every level, size nlev + 1 (one launch per colour)
  // level-gated sweeps (MODE 2): level of every slice, slices per level, slices finished per level summed over all sweeps
  int32_t *d_slice_lev = nullptr;
  uint32_t *d_lev_n = nullptr;            // [nlev * GS_SUB] slices counting into every sub-counter
  unsigned long long *d_done = nullptr;   // [nlev * GS_SUB]
  int64_t *d_lev_first = nullptr;         // [nlev + 1]
  uint32_t *d_lev_tot = nullptr;          // [nlev] (fenced gate variant)
  unsigned long long *d_done2 = nullptr;  // [nlev]
  unsigned long long sweeps = 0, sweeps2 = 0;
  // row patterns of the colour kernel (one byte per row instead of a column word per entry): per level a table of
  // [npat][W] packed words ((col - row) mod 2^29 | flags, -1 = empty slot); pat[pos] = id within the row's level, 255 = none
  unsigned char *d_pat = nullptr;      // [npad]
  int32_t *d_ptab = nullptr;           // tables of all levels, back to back
  std::vector<int> lev_npat;           // [nlev]
  std::vector<int64_t> lev_ptab_off;   // [nlev] first word of the level's table
  bool patterns = false;
  // strip order (k_gs_strip): ntask tasks of nsteps slices each, task-major
  int ntask = 0;
  int64_t nsteps = 0;
};
#define GS_COL_FRESH (1 << 30)
#define GS_COL_OWN (1 << 29)
#define GS_COL_MASK ((1 << 29) - 1)

struct pa_gs {
  pa_mat *A = nullptr;
  std::vector<GsPart> parts;
  bool committed = false;
  int order = 0;  // PA_GS_LEXICOGRAPHIC (bit-exact with the reference) or PA_GS_MULTICOLOR
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
  int epoch, backward, zero_guess, prefetch;
  int keep;  // mark x and the published pairs evict-last in L2 (they are re-read by the rows of the next 7 levels)
};

// Publishing a row: the new value travels WITH its "done in this sweep" mark, so a waiting reader needs one
// round trip and the writer needs no fence between value and flag.  The 64-bit value is split over two 8-byte
// words, each carrying the sweep epoch in its upper half ({lo32, epoch}, {hi32, epoch}); an aligned 8-byte
// store is single-copy atomic, so a reader that sees the current epoch in BOTH words has both halves of the
// new value, however the 16-byte access is split on the way.
__device__ __forceinline__ double gs_wait_value(const ulonglong2 *slot, unsigned epoch, int *err, uint64_t keep) {
  unsigned long long w0, w1;
  long long t0 = 0;
  for (;;) {
    asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(w0), "=l"(w1) : "l"(slot), "l"(keep) : "memory");
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
__device__ __forceinline__ void gs_publish(ulonglong2 *slot, double *x, double s, unsigned epoch, uint64_t keep) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(s), e = (unsigned long long)epoch << 32;
  asm volatile("st.relaxed.gpu.global.L2::cache_hint.v2.u64 [%0], {%1, %2}, %3;" ::"l"(slot), "l"((bits & 0xffffffffull) | e), "l"((bits >> 32) | e), "l"(keep)
               : "memory");
  *x = s;  // the plain vector: read by later rows as an OLD value never, by the next kernels always
}
// x changes during the sweep: always through L2; the rows of the neighbouring levels come back for the same sectors
__device__ __forceinline__ double gs_ld_x(const double *p, uint64_t keep) {
  double v;
  asm volatile("ld.global.cg.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(keep));
  return v;
}
// L2 retention of x and the published (value, epoch) pairs: evict-last when asked for, else no preference
__device__ __forceinline__ uint64_t gs_keep_policy(int keep) {
  uint64_t pol;
  if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
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
  const uint64_t pol = gs_stream_policy(), keep = gs_keep_policy(a.keep);
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
      if (fresh) xv = gs_wait_value(a.xe + col, (unsigned)a.epoch, a.err, keep);
      else if (use) xv = gs_ld_x(a.x + col, keep);
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
      gs_publish(a.xe + row, a.x + row, s, (unsigned)a.epoch, keep);
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
__global__ void __launch_bounds__(GS_THREADS, (G >= 16 ? 6 : (G == 8 ? GS_MINB8 : 2))) k_gs_flow_pipe(const GsArgs<PtrT> a) {
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
  const uint64_t pol = gs_stream_policy(), keep = gs_keep_policy(a.keep);
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
    xoC = gs_ld_x(a.x + rC, keep);
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
      xoB = gs_ld_x(a.x + rB, keep);
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
      xv[j] = (use[j] && !fresh[j]) ? gs_ld_x(a.x + col, keep) : 0.0;  // OLD values and ghosts
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j)
      if (fresh[j]) xv[j] = gs_wait_value(a.xe + colC[j], epoch, a.err, keep);
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
      gs_publish(a.xe + row, a.x + row, s, epoch, keep);
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

// ------------------------------------------------------------------ the batch sweep kernel (rows of <= 32 entries)
// An alternative schedule of the same sweep, selected with the knob gs_kernel=1 (NOT the default: see the numbers
// below).  k_gs_flow_pipe spends ~300 issue slots per row (the ordered subtraction runs on one lane of a warp that
// holds a few rows).  Here a CTA takes a BATCH of GSB_ROWS = 256 rows of one wavefront level, 32 per warp (the batch
// list pads every level to a multiple of 256, so the rows of a batch never depend on each other), and the
// dependency is tracked per LEVEL instead of per row: a batch of level L gathers its NEW x values when the counter
// of level L-1 (L+1 in the backward sweep) says that all batches of that level have stored their rows -- which, by
// induction, covers every earlier level.  One thread per CTA polls the counter and releases (fence + increment)
// after the CTA barrier.  Per batch and warp:
//   before the gate  the 32 rows' entries are streamed from the matrix into registers, 8 lanes per row and 4 rows
//                    per step (coalesced 64-byte pieces of nzval / 32-byte pieces of colval), and the OLD x values
//                    (rows of later levels, ghosts) are requested;
//   behind the gate  the NEW x values are gathered (simply what x holds now: rows of later levels have not started),
//                    the products are parked in shared memory, one line per row, and every lane walks ITS OWN row's
//                    line: s = b - p0 - p1 - ... in CSR order, + d*x_old, / d, store; CTA barrier, release.
// The ordered chain costs 2-3 issue slots per row instead of ~80 and the whole sweep ~25 per row; arithmetic and
// order per row are those of k_gs_flow (bit-identical results, tests/test_gpu_hpcg_mg.py).
// Measured on B200 (27-pt, symmetric sweep, 512^3 / 256^3 / 128^3 / 64^3 rows; the dataflow kernel: 60 / 11 / 2.2 /
// 0.9 ms):  this kernel 88 / 16 / 7.2 / 3.3 ms; with the OLD values gathered behind the gate too 70 / 16 / 7.4 / 3.5;
// one warp per 32-row batch with per-level counters 93 / 33 / 12 / 4.5; with per-row (value, epoch) pairs 106 / 45 /
// 20 / 9.  Phase clocks of one CTA (gs_trace knob): gate >= 2000 cycles, NEW gather 1250 + products 1650 (LSU
// wavefronts and 32-byte sectors for 8-byte values), chain 900, barrier + fence + increment 1450: one level costs
// ~3 us however few rows it has, and 3578 levels x 1.5 rounds (148 batch slots, levels of up to 256 batches) add up
// to more than the per-row dataflow needs, whose hop is ~1 us because only the LAST missing value of a row sits on
// the critical path.
#define GSB_THREADS 256
#define GSB_ROWS 256
#define GSB_WARP_DOUBLES(nj) (32 * (8 * (nj) + 2) + 32)  // shared memory per warp, in doubles

template <typename PtrT>
struct GsLevelArgs {
  GsArgs<PtrT> g;                   // rows/npad = the batch list
  const int32_t *batch_lev;         // [npad/GSB_ROWS] wavefront level of every batch
  const uint32_t *lev_nb;           // [nlev] batches per level
  unsigned long long *done;         // [nlev] batches finished, summed over all batch sweeps so far
  unsigned long long sweep;         // 1-based count of batch sweeps on this part (target = sweep * lev_nb)
  int nlev;
  int trace;                        // > 0: CTA trace-1 prints its phase clocks (debugging aid)
  int sync_mode;                    // 0 = release + acquire fences; 1 / 2 = timing experiments only (drop the acquire / both fences)
};

__device__ __forceinline__ double gs_ld_old(const double *p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void gs_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned long long gs_ld_relaxed(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void gs_fence_acq_rel() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// Polling is a relaxed load (an acquire load costs an L1 invalidation per poll); the acquire fence follows once.
__device__ __noinline__ void gs_wait_level(const unsigned long long *cnt, unsigned long long target, int *err) {
  long long t0 = 0;
  while (gs_ld_relaxed(cnt) < target) {
    if (!t0) {
      t0 = clock64();
    } else if (clock64() - t0 > GS_SPIN_LIMIT) {
      *err = 3;
      break;
    }
  }
}

template <typename PtrT, int NJ, bool BACKWARD, bool ZERO>
__global__ void __launch_bounds__(GSB_THREADS, 1) k_gs_level(const GsLevelArgs<PtrT> la) {
  const GsArgs<PtrT> &a = la.g;
  constexpr int KMAX = 8 * NJ;      // entry slots per row
  constexpr int STRIDE = KMAX + 2;  // doubles per row line: 16-byte aligned and conflict-free for the 16-byte chain reads
  constexpr int NT = 8;             // steps per warp and batch (4 rows each)
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) double gsw_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *prod = gsw_smem + (size_t)warp * GSB_WARP_DOUBLES(NJ);  // [32][STRIDE] products of the warp's rows
  double *dg = prod + 32 * STRIDE;                              // [32] their diagonal coefficients
  const int sub = lane >> 3, hl = lane & 7;
  const int64_t nb = a.npad / GSB_ROWS;
  const int64_t W = gridDim.x;
  const uint64_t pol = gs_stream_policy();

  // batch bt of the sweep = batch bt of the list (forward) or nb-1-bt (backward: the list is walked from its end)
  auto row_at = [&](int64_t bt) -> int32_t {  // this thread's row of batch bt (-1 = padding or past the end)
    if (bt >= nb) return -1;
    const int64_t lb = BACKWARD ? nb - 1 - bt : bt;
    return a.rows[lb * GSB_ROWS + threadIdx.x];
  };
  auto meta = [&](int32_t r, int64_t &ps, int &cnt, double &bv, double &xo) {
    ps = 0; cnt = 0; bv = 0.0; xo = 0.0;
    if (r >= 0) {
      ps = (int64_t)a.rowptr[r];
      cnt = (int)((int64_t)a.rowptr[r + 1] - ps);
      bv = a.b[r];
      if (!ZERO) xo = gs_ld_old(a.x + r);  // nobody writes x[row] before this row does
    }
  };
  auto prefetch = [&](int64_t ps, int cnt) {
    if (a.prefetch && cnt > 0) {
      const double *vp = a.nzval + ps;
      const int32_t *cp = a.colval + ps;
      gs_prefetch_l2(vp);
      gs_prefetch_l2(cp);
      if (cnt > 16) gs_prefetch_l2(vp + 16);
      gs_prefetch_l2(vp + cnt - 1);
      gs_prefetch_l2(cp + cnt - 1);
    }
  };

  dg[lane] = 0.0;
  const int64_t first = blockIdx.x;
  int32_t r0 = row_at(first), r1 = row_at(first + W), r2 = row_at(first + 2 * W);
  int64_t ps0, ps1;
  int cnt0, cnt1;
  double b0, b1, xo0, xo1;
  meta(r0, ps0, cnt0, b0, xo0);
  meta(r1, ps1, cnt1, b1, xo1);
  __syncwarp();
  for (int64_t bt = first; bt < nb; bt += W) {
    // ---- keep the next batches in flight
    const int32_t r3 = row_at(bt + 3 * W);
    int64_t ps2;
    int cnt2;
    double b2, xo2;
    meta(r2, ps2, cnt2, b2, xo2);
    prefetch(ps1, cnt1);
    const int lev = la.batch_lev[BACKWARD ? nb - 1 - bt : bt];
    const int glev = BACKWARD ? lev + 1 : lev - 1;  // the level whose completion opens this one
    const bool tr = la.trace > 0 && (int)blockIdx.x == la.trace - 1 && threadIdx.x == 0;
    long long tk[7];
    if (tr) tk[0] = clock64();
    // ---- before the gate: the batch's matrix entries, 4 rows per step, and everything that does not depend on the
    // previous level: the OLD x values (rows of later levels, ghosts) are requested here, so that only the NEW
    // values are gathered on the critical path behind the gate
    const int kend = (__reduce_max_sync(FULL, cnt0) + 1) & ~1;  // chain length of the warp's rows (even)
    int32_t rowT[NT], code[NT][NJ];
    double v[NT][NJ], xv[NT][NJ];
#pragma unroll
    for (int t = 0; t < NT; ++t) {  // the rows' extents travel by shuffle
      const int rr = t * 4 + sub;
      rowT[t] = __shfl_sync(FULL, r0, rr);
      const int cntT = __shfl_sync(FULL, cnt0, rr);
      const int64_t psT = __shfl_sync(FULL, ps0, rr);
      const int32_t *cp = a.colval + (psT + hl);  // this lane's first entry of the row; the others are 8, 16, 24 further
      const double *vp = a.nzval + (psT + hl);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        // no predicate: slots past the row end read the entries behind it (PA_MAT_PAD entries of padding follow the
        // last row) and are discarded below.  (A predicated load makes ptxas put a select right behind the load,
        // which waits for it.)
        const int32_t c = gs_ld_stream(cp + 8 * j, pol);
        v[t][j] = gs_ld_stream(vp + 8 * j, pol);
        code[t][j] = hl + 8 * j < cntT ? c : -1;
      }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int32_t c = code[t][j], row = rowT[t];
        const bool valid = c >= 0;
        // NEW value needed: an own row that precedes this one in the sweep order
        const bool isnew = valid && (BACKWARD ? (c > row && c < (int32_t)a.n) : (c < row));
        // slot code: >= 0 column of a NEW value, -1 unused, -2 OLD value, -3 OLD value + diagonal, -4 diagonal of the
        // zero-guess sweep (not subtracted)
        int32_t cd;
        if (!valid) cd = -1;
        else if (isnew) cd = c;
        else if (ZERO) cd = c == row ? -4 : -1;
        else cd = c == row ? -3 : -2;
        xv[t][j] = gs_ld_old(a.x + (cd == -2 || cd == -3 ? c : 0));  // slots without an OLD value read x[0]: harmless
        code[t][j] = cd;
      }
    }
    if (tr) tk[1] = clock64();
    // ---- the gate
    if (glev >= 0 && glev < la.nlev) {
      if (threadIdx.x == 0) {
        gs_wait_level(la.done + glev, la.sweep * (unsigned long long)la.lev_nb[glev], a.err);
        if (la.sync_mode == 0) gs_fence_acq_rel();  // acquire: relaxed poll + fence; the barrier extends it to the CTA
      }
      __syncthreads();
    }
    if (tr) tk[2] = clock64();
    // ---- behind the gate: the NEW values (what x holds now: rows of later levels have not started), then the products
#pragma unroll
    for (int t = 0; t < NT; ++t) {
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (code[t][j] >= 0) xv[t][j] = gs_ld_old(a.x + code[t][j]);  // through L2: x changes during the sweep
    }
    if (tr) tk[3] = clock64();
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int rr = t * 4 + sub;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int32_t cd = code[t][j];
        // unused entries contribute +0.0: s - (+0.0) == s bit for bit
        prod[rr * STRIDE + hl + 8 * j] = (cd >= 0 || cd == -2 || cd == -3) ? __dmul_rn(v[t][j], xv[t][j]) : 0.0;
        if (cd == -3 || cd == -4) dg[rr] = v[t][j];
      }
    }
    __syncwarp();
    if (tr) tk[4] = clock64();
    // ---- every lane finishes its own row
    {
      double s = b0;
      const double2 *pp = reinterpret_cast<const double2 *>(prod + lane * STRIDE);
#pragma unroll 4
      for (int k = 0; k < kend; k += 2) {  // s -= a*x[col], in CSR order
        const double2 pk = pp[k >> 1];
        s = __dsub_rn(__dsub_rn(s, pk.x), pk.y);
      }
      const double d = dg[lane];
      dg[lane] = 0.0;
      if (!ZERO) s = __dadd_rn(s, __dmul_rn(d, xo0));  // s += d*x[row]
      s = __ddiv_rn(s, d);
      if (r0 >= 0) a.x[r0] = s;
    }
    // ---- publish: CTA barrier, then one thread releases (fence + relaxed increment): the release is cumulative over
    // the stores the barrier ordered before it.  (__threadfence() by every thread = MEMBAR.SC + L1 invalidation in
    // all 8 warps was measured at ~3 us per level.)
    if (tr) tk[5] = clock64();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (la.sync_mode <= 1) gs_fence_acq_rel();
      asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" ::"l"(la.done + lev) : "memory");
    }
    if (tr) {
      tk[6] = clock64();
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      printf("gs-trace bt %lld lev %d: loads-issued %lld gate %lld xloads-issued %lld products %lld chain %lld publish %lld | globaltimer %llu\n",
             (long long)bt, lev, tk[1] - tk[0], tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3], tk[5] - tk[4], tk[6] - tk[5], gt);
    }
    // ---- rotate the pipeline
    r0 = r1; ps0 = ps1; cnt0 = cnt1; b0 = b1; xo0 = xo1;
    r1 = r2; ps1 = ps2; cnt1 = cnt2; b1 = b2; xo1 = xo2;
    r2 = r3;
  }
}


// ------------------------------------------------------------------ the SELL sweep kernel (rows of <= 32 entries)
// One THREAD per row, one warp per 32-row slice of the sweep-ordered SELL copy (GsOrder).  Every entry slot of the slice
// is one coalesced warp load; the x values are gathered by the lane that owns the row; the ordered chain
// s -= a*x[col] runs on all 32 lanes at once (the dataflow kernel k_gs_flow_pipe runs it on 4 of 32 lanes and spends
// ~110 warp instructions per row; here it is ~5).  SYNC: rows of earlier levels are awaited through their published
// (value, sweep epoch) pair — the per-row dataflow of the sequential sweep's dependency DAG, one launch per sweep.
// !SYNC: one launch per level (multi-colour order: 8 colours), no waiting at all: every read of x sees either a value of
// an earlier launch (a colour already updated) or a value no row of this launch writes.
struct GsSellArgs {
  const int32_t *rows;
  const int32_t *cols;
  const double *vals;
  const double *b;
  double *x;
  ulonglong2 *xe;
  int *err;
  int64_t g0, g1;    // sweep positions [g0, g1) of this launch (slice = position, or ngroups-1-position when backward)
  int64_t ngroups;
  int W;             // entry slots per row (runtime copy of the template parameter; used when W == 0)
  int epoch, backward, zero_guess;
  // MODE 2: per-level completion counters
  const int32_t *slice_lev;
  const uint32_t *lev_n;
  unsigned long long *done;
  unsigned long long sweep;
  int nlev;
  const int64_t *lev_first;  // first slice of every level (k_gs_sell_flow: which sub-counter a slice counts into)
  int gate, trace;
  const uint32_t *lev_tot;   // MODE 2: slices per level, one counter per level
  unsigned long long *done2;
  // colour kernel with row patterns
  const unsigned char *pat;
  const int32_t *ptab;       // the launch's level
  int npat;
};

// minBlocksPerSM is explicit: with maxThreads alone ptxas aims at full occupancy and sinks every load next to its use,
// which serialises the gathers of a batch (the same effect as in k_spmv_tma)
// MODE 0: no waiting (one launch per level/colour)   MODE 1: per-row dataflow (published (value, epoch) pairs)
// MODE 2: level gate — a slice of level L starts when the counter of level L-1 (L+1 backward) says all its slices are
//         stored; one lane per warp polls one hot line, the other lanes cost nothing while they wait, and the NEW values
//         are then simply what x holds (no per-row flags, no 16-byte pairs)
template <int W, int MODE>
__global__ void __launch_bounds__(GS_THREADS, (MODE == 1 ? 2 : (MODE == 0 ? GS_COLOR_MINB : 3))) k_gs_sell(const GsSellArgs a) {
  constexpr bool SYNC = MODE == 1;
  constexpr int B = W == 27 ? 9 : (W == 7 ? 7 : 8);  // slots per batch: loads of a batch are in flight together
  const int WD = W ? W : a.W;
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * (GS_THREADS / 32);
  const int64_t w = (int64_t)blockIdx.x * (GS_THREADS / 32) + (threadIdx.x >> 5);
  const unsigned epoch = (unsigned)a.epoch;
  const uint64_t pol = gs_stream_policy(), keep = gs_keep_policy(1);
  for (int64_t i = a.g0 + w; i < a.g1; i += nw) {
    const int64_t g = a.backward ? a.ngroups - 1 - i : i;
    const int32_t row = a.rows[g * 32 + lane];
    const int32_t *cp = a.cols + g * WD * 32 + lane;
    const double *vp = a.vals + g * WD * 32 + lane;
    {  // the next slice of this warp: start its HBM reads now (WD*256 + WD*128 bytes = 3*WD 128-byte lines)
      const int64_t inext = i + nw;
      if (inext < a.g1) {
        const int64_t gn = a.backward ? a.ngroups - 1 - inext : inext;
        const char *nv = reinterpret_cast<const char *>(a.vals + gn * WD * 32), *nc = reinterpret_cast<const char *>(a.cols + gn * WD * 32);
        for (int l = lane; l < 2 * WD; l += 32) gs_prefetch_l2(nv + (size_t)l * 128);
        if (lane < WD) gs_prefetch_l2(nc + (size_t)lane * 128);
      }
    }
    if (MODE != 2 && row < 0) continue;  // padding row (levels are padded to whole slices); no warp-level primitive below
    double s = row >= 0 ? __ldg(a.b + row) : 0.0, d = 0.0;
    const double xold = (a.zero_guess || row < 0) ? 0.0 : gs_ld_x(a.x + row, keep);  // nobody writes x[row] before this row does
    int lev = 0;
    if (MODE == 2) {
      lev = a.slice_lev[g];
      const int glev = a.backward ? lev + 1 : lev - 1;  // the level whose completion opens this one (and, by induction, all before it)
      if (glev >= 0 && glev < a.nlev) {
        if (lane == 0) {
          const unsigned long long target = a.sweep * (unsigned long long)a.lev_tot[glev];
          long long t0 = 0;
          while (gs_ld_relaxed(a.done2 + glev) < target) {
            if (!t0) {
              t0 = clock64();
            } else if (clock64() - t0 > GS_SPIN_LIMIT) {
              *a.err = 3;
              break;
            }
            __nanosleep(20);
          }
          gs_fence_acq_rel();  // acquire: relaxed polls + one fence; __syncwarp extends it to the warp
        }
        __syncwarp();
      }
    }
    if (row >= 0) {
#pragma unroll 1
    for (int k0 = 0; k0 < WD; k0 += B) {
      int32_t code[B];
      double v[B], xv[B];
      unsigned long long w0[B], w1[B];
      bool fresh[B], use[B];
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const bool in = W ? (k0 + u < W) : (k0 + u < WD);
        code[u] = in ? gs_ld_stream(cp + (k0 + u) * 32, pol) : -1;
        v[u] = in ? gs_ld_stream(vp + (k0 + u) * 32, pol) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const bool valid = code[u] >= 0;
        const int32_t c = code[u] & GS_COL_MASK;
        const bool ff = valid && (code[u] & GS_COL_FRESH), own = valid && (code[u] & GS_COL_OWN);
        // a value of THIS sweep is needed: forward: an own column of a lower level; backward: an own column of a higher one
        fresh[u] = SYNC && (a.backward ? (own && !ff && c != row) : ff);
        use[u] = valid && (!a.zero_guess || ff);
        xv[u] = 0.0;
        w0[u] = w1[u] = 0ull;
        if (fresh[u]) {
          asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(w0[u]), "=l"(w1[u]) : "l"(a.xe + c), "l"(keep) : "memory");
        } else if (use[u]) {
          xv[u] = MODE != 0 ? gs_ld_x(a.x + c, keep) : a.x[c];  // x changes during a one-launch sweep: through L2
        }
      }
#pragma unroll
      for (int u = 0; u < B; ++u) {
        if (fresh[u]) {
          const int32_t c = code[u] & GS_COL_MASK;
          long long t0 = 0;
          while ((unsigned)(w0[u] >> 32) != epoch || (unsigned)(w1[u] >> 32) != epoch) {
            asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(w0[u]), "=l"(w1[u]) : "l"(a.xe + c), "l"(keep) : "memory");
            if (!t0) {
              t0 = clock64();
            } else if (clock64() - t0 > GS_SPIN_LIMIT) {
              *a.err = 3;
              break;
            }
          }
          xv[u] = __longlong_as_double((long long)((w1[u] << 32) | (w0[u] & 0xffffffffull)));
        }
      }
#pragma unroll
      for (int u = 0; u < B; ++u) {  // s -= a*x[col], in CSR order
        const double t = __dsub_rn(s, __dmul_rn(v[u], xv[u]));
        s = use[u] ? t : s;
        if (code[u] >= 0 && (code[u] & GS_COL_MASK) == row) d = v[u];
      }
    }
    if (!a.zero_guess) s = __dadd_rn(s, __dmul_rn(d, xold));  // s += d*x[row]
    s = __ddiv_rn(s, d);
    if (SYNC) gs_publish(a.xe + row, a.x + row, s, epoch, keep);
    else a.x[row] = s;
    }  // row >= 0
    if (MODE == 2) {
      // publish the slice: the warp's stores are ordered before lane 0's release (syncwarp + cumulative fence)
      __syncwarp();
      if (lane == 0) {
        gs_fence_acq_rel();
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" ::"l"(a.done2 + lev) : "memory");
      }
    }
  }
}


// ------------------------------------------------------------------ the strip kernel of the lexicographic order
// The wavefront of a box operator, cut so that a WARP owns a piece of it for a long time.  A task = one strip of 32
// consecutive grid lines (y) of one plane (z); lane l walks line 32Y+l along x, w1 steps behind lane l-1 (27-pt: 2, so
// that (x+1, y-1) is finished when (x, y) starts; 7-pt: 1): step s of a task is the slice of rows (s - w1*l, 32Y+l, z),
// all on the wavefront level x + w1*y + w2*z = s + 32*w1*Y + w2*z — independent of each other.  The smoother's matrix is
// copied once in (task, step, lane) order (SELL-32: every entry slot of a step is one coalesced load, the stream of a
// warp is sequential in HBM).  Every row performs the reference's arithmetic on exactly the inputs of the sequential
// sweep (NEW values of lower-numbered own neighbours, OLD values of the others): bit-identical results.
//   * the NEW value of the row the lane finished one step ago (x-1) comes from a register;
//   * all other NEW values are read as published (value, sweep epoch) pairs: correctness never rests on a fence or on
//     the order in which warps run — a value is used only when both halves carry this sweep's epoch;
//   * OLD values are read through L1: a lane re-reads the same sectors of x for four steps in a row, and nobody
//     overwrites an OLD value before all its readers are done (anti-dependencies of a symmetric pattern follow the
//     true ones), so whatever an L1 line holds for a not-yet-updated row is the value wanted.
// Tasks are numbered by their start level 32*w1*Y + w2*z and dealt round robin to the resident warps; a step only
// depends on steps of smaller level, which belong to tasks at most (strips x 32*w1/w2) numbers ahead: as long as the grid
// holds that many warps (checked on the host) the lowest unfinished level can always proceed.
//   * sliding window: consecutive steps of a lane are consecutive rows of a grid line, so entry k of this step names the
//     column that entry k+1 (backward: k-1) of the previous step named — its value is taken from a per-warp table in shared
//     memory ([slot][lane], conflict free) instead of being gathered again: 9 instead of 26 gathers per row of the 27-pt
//     operator.  The test is on the column ids themselves (no assumption about the stencil): a match means "the same
//     column one row later", and a value that was NEW (OLD) for the previous row of the line is still NEW (OLD) for this
//     one, except the previous row itself, which is served from the register first.  The gathers of a wavefront slice go
//     to 32 different lines, and the L1 takes about one such request per cycle: the request count, not HBM, bounds this
//     sweep (profiles/), which is why the window matters.
#define GS_STRIP_SMEM(W) ((size_t)(GS_THREADS / 32) * (W) * 32 * 12)
template <int W, bool TRACE = false>
__global__ void __launch_bounds__(GS_THREADS, 2) k_gs_strip(const GsSellArgs a, const int ntask, const int64_t nsteps) {
  constexpr int B = W == 27 ? 9 : (W == 7 ? 7 : 8);
  extern __shared__ __align__(16) unsigned char gss_smem[];
  long long tk_load = 0, tk_gather = 0, tk_chain = 0, tk_rest = 0, tk0 = 0, tk1 = 0, tk_first = 0, tk_firstsum = 0, tk_old = 0;
  int tk_npoll = 0, tk_nfresh = 0;
  const int WD = W ? W : a.W;
  const int lane = threadIdx.x & 31;
  // the window of this warp: value and column of every entry slot of the lane's previous row
  double *sx = reinterpret_cast<double *>(gss_smem) + (size_t)(threadIdx.x >> 5) * WD * 32 + lane;
  int32_t *sc = reinterpret_cast<int32_t *>(gss_smem + (size_t)(GS_THREADS / 32) * WD * 32 * 8) + (size_t)(threadIdx.x >> 5) * WD * 32 + lane;
  const int64_t nw = (int64_t)gridDim.x * (GS_THREADS / 32);
  const int64_t w = (int64_t)blockIdx.x * (GS_THREADS / 32) + (threadIdx.x >> 5);
  const unsigned epoch = (unsigned)a.epoch;
  const uint64_t pol = gs_stream_policy(), keep = gs_keep_policy(1);
  const int dk = a.backward ? -1 : 1;  // the previous row's slot that names the same column
  for (int64_t t = w; t < ntask; t += nw) {
    const int64_t tq = a.backward ? ntask - 1 - t : t;
    int32_t prev_row = -1;
    double prev_val = 0.0;
    for (int k = 0; k < WD; ++k) sc[k * 32] = -1;  // empty window
    const int64_t gstep = a.backward ? -1 : 1;
    int64_t g = tq * nsteps + (a.backward ? nsteps - 1 : 0);
    int32_t row = a.rows[g * 32 + lane];
    double bv = row >= 0 ? __ldg(a.b + row) : 0.0;
    for (int64_t st = 0; st < nsteps; ++st, g += gstep) {
      const int32_t *cp = a.cols + g * WD * 32 + lane;
      const double *vp = a.vals + g * WD * 32 + lane;
      int32_t row_n = -1;
      if (st + 1 < nsteps) {  // the next step of this task: its row ids now, its matrix slice on the way to L2
        const int64_t gn = g + gstep;
        row_n = __ldg(a.rows + gn * 32 + lane);
        // ONE bulk prefetch per array (cp.async.bulk.prefetch.L2): per-line prefetch instructions (81 lines per slice, issued as
        // divergent LSU requests) queue up in front of the gathers — the first pair came back after ~2400 cycles with them
        if (a.gate == 1) {
          if (lane == 0) {
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.vals + gn * WD * 32), "r"(WD * 32 * 8) : "memory");
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.cols + gn * WD * 32), "r"(WD * 32 * 4) : "memory");
            if (st + 2 < nsteps) gs_prefetch_l2(a.rows + (gn + gstep) * 32);
          }
        } else if (a.gate == 2) {
          const char *nv = reinterpret_cast<const char *>(a.vals + gn * WD * 32), *nc = reinterpret_cast<const char *>(a.cols + gn * WD * 32);
          for (int l = lane; l < 2 * WD; l += 32) gs_prefetch_l2(nv + (size_t)l * 128);
          if (lane < WD) gs_prefetch_l2(nc + (size_t)lane * 128);
          if (lane == 0 && st + 2 < nsteps) gs_prefetch_l2(a.rows + (gn + gstep) * 32);
        }
      }
      if (TRACE) tk0 = clock64();
      if (row >= 0) {
        double s = bv, d = 0.0, xold = 0.0;
        int32_t carry_c = -1;  // backward: the slot below the batch, saved before the batch before it overwrote it
        double carry_x = 0.0;
#pragma unroll 1
        for (int k0 = 0; k0 < WD; k0 += B) {
          int32_t code[B], pc[B];
          double v[B], xv[B], px[B];
          unsigned long long w0[B], w1[B];
          bool fresh[B], use[B];
          if (TRACE) tk1 = clock64();
#pragma unroll
          for (int u = 0; u < B; ++u) {
            const bool in = W ? (k0 + u < W) : (k0 + u < WD);
            code[u] = in ? gs_ld_stream(cp + (k0 + u) * 32, pol) : -1;
            v[u] = in ? gs_ld_stream(vp + (k0 + u) * 32, pol) : 0.0;
          }
          // the window slots this batch compares with, read before the batch overwrites its own slots
#pragma unroll
          for (int u = 0; u < B; ++u) {
            const int kp = k0 + u + dk;
            const bool in = kp >= 0 && kp < WD && (dk > 0 || u > 0);
            pc[u] = in ? sc[kp * 32] : -1;
            px[u] = in ? sx[kp * 32] : 0.0;
          }
          if (dk < 0) {
            pc[0] = carry_c;
            px[0] = carry_x;
            const int kl = min(k0 + B - 1, WD - 1);
            carry_c = sc[kl * 32];
            carry_x = sx[kl * 32];
          }
          if (TRACE) {
#pragma unroll
            for (int u = 0; u < B; ++u) asm volatile("" ::"r"(code[u]), "d"(v[u]));
            const long long t2 = clock64();
            tk_load += t2 - tk1;
            tk1 = t2;
          }
#pragma unroll
          for (int u = 0; u < B; ++u) {
            const bool valid = code[u] >= 0;
            const int32_t c = valid ? (code[u] & GS_COL_MASK) : -1;
            const bool ff = valid && (code[u] & GS_COL_FRESH), own = valid && (code[u] & GS_COL_OWN);
            const bool nw_ = a.backward ? (own && !ff && c != row) : ff;  // a value of THIS sweep is needed
            use[u] = valid && (!a.zero_guess || ff);
            const bool mine = valid && c == prev_row, hit = valid && c == pc[u];
            const bool reg = mine || hit;
            xv[u] = mine ? prev_val : (hit ? px[u] : 0.0);
            fresh[u] = nw_ && !reg;
            w0[u] = w1[u] = 0ull;
            // Both loads are PREDICATED instructions, not branches: the destination registers are initialised above and keep
            // their value in the lanes that do not load.  (With `if (fresh) load; else w = 0` the compiler zeroes the registers
            // in the else path AFTER the load was issued for the other lanes: a write-after-write on the warp-wide scoreboard,
            // i.e. one full memory round trip per entry slot whenever the lanes of a slice disagree — rows on a box face
            // have fewer entries than their neighbours, so some lane always does.)
            const int32_t cs = valid ? c : 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t@p ld.relaxed.gpu.global.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;\n\t}"
                         : "+l"(w0[u]), "+l"(w1[u]) : "l"(a.xe + cs), "l"(keep), "r"((unsigned)fresh[u]) : "memory");
            const unsigned ld_old = use[u] && !nw_ && !reg;  // OLD value: L1 allowed
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.global.f64 %0, [%1];\n\t}" : "+d"(xv[u]) : "l"(a.x + cs), "r"(ld_old) : "memory");
          }
#pragma unroll
          for (int u = 0; u < B; ++u) {
            if (fresh[u]) {
              const int32_t c = code[u] & GS_COL_MASK;
              long long t0 = 0;
              if (TRACE) {
                if (!tk_first) {
                  asm volatile("" ::"l"(w0[u]));
                  tk_first = clock64() - tk1;
                }
                tk_nfresh++;
              }
              while ((unsigned)(w0[u] >> 32) != epoch || (unsigned)(w1[u] >> 32) != epoch) {
                if (TRACE) tk_npoll++;
                asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(w0[u]), "=l"(w1[u]) : "l"(a.xe + c), "l"(keep) : "memory");
                if (!t0) {
                  t0 = clock64();
                } else if (clock64() - t0 > GS_SPIN_LIMIT) {
                  *a.err = 3;
                  break;
                }
              }
              xv[u] = __longlong_as_double((long long)((w1[u] << 32) | (w0[u] & 0xffffffffull)));
            }
          }
          if (TRACE) {
            const long long t3 = clock64();
#pragma unroll
            for (int u = 0; u < B; ++u) asm volatile("" ::"d"(xv[u]));
            const long long t2 = clock64();
            tk_gather += t2 - tk1;
            tk_old += t2 - t3;
            tk_firstsum += tk_first;
            tk_first = 0;
            tk1 = t2;
          }
#pragma unroll
          for (int u = 0; u < B; ++u) {  // the window of the next row, then s -= a*x[col] in CSR order
            if (W ? (k0 + u < W) : (k0 + u < WD)) {
              sc[(k0 + u) * 32] = code[u] >= 0 ? (code[u] & GS_COL_MASK) : -1;
              sx[(k0 + u) * 32] = xv[u];
            }
            const double t2 = __dsub_rn(s, __dmul_rn(v[u], xv[u]));
            s = use[u] ? t2 : s;
            const bool dg = code[u] >= 0 && (code[u] & GS_COL_MASK) == row;
            d = dg ? v[u] : d;
            xold = dg ? xv[u] : xold;
          }
          if (TRACE) {
            asm volatile("" ::"d"(s));
            tk_chain += clock64() - tk1;
          }
        }
        if (!a.zero_guess) s = __dadd_rn(s, __dmul_rn(d, xold));  // s += d*x[row]
        s = __ddiv_rn(s, d);
        gs_publish(a.xe + row, a.x + row, s, epoch, keep);
        prev_row = row;
        prev_val = s;
      }
      row = row_n;
      bv = row >= 0 ? __ldg(a.b + row) : 0.0;
      __syncwarp();
      if (TRACE) {
        tk_rest += clock64() - tk0;
        if ((st & 63) == 63 && lane == 8 && (t == 0 || t == 40 || t == 300)) {
          printf("gs-strip task %lld step %lld (64 steps, lane 8): load %lld gather+poll %lld (first pair back after %lld, waiting for OLD values after the polls %lld; %d pair loads, %d re-polls) chain %lld total %lld cycles per step\n", (long long)t, (long long)st,
                 tk_load / 64, tk_gather / 64, tk_firstsum / 64, tk_old / 64, tk_nfresh, tk_npoll, tk_chain / 64, tk_rest / 64);
          tk_load = tk_gather = tk_chain = tk_rest = tk_firstsum = tk_old = 0;
          tk_npoll = tk_nfresh = 0;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ the multi-colour kernel: SELL slices through a TMA ring
// One launch per colour, like k_gs_sell<W, 0>, but the matrix never passes through registers on its way in: the slices
// of a colour are contiguous in the SELL copy, so a persistent CTA streams tiles of NW slices (values + column words,
// NW*W*384 bytes) HBM -> shared memory with cp.async.bulk (TMA engine, L2 evict-first) into a ring of S stages, exactly
// like k_spmv_tma; NW consumer warps (one thread per row) read their slots from shared memory ([slot][lane]: conflict
// free), gather x and run the ordered chain.  Only the x gathers, b and the row ids go through the LSU.
// Rows of a colour never read a value another row of the colour writes, so x is read with plain loads and the slices are
// walked in ascending order in both sweep directions (the order inside a colour is immaterial: same bits either way).
// PAT: the column words of a slice (a third of the stream) are replaced by ONE BYTE per row, the id of the row's pattern
// (column - row and flags of every slot) in a per-colour table held in shared memory; rows without a table entry (id 255:
// ghost columns, rare box positions) read their column words from global memory.  Same words, same arithmetic.
template <int W, int B, bool PAT = false>
__global__ void __launch_bounds__(288, 2) k_gs_color_tma(const GsSellArgs a, const int NW, const int S) {
  extern __shared__ __align__(128) unsigned char gsc_smem[];
  const int WD = W ? W : a.W;
  const int tile_e = NW * WD * 32;  // entries per stage
  double *val_s = reinterpret_cast<double *>(gsc_smem);
  int32_t *col_s = reinterpret_cast<int32_t *>(val_s + (size_t)S * tile_e);  // PAT: pattern ids, NW*32 bytes per stage
  uint64_t *full = reinterpret_cast<uint64_t *>(col_s + (size_t)S * (PAT ? NW * 8 : tile_e));
  uint64_t *empty = full + S;
  int32_t *ptab_s = reinterpret_cast<int32_t *>(empty + S);
  const int tid = threadIdx.x, nthr_c = NW * 32;
  if (PAT)
    for (int i = tid; i < a.npat * WD; i += blockDim.x) ptab_s[i] = a.ptab[i];
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, nthr_c);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t nsl = a.g1 - a.g0, ntiles = (nsl + NW - 1) / NW;
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int64_t nloc = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
  if (tid >= nthr_c) {
    if (tid == nthr_c) {  // producer: one elected lane drives the ring
      const uint64_t pol = gs_stream_policy();
      int s = 0;
      uint32_t ph = 1;  // parity of the consumers' release of stage s, one lap behind
      for (int64_t j = 0; j < nloc; ++j, ++s) {
        if (s == S) { s = 0; ph ^= 1u; }
        if (j >= S) mbar_wait(empty + s, ph);
        const int64_t g = a.g0 + (first + j * stride) * NW;
        const uint32_t nsl_t = (uint32_t)min((int64_t)NW, a.g1 - g), ne = nsl_t * WD * 32;
        if (PAT) {
          mbar_expect_tx(full + s, ne * 8u + nsl_t * 32u);
          tma_load_1d(val_s + (size_t)s * tile_e, a.vals + g * WD * 32, ne * 8u, full + s, pol);
          tma_load_1d(reinterpret_cast<unsigned char *>(col_s) + (size_t)s * NW * 32, a.pat + g * 32, nsl_t * 32u, full + s, pol);
        } else {
          mbar_expect_tx(full + s, ne * 12u);
          tma_load_1d(val_s + (size_t)s * tile_e, a.vals + g * WD * 32, ne * 8u, full + s, pol);
          tma_load_1d(col_s + (size_t)s * tile_e, a.cols + g * WD * 32, ne * 4u, full + s, pol);
        }
      }
    }
    return;
  }
  const int lane = tid & 31, wrp = tid >> 5;
  // row ids two tiles ahead, right-hand side one tile ahead: none of the three dependent reads waits inside the loop
  auto slice_row = [&](int64_t j) -> int32_t {
    if (j >= nloc) return -1;
    const int64_t g = a.g0 + (first + j * stride) * NW + wrp;
    return g < a.g1 ? __ldg(a.rows + g * 32 + lane) : -1;
  };
  int32_t row = slice_row(0), row_n = slice_row(1);
  double bv = row >= 0 ? __ldg(a.b + row) : 0.0;
  int s = 0;
  uint32_t ph = 0;
  for (int64_t j = 0; j < nloc; ++j, ++s) {
    if (s == S) { s = 0; ph ^= 1u; }
    const int32_t row_nn = slice_row(j + 2);
    const double bv_n = row_n >= 0 ? __ldg(a.b + row_n) : 0.0;
    mbar_wait(full + s, ph);
    if (row >= 0) {
      const double *vs = val_s + (size_t)s * tile_e + wrp * WD * 32 + lane;
      const int32_t *cs = col_s + (size_t)s * tile_e + wrp * WD * 32 + lane;
      double sum = bv, d = 0.0, xo = 0.0;
      const unsigned pid = PAT ? (unsigned)reinterpret_cast<const unsigned char *>(col_s)[(size_t)s * NW * 32 + wrp * 32 + lane] : 0u;
      const bool esc = PAT && pid == 255u;
      const int32_t *tab = ptab_s + (esc ? 0u : pid) * (unsigned)WD;
      const int32_t *gcode = a.cols + (a.g0 + (first + j * stride) * NW + wrp) * WD * 32 + lane;  // this row's column words in HBM
#pragma unroll 1
      for (int k0 = 0; k0 < WD; k0 += B) {
        int32_t code[B];
        double v[B], xv[B];
#pragma unroll
        for (int u = 0; u < B; ++u) {
          const int kk = W ? min(k0 + u, W - 1) : min(k0 + u, WD - 1);  // past the end: a redundant read of the last slot
          if (PAT) {
            const int32_t pk = tab[kk];
            code[u] = pk < 0 ? -1 : (((row + (pk & GS_COL_MASK)) & GS_COL_MASK) | (pk & (GS_COL_FRESH | GS_COL_OWN)));
            if (esc || pk == -2) code[u] = __ldg(gcode + kk * 32);
          } else {
            code[u] = cs[kk * 32];
          }
          v[u] = vs[kk * 32];
        }
#pragma unroll
        for (int u = 0; u < B; ++u) {
          const int32_t c = code[u] >= 0 ? (code[u] & GS_COL_MASK) : row;  // empty slot: any valid address
          asm("ld.global.f64 %0, [%1];" : "=d"(xv[u]) : "l"(a.x + c));  // scheduled like __ldg, but never the read-only path
        }
#pragma unroll
        for (int u = 0; u < B; ++u) {  // s -= a*x[col], in CSR order
          const bool in = (W ? (k0 + u < W) : (k0 + u < WD)) && code[u] >= 0;
          const bool use = in && (!a.zero_guess || (code[u] & GS_COL_FRESH));
          const double t = __dsub_rn(sum, __dmul_rn(v[u], xv[u]));
          sum = use ? t : sum;
          const bool dg = in && (code[u] & GS_COL_MASK) == row;
          d = dg ? v[u] : d;
          xo = dg ? xv[u] : xo;
        }
      }
      mbar_arrive(empty + s);
      if (!a.zero_guess) sum = __dadd_rn(sum, __dmul_rn(d, xo));  // s += d*x[row]
      a.x[row] = __ddiv_rn(sum, d);
    } else {
      mbar_arrive(empty + s);
    }
    row = row_n; row_n = row_nn; bv = bv_n;
  }
}

// ------------------------------------------------------------------ the SELL dataflow kernel of the lexicographic order
// One thread per row, one warp per slice, like k_gs_sell, but with everything that does not depend on the rows of
// earlier levels done BEFORE the warp waits:
//   before the gate  the slice's entries are read (coalesced, L2-prefetched one slice ahead) and parked in shared memory
//                    ([slot][lane]: conflict-free); entries whose column is NOT updated earlier in this sweep (later
//                    levels, ghosts, the diagonal) are multiplied with x right away — nobody writes those x before this
//                    row publishes — and parked as products; the slots that need a NEW value are listed;
//   the gate         lane 0 polls the completion counter of the previous level with relaxed loads — NO fence: the counter
//                    only keeps 31 of 32 lanes (and all their sector requests) out of the polling; whether a NEW value has
//                    really arrived is decided by its (value, sweep epoch) pair, written as one 16-byte store;
//   behind the gate  the NEW values are gathered (pairs verified, re-read in the rare case the counter ran ahead of the
//                    data), multiplied and parked; then the ordered chain s -= product[k], k = 0..W-1 in CSR order, on
//                    all 32 lanes at once; s += d*x_old; s /= d; publish (pair + plain x); relaxed increment of the level.
// Critical path per level: gather of the NEW pairs + W dependent subtractions + divide + publish — no fence, no
// matrix load.  Same arithmetic and order as k_gs_flow (bit-identical sweeps, tests/test_gpu_hpcg_mg.py).
#define GSF_WARPS (GS_THREADS / 32)
#define GS_SUB 32  // completion sub-counters per level (slice s of a level counts into s & 31): 1172 increments of ONE address per
                   // level serialise in L2 (~27 cycles each = 17 us per level at 512^3); 32 addresses polled by the 32 lanes do not
#define GSF_WARP_BYTES(W) ((size_t)(W) * 32 * (8 + 4 + 1))

template <int W>
__global__ void __launch_bounds__(GS_THREADS, 2) k_gs_sell_flow(const GsSellArgs a) {
  extern __shared__ __align__(16) unsigned char gsf_smem[];
  constexpr int B = W == 27 ? 9 : (W == 7 ? 7 : 8);
  const int WD = W ? W : a.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *sval = reinterpret_cast<double *>(gsf_smem + (size_t)warp * GSF_WARP_BYTES(WD)) + lane;   // [slot][lane]
  int32_t *scode = reinterpret_cast<int32_t *>(gsf_smem + (size_t)warp * GSF_WARP_BYTES(WD) + (size_t)WD * 32 * 8) + lane;
  unsigned char *slist = gsf_smem + (size_t)warp * GSF_WARP_BYTES(WD) + (size_t)WD * 32 * 12 + lane;
  const int64_t nw = (int64_t)gridDim.x * GSF_WARPS;
  const int64_t w = (int64_t)blockIdx.x * GSF_WARPS + warp;
  const unsigned epoch = (unsigned)a.epoch;
  const uint64_t pol = gs_stream_policy(), keep = gs_keep_policy(1);
  for (int64_t i = a.g0 + w; i < a.g1; i += nw) {
    const int64_t g = a.backward ? a.ngroups - 1 - i : i;
    const int32_t row = a.rows[g * 32 + lane];
    const int32_t *cp = a.cols + g * WD * 32 + lane;
    const double *vp = a.vals + g * WD * 32 + lane;
    {  // the next slice of this warp: start its HBM reads now
      const int64_t inext = i + nw;
      if (inext < a.g1) {
        const int64_t gn = a.backward ? a.ngroups - 1 - inext : inext;
        const char *nv = reinterpret_cast<const char *>(a.vals + gn * WD * 32), *nc = reinterpret_cast<const char *>(a.cols + gn * WD * 32);
        for (int l = lane; l < 2 * WD; l += 32) gs_prefetch_l2(nv + (size_t)l * 128);
        if (lane < WD) gs_prefetch_l2(nc + (size_t)lane * 128);
      }
    }
    // ---- before the gate
    long long tpre = 0;
    if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) tpre = clock64();
    double bval = 0.0, xold = 0.0, d = 0.0;
    int nf = 0;
    if (row >= 0) {
      bval = __ldg(a.b + row);
      if (!a.zero_guess) xold = gs_ld_x(a.x + row, keep);
    }
#pragma unroll 1
    for (int k0 = 0; k0 < WD; k0 += B) {
      int32_t code[B];
      double v[B], xv[B];
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const bool in = (W ? (k0 + u < W) : (k0 + u < WD)) && row >= 0;
        code[u] = in ? gs_ld_stream(cp + (k0 + u) * 32, pol) : -1;
        v[u] = in ? gs_ld_stream(vp + (k0 + u) * 32, pol) : 0.0;
      }
      bool fresh[B], use[B];
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const bool valid = code[u] >= 0;
        const int32_t c = code[u] & GS_COL_MASK;
        const bool ff = valid && (code[u] & GS_COL_FRESH), own = valid && (code[u] & GS_COL_OWN);
        fresh[u] = a.backward ? (own && !ff && c != row) : ff;
        use[u] = valid && (!a.zero_guess || ff);
        xv[u] = (use[u] && !fresh[u]) ? gs_ld_x(a.x + c, keep) : 0.0;  // OLD values, ghosts, the diagonal: final until this row publishes
        if (valid && c == row) d = v[u];
      }
#pragma unroll
      for (int u = 0; u < B; ++u) {
        if (W ? (k0 + u < W) : (k0 + u < WD)) {
          const int k = k0 + u;
          if (fresh[u]) {
            sval[k * 32] = v[u];
            scode[k * 32] = code[u] & GS_COL_MASK;
            slist[nf * 32] = (unsigned char)k;
            ++nf;
          } else {
            sval[k * 32] = use[u] ? __dmul_rn(v[u], xv[u]) : 0.0;  // unused slots contribute +0.0: s - (+0.0) == s bit for bit
          }
        }
      }
    }
    // ---- the gate (relaxed counter, no fence: it only spares the polling; the pairs below carry the proof)
    {
      const int lev = a.slice_lev[g];
      const int glev = a.backward ? lev + 1 : lev - 1;
      if (a.gate && glev >= 0 && glev < a.nlev) {
        // lane l polls sub-counter l of the previous level (one coalesced 256-byte read per poll)
        const unsigned long long target = a.sweep * (unsigned long long)a.lev_n[glev * GS_SUB + lane];
        long long t0 = 0;
        while (gs_ld_relaxed(a.done + glev * GS_SUB + lane) < target) {
          if (!t0) {
            t0 = clock64();
          } else if (clock64() - t0 > GS_SPIN_LIMIT) {
            *a.err = 3;
            break;
          }
        }
        __syncwarp();
      }
      long long tg = 0;
      if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) tg = clock64();
      // ---- behind the gate: the NEW values
      for (int j0 = 0; j0 < nf; j0 += 7) {
        unsigned long long w0[7], w1[7];
        int kk[7];
#pragma unroll
        for (int u = 0; u < 7; ++u) {
          kk[u] = j0 + u < nf ? (int)slist[(j0 + u) * 32] : -1;
          w0[u] = w1[u] = 0ull;
          if (kk[u] >= 0)
            asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(w0[u]), "=l"(w1[u]) : "l"(a.xe + scode[kk[u] * 32]), "l"(keep) : "memory");
        }
#pragma unroll
        for (int u = 0; u < 7; ++u) {
          if (kk[u] >= 0) {
            long long t0 = 0;
            while ((unsigned)(w0[u] >> 32) != epoch || (unsigned)(w1[u] >> 32) != epoch) {
              asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(w0[u]), "=l"(w1[u]) : "l"(a.xe + scode[kk[u] * 32]), "l"(keep) : "memory");
              if (!t0) {
                t0 = clock64();
              } else if (clock64() - t0 > GS_SPIN_LIMIT) {
                *a.err = 3;
                break;
              }
            }
            const double xn = __longlong_as_double((long long)((w1[u] << 32) | (w0[u] & 0xffffffffull)));
            sval[kk[u] * 32] = __dmul_rn(sval[kk[u] * 32], xn);
          }
        }
      }
      // ---- the ordered chain, all lanes at once
      if (row >= 0) {
        double s = bval;
        if (W) {
#pragma unroll
          for (int k = 0; k < W; ++k) s = __dsub_rn(s, sval[k * 32]);  // s -= a*x[col], in CSR order
        } else {
          for (int k = 0; k < WD; ++k) s = __dsub_rn(s, sval[k * 32]);
        }
        if (!a.zero_guess) s = __dadd_rn(s, __dmul_rn(d, xold));  // s += d*x[row]
        s = __ddiv_rn(s, d);
        gs_publish(a.xe + row, a.x + row, s, epoch, keep);
      }
      __syncwarp();
      if (lane == 0 && a.gate) asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" ::"l"(a.done + lev * GS_SUB + (int)((g - a.lev_first[lev]) & (GS_SUB - 1))) : "memory");
      if (a.trace && blockIdx.x == 0 && threadIdx.x == 0 && i < a.g0 + 40 * nw)
        printf("gs-flow slice %lld lev %d nf %d: pre-gate+wait %lld post-gate %lld cycles\n", (long long)g, lev, nf, tg - tpre, clock64() - tg);
    }
  }
}

// slot k of row r = entry k of the CSR row, flags from the levels (own columns number the own rows: square own block)
template <typename PtrT>
__global__ void k_gs_build_sell(const PtrT *rowptr, const int32_t *colval, const double *nzval, const int32_t *rows, int64_t npad, int W,
                                int64_t n_own, const int32_t *lev, int32_t *cols, double *vals) {
  for (int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pos < npad; pos += (int64_t)gridDim.x * blockDim.x) {
    const int64_t base = (pos >> 5) * W * 32 + (pos & 31);
    const int32_t row = rows[pos];
    int64_t p0 = 0;
    int len = 0, lr = 0;
    if (row >= 0) {
      p0 = (int64_t)rowptr[row];
      len = (int)((int64_t)rowptr[row + 1] - p0);
      lr = lev[row];
    }
    for (int k = 0; k < W; ++k) {
      int32_t code = -1;
      double v = 0.0;
      if (k < len) {
        const int32_t c = colval[p0 + k];
        const bool own = c < n_own;
        const bool fr = own && lev[c] < lr;
        code = c | (own ? GS_COL_OWN : 0) | (fr ? GS_COL_FRESH : 0);
        v = nzval[p0 + k];
      }
      cols[base + (int64_t)k * 32] = code;
      vals[base + (int64_t)k * 32] = v;
    }
  }
}

// multi-colour order on a box: 27-pt -> 8 colours (parity of x, y, z), 7-pt -> red/black
__global__ void k_levels_color(int32_t *lev, int32_t *rows, int64_t n, int64_t bx, int64_t by, int kind) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ix = i % bx, iy = (i / bx) % by, iz = i / (bx * by);
    lev[i] = kind == 27 ? (int32_t)((ix & 1) + 2 * (iy & 1) + 4 * (iz & 1)) : (int32_t)((ix + iy + iz) & 1);
    rows[i] = (int32_t)i;
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

static void gs_free_order(GsOrder *&o) {
  if (!o) return;
  cudaFree(o->d_rows);
  cudaFree(o->d_cols);
  cudaFree(o->d_vals);
  cudaFree(o->d_slice_lev);
  cudaFree(o->d_lev_n);
  cudaFree(o->d_done);
  cudaFree(o->d_lev_first);
  cudaFree(o->d_lev_tot);
  cudaFree(o->d_done2);
  cudaFree(o->d_pat);
  cudaFree(o->d_ptab);
  delete o;
  o = nullptr;
}

static void gs_free_part(GsPart &g) {
  gs_free_order(g.ord[0]);
  gs_free_order(g.ord[1]);
  cudaFree(g.d_rows);
  cudaFree(g.d_xe);
  cudaFree(g.d_rows_b);
  cudaFree(g.d_batch_lev);
  cudaFree(g.d_lev_nb);
  cudaFree(g.d_done);
  g = GsPart();
}

extern "C" int pa_gs_create(pa_mat *A, pa_gs **out) {
  PA_CHECK(A && out && A->committed, PA_ESTATE, "pa_gs_create: matrix missing or not committed");
  PA_CHECK(!A->subassembled, PA_EINVAL, "pa_gs_create: Gauss-Seidel needs an assembled matrix");
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
  p.kind = kind;
  for (int d = 0; d < 3; ++d) p.dims[d] = dims[d];
  p.w[0] = 1;
  p.w[1] = kind == 27 ? 2 : 1;
  p.w[2] = kind == 27 ? 4 : 1;
  return PA_OK;
}

// level-sorted rows -> padded list (every level starts at a multiple of 32) -> SELL copy with the order's flags
static int gs_build_sell(pa_ctx *c, const MatPart &m, GsPart &p, const int32_t *d_lev, const int32_t *d_lev_sorted, const int32_t *d_rows_sorted,
                         const std::vector<int> &cnt, int nlev, GsOrder **out) {
  GsOrder *o = new GsOrder();
  o->nlev = nlev;
  o->W = std::max(p.maxlen, 1);
  std::vector<int32_t> shift(nlev);
  o->lev_group.assign(nlev + 1, 0);
  int64_t at = 0, padded = 0;
  for (int l = 0; l < nlev; ++l) {
    padded = (padded + 31) / 32 * 32;
    o->lev_group[l] = padded / 32;
    if (!(padded - at < (1ll << 31) && padded + cnt[l] < (1ll << 31))) {
      delete o;
      pa_set_error("pa_gs: too many rows for the sweep-ordered copy");
      return PA_EINVAL;
    }
    shift[l] = (int32_t)(padded - at);
    at += cnt[l];
    padded += cnt[l];
  }
  o->npad = (padded + 31) / 32 * 32;
  o->lev_group[nlev] = o->npad / 32;
  int32_t *d_shift = nullptr;
  cudaError_t e = cudaMalloc((void **)&d_shift, nlev * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&o->d_rows, o->npad * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&o->d_cols, (size_t)o->npad * o->W * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&o->d_vals, (size_t)o->npad * o->W * sizeof(double));
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaFree(d_shift);
    gs_free_order(o);
    pa_set_error("pa_gs: cannot allocate the sweep-ordered matrix copy (%.2f GiB): %s", (double)p.n * o->W * 12 / 1073741824.0, cudaGetErrorString(e));
    return PA_ENOMEM;
  }
  PA_CUDA(cudaMemcpyAsync(d_shift, shift.data(), nlev * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  PA_CUDA(cudaMemsetAsync(o->d_rows, 0xff, o->npad * sizeof(int32_t), c->stream));
  k_pad_levels<<<148 * 8, 256, 0, c->stream>>>(d_lev_sorted, d_rows_sorted, d_shift, p.n, o->d_rows);
  if (m.ptr64)
    k_gs_build_sell<int64_t><<<148 * 8, 256, 0, c->stream>>>((const int64_t *)m.d_rowptr, m.d_colval, m.d_nzval, o->d_rows, o->npad, o->W, p.n, d_lev, o->d_cols, o->d_vals);
  else
    k_gs_build_sell<int32_t><<<148 * 8, 256, 0, c->stream>>>((const int32_t *)m.d_rowptr, m.d_colval, m.d_nzval, o->d_rows, o->npad, o->W, p.n, d_lev, o->d_cols, o->d_vals);
  PA_CUDA(cudaGetLastError());
  {  // level of every slice, slices per level, completion counters (level-gated sweeps)
    const int64_t ng = o->npad / 32;
    std::vector<int32_t> sl((size_t)ng);
    std::vector<uint32_t> ln((size_t)nlev * GS_SUB, 0u), lt((size_t)nlev);
    for (int l = 0; l < nlev; ++l) {
      lt[l] = (uint32_t)(o->lev_group[l + 1] - o->lev_group[l]);
      for (int64_t q = o->lev_group[l]; q < o->lev_group[l + 1]; ++q) {
        sl[(size_t)q] = l;
        ln[(size_t)l * GS_SUB + (size_t)((q - o->lev_group[l]) & (GS_SUB - 1))]++;
      }
    }
    PA_CUDA(cudaMalloc((void **)&o->d_slice_lev, std::max<int64_t>(ng, 1) * sizeof(int32_t)));
    PA_CUDA(cudaMalloc((void **)&o->d_lev_n, ln.size() * sizeof(uint32_t)));
    PA_CUDA(cudaMalloc((void **)&o->d_done, ln.size() * sizeof(unsigned long long)));
    PA_CUDA(cudaMalloc((void **)&o->d_lev_first, (size_t)(nlev + 1) * sizeof(int64_t)));
    PA_CUDA(cudaMalloc((void **)&o->d_lev_tot, (size_t)nlev * sizeof(uint32_t)));
    PA_CUDA(cudaMalloc((void **)&o->d_done2, (size_t)nlev * sizeof(unsigned long long)));
    PA_CUDA(cudaMemcpyAsync(o->d_slice_lev, sl.data(), (size_t)ng * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    PA_CUDA(cudaMemcpyAsync(o->d_lev_n, ln.data(), ln.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    PA_CUDA(cudaMemcpyAsync(o->d_lev_first, o->lev_group.data(), (size_t)(nlev + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    PA_CUDA(cudaMemcpyAsync(o->d_lev_tot, lt.data(), (size_t)nlev * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    PA_CUDA(cudaMemsetAsync(o->d_done, 0, ln.size() * sizeof(unsigned long long), c->stream));
    PA_CUDA(cudaMemsetAsync(o->d_done2, 0, (size_t)nlev * sizeof(unsigned long long), c->stream));
    PA_CUDA(cudaStreamSynchronize(c->stream));
  }
  PA_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_shift);
  c->launches += 2;
  *out = o;
  return PA_OK;
}

// strip order of a box: slice (task, step) holds rows (step - w1*lane, 32*Y + lane, z); -1 where that leaves the box
__global__ void k_rows_strip(int32_t *rows, int64_t nslices, int64_t nsteps, const int32_t *task_y, const int32_t *task_z, int64_t nx, int64_t ny, int w1) {
  for (int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pos < nslices * 32; pos += (int64_t)gridDim.x * blockDim.x) {
    const int lane = (int)(pos & 31);
    const int64_t sl = pos >> 5, task = sl / nsteps, st = sl % nsteps;
    const int64_t y = 32 * (int64_t)task_y[task] + lane, x = st - (int64_t)w1 * lane;
    rows[pos] = (y < ny && x >= 0 && x < nx) ? (int32_t)(x + nx * (y + ny * (int64_t)task_z[task])) : -1;
  }
}

static int gs_make_strip_order(pa_ctx *c, const MatPart &m, GsPart &p, const int32_t *d_lev, GsOrder **out) {
  const int64_t nx = p.dims[0], ny = p.dims[1], nz = p.dims[2];
  const int w1 = (int)p.w[1], w2 = (int)p.w[2];
  const int64_t nstrips = (ny + 31) / 32;
  GsOrder *o = new GsOrder();
  o->nlev = p.nlev;
  o->W = std::max(p.maxlen, 1);
  o->ntask = (int)(nstrips * nz);
  o->nsteps = nx + 31 * (int64_t)w1;
  std::vector<std::array<int64_t, 3>> key((size_t)o->ntask);  // (start level, strip, plane)
  for (int64_t z = 0; z < nz; ++z)
    for (int64_t Y = 0; Y < nstrips; ++Y) key[(size_t)(z * nstrips + Y)] = {32 * w1 * Y + w2 * z, Y, z};
  std::sort(key.begin(), key.end());
  std::vector<int32_t> ty((size_t)o->ntask), tz((size_t)o->ntask);
  for (int q = 0; q < o->ntask; ++q) {
    ty[q] = (int32_t)key[q][1];
    tz[q] = (int32_t)key[q][2];
  }
  const int64_t nslices = (int64_t)o->ntask * o->nsteps;
  o->npad = nslices * 32;
  int32_t *d_ty = nullptr, *d_tz = nullptr;
  cudaError_t e = cudaMalloc((void **)&d_ty, ty.size() * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_tz, tz.size() * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&o->d_rows, o->npad * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&o->d_cols, (size_t)o->npad * o->W * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc((void **)&o->d_vals, (size_t)o->npad * o->W * sizeof(double));
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaFree(d_ty); cudaFree(d_tz);
    gs_free_order(o);
    pa_set_error("pa_gs: cannot allocate the strip-ordered matrix copy (%.2f GiB): %s", (double)nslices * 32 * p.maxlen * 12 / 1073741824.0, cudaGetErrorString(e));
    return PA_ENOMEM;
  }
  PA_CUDA(cudaMemcpyAsync(d_ty, ty.data(), ty.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  PA_CUDA(cudaMemcpyAsync(d_tz, tz.data(), tz.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  k_rows_strip<<<148 * 8, 256, 0, c->stream>>>(o->d_rows, nslices, o->nsteps, d_ty, d_tz, nx, ny, w1);
  if (m.ptr64)
    k_gs_build_sell<int64_t><<<148 * 8, 256, 0, c->stream>>>((const int64_t *)m.d_rowptr, m.d_colval, m.d_nzval, o->d_rows, o->npad, o->W, p.n, d_lev, o->d_cols, o->d_vals);
  else
    k_gs_build_sell<int32_t><<<148 * 8, 256, 0, c->stream>>>((const int32_t *)m.d_rowptr, m.d_colval, m.d_nzval, o->d_rows, o->npad, o->W, p.n, d_lev, o->d_cols, o->d_vals);
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_ty); cudaFree(d_tz);
  c->launches += 2;
  *out = o;
  return PA_OK;
}

// tasks that can be co-dependent: strips x (32*w1/w2) start levels; the grid must hold that many warps (see k_gs_strip)
static bool gs_strip_ok(const GsPart &p, const MatPart &m, int64_t n_local_cols, int64_t resident_warps) {
  if (!p.geom || p.maxlen < 1 || p.maxlen > 32 || m.nnz == 0 || n_local_cols >= (1ll << 29)) return false;
  const int64_t nstrips = (p.dims[1] + 31) / 32;
  const int64_t nslices = nstrips * p.dims[2] * (p.dims[0] + 31 * p.w[1]);
  return nstrips * (32 * p.w[1] / std::max<int64_t>(p.w[2], 1) + 1) * 2 <= resident_warps && nslices * 32 < (1ll << 31) * 4;
}

static bool gs_sell_ok(const GsPart &p, const MatPart &m, int64_t n_local_cols) {
  return p.maxlen >= 1 && p.maxlen <= 32 && m.nnz > 0 && n_local_cols < (1ll << 29);
}

// ---- row patterns of a SELL copy, per level (colour)
// -2 = a ghost column (no OWN flag): ghost ids follow no pattern, the kernel reads that word from the SELL copy
__device__ __forceinline__ int32_t gs_pack_code(int32_t code, int32_t row) {
  if (code < 0) return -1;
  if (!(code & GS_COL_OWN)) return -2;
  return (((code & GS_COL_MASK) - row) & GS_COL_MASK) | (code & (GS_COL_FRESH | GS_COL_OWN));
}
__global__ void k_gs_pat_sample(const int32_t *rows, const int32_t *cols, int W, int64_t pos0, int64_t npos, int64_t nsample, int32_t *out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nsample; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t stride = npos / nsample > 0 ? npos / nsample : 1;
    int64_t q = i * stride + (int64_t)((unsigned long long)(i * 2654435761ull) % (unsigned long long)stride);
    if (q >= npos) q = npos - 1;
    const int64_t pos = pos0 + q;
    const int32_t row = rows[pos];
    int32_t *o = out + i * (1 + W);
    o[0] = row >= 0 ? 1 : 0;
    for (int k = 0; k < W; ++k) o[1 + k] = row >= 0 ? gs_pack_code(cols[((pos >> 5) * W + k) * 32 + (pos & 31)], row) : 0;
  }
}
__global__ void k_gs_pat_assign(const int32_t *rows, const int32_t *cols, int W, int64_t pos0, int64_t npos, const int32_t *ptab, int npat, unsigned char *pat,
                                unsigned long long *n_esc) {
  unsigned long long esc = 0;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < npos; q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pos = pos0 + q;
    const int32_t row = rows[pos];
    int id = 255;
    if (row >= 0) {
      const int32_t *cp = cols + ((pos >> 5) * W) * 32 + (pos & 31);
      const int32_t first = gs_pack_code(cp[0], row), last = gs_pack_code(cp[(W - 1) * 32], row);
      for (int t = 0; t < npat && id == 255; ++t) {
        if (ptab[t * W] != first || ptab[t * W + W - 1] != last) continue;
        bool same = true;
        for (int k = 1; k < W - 1; ++k) same &= ptab[t * W + k] == gs_pack_code(cp[k * 32], row);
        if (same) id = t;
      }
      esc += id == 255;
    }
    pat[pos] = (unsigned char)id;
  }
  if (esc) atomicAdd(n_esc, esc);
}

static int gs_build_patterns(pa_ctx *c, GsOrder *o, int64_t n_rows) {
  const int W = o->W, nlev = o->nlev;
  o->patterns = false;
  if (W < 2 || n_rows < pa_knob(c, "gs_pattern_min_rows", 4096)) return PA_OK;
  const int max_pat = std::min(254, (16 * 1024) / (W * 4));  // the level's table lives in shared memory
  o->lev_npat.assign(nlev, 0);
  o->lev_ptab_off.assign(nlev, 0);
  std::vector<int32_t> all;
  const int64_t nsample = 16384;
  int32_t *d_s = nullptr;
  PA_CUDA(cudaMalloc((void **)&d_s, nsample * (1 + W) * sizeof(int32_t)));
  std::vector<int32_t> hs((size_t)nsample * (1 + W));
  for (int l = 0; l < nlev; ++l) {
    const int64_t pos0 = o->lev_group[l] * 32, npos = (o->lev_group[l + 1] - o->lev_group[l]) * 32;
    o->lev_ptab_off[l] = (int64_t)all.size();
    if (npos == 0) continue;
    const int64_t ns = std::min<int64_t>(nsample, npos);
    k_gs_pat_sample<<<64, 256, 0, c->stream>>>(o->d_rows, o->d_cols, W, pos0, npos, ns, d_s);
    PA_CUDA(cudaMemcpyAsync(hs.data(), d_s, (size_t)ns * (1 + W) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    PA_CUDA(cudaStreamSynchronize(c->stream));
    std::map<std::vector<int32_t>, int64_t> freq;
    for (int64_t i = 0; i < ns; ++i) {
      const int32_t *q = hs.data() + i * (1 + W);
      if (q[0]) freq[std::vector<int32_t>(q + 1, q + 1 + W)]++;
    }
    const int64_t min_count = std::max<int64_t>(1, ns / 4096);  // one-off tuples (ghost columns) keep their column words
    std::vector<std::pair<int64_t, std::vector<int32_t>>> order;
    for (auto &kv : freq)
      if (kv.second >= min_count) order.emplace_back(-kv.second, kv.first);
    std::sort(order.begin(), order.end());
    if ((int)order.size() > max_pat) order.resize(max_pat);
    o->lev_npat[l] = (int)order.size();
    for (auto &e : order) all.insert(all.end(), e.second.begin(), e.second.end());
  }
  cudaFree(d_s);
  if (all.empty()) return PA_OK;
  unsigned long long *d_esc = nullptr, h_esc = 0;
  PA_CUDA(cudaMalloc((void **)&o->d_ptab, all.size() * sizeof(int32_t)));
  PA_CUDA(cudaMalloc((void **)&o->d_pat, (size_t)o->npad + 512));
  PA_CUDA(cudaMalloc((void **)&d_esc, sizeof(unsigned long long)));
  PA_CUDA(cudaMemcpyAsync(o->d_ptab, all.data(), all.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  PA_CUDA(cudaMemsetAsync(o->d_pat, 255, (size_t)o->npad + 512, c->stream));
  PA_CUDA(cudaMemsetAsync(d_esc, 0, sizeof(unsigned long long), c->stream));
  for (int l = 0; l < nlev; ++l) {
    const int64_t pos0 = o->lev_group[l] * 32, npos = (o->lev_group[l + 1] - o->lev_group[l]) * 32;
    if (npos == 0 || o->lev_npat[l] == 0) continue;
    k_gs_pat_assign<<<148 * 8, 256, 0, c->stream>>>(o->d_rows, o->d_cols, W, pos0, npos, o->d_ptab + o->lev_ptab_off[l], o->lev_npat[l], o->d_pat, d_esc);
  }
  PA_CUDA(cudaMemcpyAsync(&h_esc, d_esc, sizeof(h_esc), cudaMemcpyDeviceToHost, c->stream));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_esc);
  c->launches += 2 * nlev;
  if ((double)h_esc > 0.08 * (double)n_rows) {  // too many rows without a pattern: keep the column words
    cudaFree(o->d_pat); cudaFree(o->d_ptab);
    o->d_pat = nullptr; o->d_ptab = nullptr;
    return PA_OK;
  }
  o->patterns = true;
  return PA_OK;
}

// the multi-colour order of a box operator (built on first use: it costs a second copy of the matrix)
static int gs_make_color_order(pa_gs *g, int k) {
  pa_ctx *c = g->A->ctx;
  GsPart &p = g->parts[k];
  const MatPart &m = g->A->parts[k];
  PA_CHECK(p.geom, PA_EINVAL, "multi-colour Gauss-Seidel needs the box geometry hint (pa_gs_set_box)");
  PA_CHECK(gs_sell_ok(p, m, g->A->cols->parts[k].n_local), PA_EINVAL, "multi-colour Gauss-Seidel needs rows of at most 32 entries");
  int32_t *d_lev = nullptr, *d_lev2 = nullptr, *d_rows0 = nullptr, *d_rows1 = nullptr;
  PA_CUDA(cudaMalloc((void **)&d_lev, p.n * sizeof(int32_t)));
  PA_CUDA(cudaMalloc((void **)&d_lev2, p.n * sizeof(int32_t)));
  PA_CUDA(cudaMalloc((void **)&d_rows0, p.n * sizeof(int32_t)));
  PA_CUDA(cudaMalloc((void **)&d_rows1, p.n * sizeof(int32_t)));
  const int nlev = p.kind == 27 ? 8 : 2;
  k_levels_color<<<148 * 8, 256, 0, c->stream>>>(d_lev, d_rows0, p.n, p.dims[0], p.dims[1], p.kind);
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_lev, d_lev2, d_rows0, d_rows1, (int)p.n, 0, 3, c->stream);
  void *d_tmp = nullptr;
  PA_CUDA(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
  PA_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_lev, d_lev2, d_rows0, d_rows1, (int)p.n, 0, 3, c->stream));  // stable: rows ascending within a colour
  int *d_cnt = nullptr;
  PA_CUDA(cudaMalloc((void **)&d_cnt, nlev * sizeof(int)));
  PA_CUDA(cudaMemsetAsync(d_cnt, 0, nlev * sizeof(int), c->stream));
  k_level_hist<<<148 * 8, 256, 0, c->stream>>>(d_lev, p.n, d_cnt);
  std::vector<int> cnt(nlev);
  PA_CUDA(cudaMemcpyAsync(cnt.data(), d_cnt, nlev * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  int rc = gs_build_sell(c, m, p, d_lev, d_lev2, d_rows1, cnt, nlev, &p.ord[1]);
  cudaFree(d_tmp); cudaFree(d_lev); cudaFree(d_lev2); cudaFree(d_rows0); cudaFree(d_rows1); cudaFree(d_cnt);
  c->launches += 3;
  if (rc == PA_OK && pa_knob(c, "gs_color_patterns", 1) != 0) rc = gs_build_patterns(c, p.ord[1], p.n);
  return rc;
}

/* Order of the sweeps: PA_GS_LEXICOGRAPHIC (default) = the reference's sequential order 1:n / n:-1:1 executed as a
 * wavefront dataflow — iterates bit-identical to the reference; PA_GS_MULTICOLOR = colour by colour (8 colours for the
 * 27-pt operator, red/black for 7-pt; needs pa_gs_set_box): a different but equally valid Gauss-Seidel order — same fixed
 * point, convergence-level parity (SURVEY 8f-1 allows it; gated by the reference's own HPCG test).  Switching releases the
 * matrix copy of the other order. */
extern "C" int pa_gs_set_order(pa_gs *g, int32_t order) {
  PA_CHECK(g && g->committed, PA_ESTATE, "pa_gs_set_order: smoother missing or not committed");
  PA_CHECK(order == PA_GS_LEXICOGRAPHIC || order == PA_GS_MULTICOLOR, PA_EINVAL, "pa_gs_set_order: unknown order %d", order);
  pa_ctx *c = g->A->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  if (order == g->order) return PA_OK;
  PA_CUDA(cudaStreamSynchronize(c->stream));
  if (order == PA_GS_MULTICOLOR) {
    for (int k = 0; k < c->nlocal; ++k) {
      if (g->parts[k].n == 0) continue;
      PA_CHECK(g->parts[k].geom, PA_EINVAL, "pa_gs_set_order: the multi-colour order needs pa_gs_set_box before pa_gs_commit");
    }
    for (int k = 0; k < c->nlocal; ++k) {
      if (g->parts[k].n == 0) continue;
      gs_free_order(g->parts[k].ord[0]);  // the wavefront copy is rebuilt if the caller switches back
      if (!g->parts[k].ord[1]) PA_TRY(gs_make_color_order(g, k));
    }
  } else {
    for (int k = 0; k < c->nlocal; ++k) gs_free_order(g->parts[k].ord[1]);
    // (the wavefront SELL copy is not rebuilt: the dataflow kernel on the CSR itself serves the lexicographic order)
  }
  g->order = order;
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
    // level histogram and the longest row
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
    // the level-sorted row list with every level starting at a multiple of `al` (-1 = padding)
    auto padded_list = [&](int64_t al, int32_t **d_list, int64_t *npad) -> int {
      std::vector<int32_t> shift(p.nlev);
      int64_t at = 0, padded = 0;
      for (int l = 0; l < p.nlev; ++l) {
        padded = (padded + al - 1) / al * al;
        PA_CHECK(padded - at < (1ll << 31) && padded + cnt[l] < (1ll << 31), PA_EINVAL, "pa_gs_commit: too many rows");
        shift[l] = (int32_t)(padded - at);
        at += cnt[l];
        padded += cnt[l];
      }
      PA_CHECK(at == p.n, PA_ESTATE, "pa_gs_commit: level histogram does not add up");
      *npad = (padded + al - 1) / al * al;
      int32_t *d_shift = nullptr;
      PA_CUDA(cudaMalloc((void **)&d_shift, p.nlev * sizeof(int32_t)));
      PA_CUDA(cudaMemcpyAsync(d_shift, shift.data(), p.nlev * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
      PA_CUDA(cudaMalloc((void **)d_list, *npad * sizeof(int32_t)));
      PA_CUDA(cudaMemsetAsync(*d_list, 0xff, *npad * sizeof(int32_t), c->stream));
      k_pad_levels<<<148 * 8, 256, 0, c->stream>>>(d_lev2, d_rows1, d_shift, p.n, *d_list);
      PA_CUDA(cudaStreamSynchronize(c->stream));
      cudaFree(d_shift);
      return PA_OK;
    };
    PA_TRY(padded_list(8, &p.d_rows, &p.npad));
    // gs_kernel = 2: sweep-ordered SELL copy of the wavefront order + thread-per-row kernels.  NOT the default: measured on
    // B200 (27-pt 512^3, symmetric sweep) 76-99 ms against 58 ms of the warp-per-row dataflow kernel on the CSR — with one
    // thread per row a lane collects its 13 NEW values itself (two dependent L2 round trips behind the gate, ~6500 cycles per
    // level, profiles/r02_gs_sell_trace.log) where the dataflow kernel spreads them over 8 lanes that are already polling.
    // The SELL copy is what the multi-colour order runs on (no waiting at all there: 27.8 ms).
    if (gs_sell_ok(p, m, g->A->cols->parts[k].n_local) && pa_knob(c, "gs_kernel", 0) == 2) {
      PA_TRY(gs_build_sell(c, m, p, d_lev, d_lev2, d_rows1, cnt, p.nlev, &p.ord[0]));
    }
    // gs_kernel = 3: strip order (a warp owns 32 grid lines of a plane) + k_gs_strip
    if (pa_knob(c, "gs_kernel", 0) == 3 && gs_strip_ok(p, m, g->A->cols->parts[k].n_local, 148 * 2 * (GS_THREADS / 32))) {
      PA_TRY(gs_make_strip_order(c, m, p, d_lev, &p.ord[0]));
    }
    // batch kernel tables (only when that kernel is selected: it is not the default): level l occupies
    // ceil(cnt_l / GSB_ROWS) consecutive batches
    if (p.maxlen <= 32 && m.nnz > 0 && pa_knob(c, "gs_kernel", 0) == 1) {
      const int64_t al = GSB_ROWS;
      PA_TRY(padded_list(al, &p.d_rows_b, &p.npad_b));
      std::vector<int32_t> batch_lev((size_t)(p.npad_b / al));
      std::vector<uint32_t> lev_nb(p.nlev);
      size_t bt = 0;
      for (int l = 0; l < p.nlev; ++l) {
        lev_nb[l] = (uint32_t)((cnt[l] + al - 1) / al);
        PA_CHECK(lev_nb[l] > 0, PA_ESTATE, "pa_gs_commit: empty wavefront level %d", l);
        for (uint32_t q = 0; q < lev_nb[l]; ++q) batch_lev[bt++] = l;
      }
      PA_CHECK(bt == batch_lev.size(), PA_ESTATE, "pa_gs_commit: batch table does not add up");
      PA_CUDA(cudaMalloc((void **)&p.d_batch_lev, batch_lev.size() * sizeof(int32_t)));
      PA_CUDA(cudaMalloc((void **)&p.d_lev_nb, lev_nb.size() * sizeof(uint32_t)));
      PA_CUDA(cudaMalloc((void **)&p.d_done, (size_t)p.nlev * sizeof(unsigned long long)));
      PA_CUDA(cudaMemcpyAsync(p.d_batch_lev, batch_lev.data(), batch_lev.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
      PA_CUDA(cudaMemcpyAsync(p.d_lev_nb, lev_nb.data(), lev_nb.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
      PA_CUDA(cudaMemsetAsync(p.d_done, 0, (size_t)p.nlev * sizeof(unsigned long long), c->stream));
      PA_CUDA(cudaStreamSynchronize(c->stream));
      p.sweeps = 0;
    }
    PA_CUDA(cudaMalloc((void **)&p.d_xe, p.n * sizeof(ulonglong2)));
    PA_CUDA(cudaMemsetAsync(p.d_xe, 0, p.n * sizeof(ulonglong2), c->stream));
    PA_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_tmp); cudaFree(d_lev); cudaFree(d_lev2); cudaFree(d_rows0); cudaFree(d_rows1); cudaFree(d_cnt);
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

// lanes per row of the dataflow kernel: 8 (four rows per warp) on the widest grids, 16 (two rows per warp) once the
// levels are wide enough to be throughput bound, 32 where the sweep is bound by the level-to-level hop; 0 = the
// unpipelined warp-per-row kernel (any row length).
// Measured on B200 (27-pt, symmetric sweep, ms):   rows      4 lanes   8 lanes   16 lanes   32 lanes
//                                                  134.2M     73.6      58.0      73.3        -     (8 lanes: 60.1 at 3 CTAs/SM)
//                                                   16.8M     21.8      12.5      11.3       13.5
//                                                    2.1M     10.4       5.8       5.0        2.2 (3.7 before the pipeline)
//                                                    262k      5.0       2.7        -         0.9
// (One THREAD per row was tried for the widest levels and is 2.7x slower: 205 ms vs 75 ms at 134M rows.)
static int gs_default_lanes(pa_ctx *c, const GsPart &p) {
  return (int)pa_knob(c, "gs_lanes", p.n >= (64ll << 20) ? 8 : (p.n >= (4ll << 20) ? 16 : 32));
}

// one sweep over the own rows of every local part (ghost entries of x are inputs only)
static int gs_sweep(pa_gs *g, pa_vec *x, const pa_vec *b, int backward, int zero_guess) {
  pa_ctx *c = g->A->ctx;
  for (int k = 0; k < c->nlocal; ++k) {
    GsPart &p = g->parts[k];
    const MatPart &m = g->A->parts[k];
    if (p.n == 0) continue;
    p.epoch += 1;
    const int64_t gsk = pa_knob(c, "gs_kernel", 0);
    GsOrder *o = g->order == PA_GS_MULTICOLOR ? p.ord[1] : ((gsk == 2 || gsk == 3) ? p.ord[0] : nullptr);
    PA_CHECK(g->order != PA_GS_MULTICOLOR || o, PA_ESTATE, "gs_sweep: the multi-colour copy of the matrix is missing");
    if (o) {
      PA_CHECK(!(zero_guess && backward), PA_ESTATE, "gs_sweep: the zero-guess sweep is a forward sweep");
      GsSellArgs a;
      a.rows = o->d_rows;
      a.cols = o->d_cols;
      a.vals = o->d_vals;
      a.b = b->d[k];
      a.x = x->d[k];
      a.xe = p.d_xe;
      a.err = c->d_err;
      a.ngroups = o->npad / 32;
      a.W = o->W;
      a.epoch = p.epoch;
      a.backward = backward;
      a.zero_guess = zero_guess;
      a.slice_lev = nullptr;
      a.lev_n = nullptr;
      a.done = nullptr;
      a.sweep = 0;
      a.nlev = o->nlev;
      a.lev_first = nullptr;
      a.gate = a.trace = 0;
      a.lev_tot = nullptr;
      a.done2 = nullptr;
      int nsm = 148;
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
      const int wpc = GS_THREADS / 32;
      if (g->order == PA_GS_MULTICOLOR) {
        // one launch per colour, in sweep order; within a colour the rows are independent
        void (*kern)(const GsSellArgs) = o->W == 27 ? k_gs_sell<27, 0> : (o->W == 7 ? k_gs_sell<7, 0> : k_gs_sell<0, 0>);
        // gs_color_kernel 1 (default): the slices stream through a TMA ring (k_gs_color_tma); 0: every lane loads its own entries
        const bool ring = pa_knob(c, "gs_color_kernel", 1) != 0;
        const bool pat0 = pa_knob(c, "gs_color_kernel", 1) != 0 && o->patterns && pa_knob(c, "gs_color_patterns", 1) != 0 && (o->W == 27 || o->W == 7);
        // measured at 27-pt 512^3 (symmetric sweep): column words 2 slices x 2 stages x 14: 18.7 ms; row patterns 4 x 2 x 27: 14.0 ms
        const int batch = (int)pa_knob(c, "gs_color_batch", pat0 ? 27 : 14);
        void (*tk)(const GsSellArgs, int, int) = o->W == 27 ? (batch >= 27 ? k_gs_color_tma<27, 27> : (batch >= 14 ? k_gs_color_tma<27, 14> : k_gs_color_tma<27, 9>))
                                                 : (o->W == 7 ? k_gs_color_tma<7, 7> : k_gs_color_tma<0, 8>);
        const bool pat = ring && o->patterns && pa_knob(c, "gs_color_patterns", 1) != 0 && (o->W == 27 || o->W == 7);
        if (pat) tk = o->W == 27 ? (batch >= 27 ? k_gs_color_tma<27, 27, true> : (batch >= 14 ? k_gs_color_tma<27, 14, true> : k_gs_color_tma<27, 9, true>)) : k_gs_color_tma<7, 7, true>;
        int NW = (int)pa_knob(c, "gs_color_slices", pat ? 4 : 2), S = (int)pa_knob(c, "gs_color_stages", 2);
        NW = std::max(1, std::min(NW, 8));
        S = std::max(2, std::min(S, 8));
        size_t smem = 0;
        int per_sm = 1;
        if (ring) {
          int max_npat = 0;
          if (pat) for (int l = 0; l < o->nlev; ++l) max_npat = std::max(max_npat, o->lev_npat[l]);
          while ((smem = (size_t)S * ((size_t)NW * o->W * 32 * (pat ? 8 : 12) + (pat ? NW * 32 : 0) + 16) + (size_t)max_npat * o->W * 4) > 200 * 1024 && (NW > 1 || S > 2)) {
            if (NW > 1) NW /= 2; else --S;
          }
          PA_CHECK(smem <= 220 * 1024, PA_ESTATE, "gs_sweep: a slice of the multi-colour copy does not fit shared memory");
          PA_CUDA(cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          PA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tk, NW * 32 + 32, smem));
          const int64_t cap = pa_knob(c, "gs_color_ctas", 0);
          if (cap > 0 && cap < per_sm) per_sm = (int)cap;
          if (per_sm < 1) per_sm = 1;
        }
        for (int q = 0; q < o->nlev; ++q) {
          const int l = backward ? o->nlev - 1 - q : q;
          const int64_t s0 = o->lev_group[l], s1 = o->lev_group[l + 1];
          if (s1 == s0) continue;
          if (ring) {
            a.g0 = s0;
            a.g1 = s1;
            a.pat = pat ? o->d_pat : nullptr;
            a.ptab = pat ? o->d_ptab + o->lev_ptab_off[l] : nullptr;
            a.npat = pat ? o->lev_npat[l] : 0;
            const int64_t grid = std::min<int64_t>((s1 - s0 + NW - 1) / NW, (int64_t)nsm * per_sm);
            tk<<<(unsigned)grid, NW * 32 + 32, smem, c->stream>>>(a, NW, S);
            c->launches++;
            continue;
          }
          a.g0 = backward ? a.ngroups - s1 : s0;
          a.g1 = backward ? a.ngroups - s0 : s1;
          const int64_t grid = std::min<int64_t>((s1 - s0 + wpc - 1) / wpc, (int64_t)nsm * 16);
          kern<<<(unsigned)grid, GS_THREADS, 0, c->stream>>>(a);
          c->launches++;
        }
      } else if (o->ntask) {
        void (*kern)(const GsSellArgs, int, int64_t) = o->W == 27 ? (pa_knob(c, "gs_trace", 0) ? k_gs_strip<27, true> : k_gs_strip<27>)
                                                                 : (o->W == 7 ? k_gs_strip<7> : k_gs_strip<0>);
        const size_t smem = GS_STRIP_SMEM(o->W);
        a.gate = (int)pa_knob(c, "gs_strip_prefetch", 1);  // 1 = bulk L2 prefetch of the next slice, 2 = per-line prefetches, 0 = none
        PA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int ctas_per_sm = 0;
        PA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, GS_THREADS, smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        const int64_t cap = pa_knob(c, "gs_ctas", 0);
        if (cap > 0 && cap < ctas_per_sm) ctas_per_sm = (int)cap;
        a.g0 = 0;
        a.g1 = a.ngroups;
        // all CTAs co-resident (see k_gs_strip): the grid never exceeds what the device holds at once
        const int64_t grid = std::min<int64_t>((o->ntask + wpc - 1) / wpc, (int64_t)nsm * ctas_per_sm);
        PA_CHECK(gs_strip_ok(p, m, g->A->cols->parts[k].n_local, grid * wpc) || grid * wpc >= o->ntask, PA_ESTATE, "gs_sweep: the strip kernel needs more resident warps than the device holds");
        kern<<<(unsigned)grid, GS_THREADS, smem, c->stream>>>(a, o->ntask, o->nsteps);
        c->launches++;
      } else {
        // gs_sell_mode 3 (default): staged dataflow kernel (relaxed level gate + per-row pairs); 2: fenced level gate; 1: per-row pairs only
        const int mode = (int)pa_knob(c, "gs_sell_mode", 3);
        void (*kern)(const GsSellArgs) = mode == 3 ? (o->W == 27 ? k_gs_sell_flow<27> : (o->W == 7 ? k_gs_sell_flow<7> : k_gs_sell_flow<0>))
                                       : mode == 2 ? (o->W == 27 ? k_gs_sell<27, 2> : (o->W == 7 ? k_gs_sell<7, 2> : k_gs_sell<0, 2>))
                                                   : (o->W == 27 ? k_gs_sell<27, 1> : (o->W == 7 ? k_gs_sell<7, 1> : k_gs_sell<0, 1>));
        a.slice_lev = o->d_slice_lev;
        a.lev_n = o->d_lev_n;
        a.done = o->d_done;
        a.gate = (int)pa_knob(c, "gs_gate", 1);
        a.trace = (int)pa_knob(c, "gs_trace", 0);
        a.lev_first = o->d_lev_first;
        a.lev_tot = o->d_lev_tot;
        a.done2 = o->d_done2;
        // every gated sweep advances every (sub-)counter by its slice count, once: the target is sweep number x count
        a.sweep = mode == 3 ? (a.gate ? ++o->sweeps : 0) : (mode == 2 ? ++o->sweeps2 : 0);
        a.nlev = o->nlev;
        const size_t smem = mode == 3 ? GSF_WARPS * GSF_WARP_BYTES(o->W) : 0;
        if (smem) PA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int ctas_per_sm = 0;
        PA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, GS_THREADS, smem));
        if (ctas_per_sm < 1) ctas_per_sm = 1;
        const int64_t cap = pa_knob(c, "gs_ctas", 0);
        if (cap > 0 && cap < ctas_per_sm) ctas_per_sm = (int)cap;
        a.g0 = 0;
        a.g1 = a.ngroups;
        // all CTAs co-resident: a waiting row only waits for rows of earlier slices, held by running warps
        const int64_t grid = std::min<int64_t>((a.ngroups + wpc - 1) / wpc, (int64_t)nsm * ctas_per_sm);
        kern<<<(unsigned)grid, GS_THREADS, smem, c->stream>>>(a);
        c->launches++;
      }
      continue;
    }
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
      a.prefetch = (int)pa_knob(c, "gs_prefetch", 1);
      a.keep = (int)pa_knob(c, "gs_keep", 1);
      int nsm = 148;
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
      // batch kernel (32 rows of one level per warp) where the levels are wide; the warp-per-row dataflow kernel where
      // the sweep is bound by the level-to-level hop (coarse grids) or rows are longer than 32 entries
      if (p.d_rows_b && pa_knob(c, "gs_kernel", 0) == 1) {
        const int nj = p.maxlen <= 8 ? 1 : (p.maxlen <= 16 ? 2 : 4);
        void (*wk)(const GsLevelArgs<PtrT>) = nullptr;
        PA_CHECK(!(zero_guess && backward), PA_ESTATE, "gs_sweep: the zero-guess sweep is a forward sweep");
        if (zero_guess)
          wk = nj == 1 ? k_gs_level<PtrT, 1, false, true> : nj == 2 ? k_gs_level<PtrT, 2, false, true> : k_gs_level<PtrT, 4, false, true>;
        else if (!backward)
          wk = nj == 1 ? k_gs_level<PtrT, 1, false, false> : nj == 2 ? k_gs_level<PtrT, 2, false, false> : k_gs_level<PtrT, 4, false, false>;
        else
          wk = nj == 1 ? k_gs_level<PtrT, 1, true, false> : nj == 2 ? k_gs_level<PtrT, 2, true, false> : k_gs_level<PtrT, 4, true, false>;
        GsLevelArgs<PtrT> la;
        la.g = a;
        la.g.rows = p.d_rows_b;
        la.g.npad = p.npad_b;
        la.batch_lev = p.d_batch_lev;
        la.lev_nb = p.d_lev_nb;
        la.done = p.d_done;
        la.sweep = ++p.sweeps;
        la.nlev = p.nlev;
        la.sync_mode = (int)pa_knob(c, "gs_sync", 0);
        la.trace = (int)pa_knob(c, "gs_trace", 0);
        const size_t smem = (size_t)(GSB_THREADS / 32) * GSB_WARP_DOUBLES(nj) * sizeof(double);
        PA_CUDA(cudaFuncSetAttribute(wk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        PA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wk, GSB_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        const int64_t cap = pa_knob(c, "gs_ctas", 0);
        if (cap > 0 && cap < per_sm) per_sm = (int)cap;
        // all CTAs co-resident: a waiting batch only waits for batches held by running CTAs
        const int64_t grid = std::min<int64_t>(p.npad_b / GSB_ROWS, (int64_t)nsm * per_sm);
        wk<<<(unsigned)grid, GSB_THREADS, smem, c->stream>>>(la);
        return PA_OK;
      }
      int lanes = gs_default_lanes(c, p);
      if (p.maxlen > 32) lanes = 0;
      void (*kern)(const GsArgs<PtrT>) = lanes == 4 ? k_gs_flow_pipe<PtrT, 4> : lanes == 8 ? k_gs_flow_pipe<PtrT, 8>
                                         : lanes == 16 ? k_gs_flow_pipe<PtrT, 16> : lanes == 32 ? k_gs_flow_pipe<PtrT, 32> : k_gs_flow<PtrT>;
      const int rpw = lanes == 4 ? 8 : lanes == 8 ? 4 : lanes == 16 ? 2 : 1;
      int ctas_per_sm = 0;
      PA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, GS_THREADS, 0));
      if (ctas_per_sm < 1) ctas_per_sm = 1;
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

// mul_no_lat!(Axf, A, x) followed by restrict! reads Axf only at the injection points (one fine row in eight): this
// kernel computes exactly those rows — r_c[i] = b_f[v] - (A x)[v], v = f2c[i] — with the arithmetic of spmv_csr!
// (acc += a*x[col] in CSR order, separate multiply and add), so r_c has the bits of the two-step reference sequence while
// 7/8 of the residual SpMV's matrix traffic is never read.  One warp per coarse row: lane k loads entry k (one coalesced
// read of the row), multiplies, and the products are added in column order by passing them to lane 0.
template <typename PtrT>
__global__ void __launch_bounds__(256, 4) k_residual_restrict(double *rc, const double *bf, const double *x, const PtrT *rowptr, const int32_t *colval,
                                                           const double *nzval, int64_t nc, int64_t cx, int64_t cy, int64_t fx, int64_t fy) {
  constexpr int R = 4;  // coarse rows in flight per warp: the three dependent reads (row pointer, entries, x) of R rows overlap
  __shared__ double rr_prod[8 * R * 33];
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
  const uint32_t ucx = (uint32_t)cx, ucy = (uint32_t)cy;
  for (int64_t i0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R; i0 < nc; i0 += nw * R) {
    int64_t f[R], p0[R];
    int len[R];
    bool shortrows = true;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const uint32_t i = (uint32_t)min(i0 + r, nc - 1);  // (rows < 2^31: pa_gs_commit)
      const uint32_t ix = i % ucx, t = i / ucx, iy = t % ucy, iz = t / ucy;
      f[r] = 2 * (int64_t)ix + fx * (2 * (int64_t)iy + fy * 2 * (int64_t)iz);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      p0[r] = (int64_t)rowptr[f[r]];
      len[r] = (int)((int64_t)rowptr[f[r] + 1] - p0[r]);
      shortrows &= len[r] <= 32;
    }
    if (shortrows) {
      double v[R], xv[R];
      int32_t c[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const bool in = lane < len[r];
        v[r] = in ? nzval[p0[r] + lane] : 0.0;
        c[r] = in ? colval[p0[r] + lane] : 0;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) xv[r] = x[c[r]];
      // products parked in shared memory, then lane r adds row r's products in column order (R chains side by side)
      double *sp = rr_prod + (size_t)(threadIdx.x >> 5) * (R * 33);
#pragma unroll
      for (int r = 0; r < R; ++r) sp[r * 33 + lane] = __dmul_rn(v[r], xv[r]);
      __syncwarp();
      if (lane < R) {
        int mylen = len[0];
        int64_t myf = f[0];
#pragma unroll
        for (int r = 1; r < R; ++r) {
          mylen = lane == r ? len[r] : mylen;
          myf = lane == r ? f[r] : myf;
        }
        const double bv = bf[myf];
        double acc = 0.0;
        for (int k = 0; k < mylen; ++k) acc = __dadd_rn(acc, sp[lane * 33 + k]);
        if (i0 + lane < nc) rc[i0 + lane] = __dsub_rn(bv, acc);
      }
      __syncwarp();
    } else {
      for (int r = 0; r < R && i0 + r < nc; ++r) {  // rows of any length: 32 entries per trip, the chain continues across trips
        double acc = 0.0;
        for (int64_t q = p0[r]; q < p0[r] + len[r]; q += 32) {
          const int n = (int)min((int64_t)32, p0[r] + len[r] - q);
          double prod = 0.0;
          if (lane < n) prod = __dmul_rn(nzval[q + lane], x[colval[q + lane]]);
          for (int k = 0; k < n; ++k) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, prod, k));
        }
        if (lane == 0) rc[i0 + r] = __dsub_rn(bf[f[r]], acc);
      }
    }
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
  // mg_fused_restrict 1 (default): residual at the injection points only (k_residual_restrict); 0: the reference's two steps
  const bool fused = pa_knob(c, "mg_fused_restrict", 1) != 0 && !M->A[l]->subassembled;
  if (fused) PA_TRY(pa_vec_consistent(x));  // mul_no_lat! = consistent!(x), then the local product
  else PA_TRY(pa_spmv(M->A[l], x, M->Axf[l], 1.0, 0.0, PA_SPMV_DEFAULT));  // mul_no_lat!(Axf, A, x)
  PA_TRY(pa_before_write(c));
  for (int k = 0; k < c->nlocal; ++k) {
    const auto &dc = M->dims[l - 1][k], &df = M->dims[l][k];
    const int64_t nc = dc[0] * dc[1] * dc[2];
    if (!nc) continue;
    if (fused) {
      const MatPart &m = M->A[l]->parts[k];
      const int64_t grid = std::min<int64_t>((nc + 31) / 32, 148 * 8);
      if (m.ptr64)
        k_residual_restrict<int64_t><<<(unsigned)grid, 256, 0, c->stream>>>(M->r[l - 1]->d[k], b->d[k], x->d[k], (const int64_t *)m.d_rowptr, m.d_colval, m.d_nzval,
                                                                          nc, dc[0], dc[1], df[0], df[1]);
      else
        k_residual_restrict<int32_t><<<(unsigned)grid, 256, 0, c->stream>>>(M->r[l - 1]->d[k], b->d[k], x->d[k], (const int32_t *)m.d_rowptr, m.d_colval, m.d_nzval,
                                                                          nc, dc[0], dc[1], df[0], df[1]);
    } else {
      k_restrict<<<small_grid(nc), 256, 0, c->stream>>>(M->r[l - 1]->d[k], b->d[k], M->Axf[l]->d[k], nc, dc[0], dc[1], df[0], df[1]);
    }
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
