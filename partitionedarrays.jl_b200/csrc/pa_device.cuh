// Device-side helpers shared by the kernels that carry a folded scalar all-reduce or the epoch signalling in their
// prologue / epilogue (see RedPush / RedWait / DoneWait in pa_internal.h).
#pragma once
#include "pa_internal.h"

#define PA_DEV_SPIN_LIMIT (200000000000LL)  // ~100 s of SM cycles: a protocol bug must not hang the box

__device__ __forceinline__ unsigned long long pa_ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void pa_st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool pa_spin_until(const unsigned long long *f, unsigned long long want, int *err) {
  long long t0 = 0;
  for (;;) {
    if (pa_ld_acquire_sys(f) >= want) return true;
    if (!t0) {
      t0 = clock64();
    } else if (clock64() - t0 > PA_DEV_SPIN_LIMIT) {
      *err = 1;
      return false;
    }
    __nanosleep(32);
  }
}

// Producer: called by the first `nparts` lanes of ONE warp of the kernel's last CTA (after the grid-wide fold), value
// uniform over the lanes.  Reduction e = *epoch + 1 goes to buffer e & 1: a part can only post e + 2 after it consumed
// e + 1 from every peer, and a peer posts e + 1 from the kernel whose prologue consumed e, so nobody still reads e.
__device__ __forceinline__ void pa_red_post(const RedPush &p, double v, int lane) {
  const unsigned long long e = *p.epoch + 1ull;
  const int par = (int)(e & 1ull);
  __syncwarp();
  if (lane < p.nparts) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p.val[lane] + par * p.nparts + p.me), "d"(v) : "memory");
    pa_st_release_sys(p.flag[lane] + par * p.nparts + p.me, e);
  }
  __syncwarp();
  if (lane == 0) *p.epoch = e;
}

// Consumer: every thread of the CTA calls it (contains __syncthreads); returns the sum over parts of the most recent
// reduction posted by this process' stream (stream order: the producing kernel has completed, so *epoch is final).
__device__ __forceinline__ double pa_red_sum(const RedWait &w, double *sm /* [PA_MAX_NBR + 1] shared */) {
  const unsigned long long e = *w.epoch;
  const int par = (int)(e & 1ull);
  const int q = threadIdx.x;
  if (q < w.nparts) {
    pa_spin_until(w.flag + par * w.nparts + q, e, w.err);
    double x;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(x) : "l"(w.val + par * w.nparts + q) : "memory");
    sm[q] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < w.nparts; ++i) s += sm[i];  // part order
    sm[PA_MAX_NBR] = s;
  }
  __syncthreads();
  return sm[PA_MAX_NBR];
}

// WAR guard: wait until every neighbour has reported "done reading your vectors" for the current exchange epoch
// (every thread of the CTA calls it; contains __syncthreads)
__device__ __forceinline__ void pa_wait_done(const DoneWait &d) {
  if ((int)threadIdx.x < d.n) pa_spin_until(d.flag[threadIdx.x], *d.epoch, d.err);
  __syncthreads();
}

// ------------------------------------------------------------------ mbarrier + 1-D bulk copies (TMA engine; SASS UBLKCP / SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
