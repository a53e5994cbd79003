// Backend runtime: context (array of parts), symmetric peer-mapped arena, cross-GPU signalling,
// scalar all-reduce, and the exchange plan (PRange + VectorAssemblyCache ingest).
// Reference behaviour mirrored: src/debug_array.jl, src/mpi_array.jl (backends),
// src/p_range.jl:417-531 + src/p_vector.jl:418-468 (plan), src/primitives.jl:992-1042 (exchange).
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "pa_internal.h"

// ------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";

void pa_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int pa_cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  pa_set_error("CUDA error %s (%s) at %s:%d in %s", cudaGetErrorName(e), cudaGetErrorString(e), file, line, what);
  return PA_ECUDA;
}
extern "C" const char *pa_last_error(void) { return g_err; }
extern "C" int pa_abi_version(void) { return PA_ABI_VERSION; }

int64_t pa_knob(pa_ctx *ctx, const char *key, int64_t dflt) {
  auto it = ctx->knobs.find(key);
  if (it != ctx->knobs.end()) return it->second;
  std::string env = std::string("PA_") + key;
  for (auto &c : env) c = (char)toupper(c);
  const char *v = getenv(env.c_str());
  if (!v) return dflt;  // defaults may depend on the caller's matrix: never cached
  int64_t r = atoll(v);
  ctx->knobs[key] = r;
  return r;
}
extern "C" int pa_ctx_set_knob(pa_ctx *ctx, const char *key, int64_t value) {
  PA_CHECK(ctx && key, PA_EINVAL, "pa_ctx_set_knob: null argument");
  ctx->knobs[key] = value;
  return PA_OK;
}

// ------------------------------------------------------------------ signalling kernels
// One u64 "arrive" and one u64 "done" flag per (part, neighbour part) live in the header of the
// owner's arena.  A neighbour writes them with system-scope stores over NVLink; the owner polls its
// own HBM.  Epoch counters are device resident so that a captured CUDA graph can be replayed.
#define PA_SPIN_LIMIT (200000000000LL)  // ~100 s of SM cycles: a protocol bug must not hang the box

__global__ void k_signal(unsigned long long *epoch, int bump, FlagPtrs dst, int n) {
  // all earlier kernels of this stream are complete (stream order) -> make their writes visible
  // system wide, then publish the new epoch to every neighbour.
  unsigned long long e = *epoch + (unsigned long long)bump;
  __syncthreads();
  if (threadIdx.x == 0 && bump) *epoch = e;
  __threadfence_system();
  if ((int)threadIdx.x < n) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst.p[threadIdx.x]), "l"(e) : "memory");
  }
}

__global__ void k_wait(const unsigned long long *epoch, FlagPtrs src, int n, int *err) {
  if ((int)threadIdx.x < n) {
    const unsigned long long want = *epoch;
    const unsigned long long *f = src.p[threadIdx.x];
    unsigned long long got;
    long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(f) : "memory");
      if (got >= want) break;
      if (clock64() - t0 > PA_SPIN_LIMIT) {
        *err = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __threadfence_system();
}

static uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------ context
extern "C" int pa_ctx_create(int32_t nparts_global, int32_t nlocal, const int32_t *part_ids, int32_t device,
                             uint64_t arena_bytes, void *stream, pa_ctx **out) {
  PA_CHECK(out && part_ids && nparts_global > 0 && nlocal > 0 && nlocal <= nparts_global, PA_EINVAL,
           "pa_ctx_create: bad arguments (nparts=%d nlocal=%d)", nparts_global, nlocal);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    pa_set_error("pa_ctx_create: no CUDA device available (%s) — this backend has no CPU fallback",
                 cudaGetErrorString(e));
    return PA_ECUDA;
  }
  PA_CHECK(device >= 0 && device < ndev, PA_EINVAL, "pa_ctx_create: device %d out of range (%d devices)", device, ndev);
  PA_CUDA(cudaSetDevice(device));
  pa_ctx *c = new pa_ctx();
  c->nparts = nparts_global;
  c->nlocal = nlocal;
  c->device = device;
  c->local_of_part.assign(nparts_global, -1);
  for (int k = 0; k < nlocal; ++k) {
    int p = part_ids[k] - 1;
    if (p < 0 || p >= nparts_global || c->local_of_part[p] != -1) {
      delete c;
      pa_set_error("pa_ctx_create: invalid or repeated part id %d", part_ids[k]);
      return PA_EINVAL;
    }
    c->part_ids.push_back(p);
    c->local_of_part[p] = k;
  }
  if (stream) {
    c->stream = (cudaStream_t)stream;
  } else {
    PA_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  PA_CUDA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
  PA_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  PA_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  c->arena_bytes = align_up(arena_bytes ? arena_bytes : (1ull << 30), 1 << 20);
  // header: arrive[nparts] | done[nparts] | red_flag[2][nparts] | red_val[2][nparts]   (8 bytes each)
  c->hdr_bytes = align_up(6ull * nparts_global * sizeof(unsigned long long), 4096);
  PA_CHECK(c->arena_bytes > c->hdr_bytes, PA_EINVAL, "pa_ctx_create: arena too small");
  c->bump = c->hdr_bytes;
  c->peer_base.assign(nparts_global, nullptr);
  c->peer_ipc.assign(nparts_global, false);
  for (int k = 0; k < nlocal; ++k) {
    char *p = nullptr;
    cudaError_t me = cudaMalloc((void **)&p, c->arena_bytes);
    if (me != cudaSuccess) {
      pa_set_error("pa_ctx_create: cannot allocate a %.2f GiB arena for part %d (%s)", c->arena_bytes / 1073741824.0,
                   c->part_ids[k] + 1, cudaGetErrorString(me));
      cudaGetLastError();
      return PA_ENOMEM;
    }
    PA_CUDA(cudaMemsetAsync(p, 0, c->hdr_bytes, c->stream));
    c->arena.push_back(p);
    c->peer_base[c->part_ids[k]] = p;
  }
  PA_CUDA(cudaMalloc((void **)&c->d_scal, PA_NSCAL * sizeof(double)));
  PA_CUDA(cudaMemsetAsync(c->d_scal, 0, PA_NSCAL * sizeof(double), c->stream));
  PA_CUDA(cudaMalloc((void **)&c->d_partial, nlocal * sizeof(double)));
  PA_CUDA(cudaMalloc((void **)&c->d_blockpart, (size_t)nlocal * PA_RED_BLOCKS * sizeof(double)));
  PA_CUDA(cudaMalloc((void **)&c->d_ticket, nlocal * sizeof(unsigned)));
  PA_CUDA(cudaMemsetAsync(c->d_ticket, 0, nlocal * sizeof(unsigned), c->stream));
  PA_CUDA(cudaMalloc((void **)&c->d_epoch, nlocal * sizeof(unsigned long long)));
  PA_CUDA(cudaMemsetAsync(c->d_epoch, 0, nlocal * sizeof(unsigned long long), c->stream));
  PA_CUDA(cudaMalloc((void **)&c->d_red_epoch, sizeof(unsigned long long)));
  PA_CUDA(cudaMemsetAsync(c->d_red_epoch, 0, sizeof(unsigned long long), c->stream));
  PA_CUDA(cudaMalloc((void **)&c->d_err, sizeof(int)));
  PA_CUDA(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
  PA_CUDA(cudaMalloc((void **)&c->d_cons_ticket, 2 * sizeof(unsigned)));
  PA_CUDA(cudaMemsetAsync(c->d_cons_ticket, 0, 2 * sizeof(unsigned), c->stream));
  PA_CUDA(cudaHostAlloc((void **)&c->h_scal, PA_NSCAL * sizeof(double), cudaHostAllocDefault));
  PA_CUDA(cudaHostAlloc((void **)&c->h_err, sizeof(int), cudaHostAllocDefault));
  *c->h_err = 0;
  PA_CUDA(cudaStreamSynchronize(c->stream));
  *out = c;
  return PA_OK;
}

typedef int (*nccl_destroy_t)(void *);
static void *g_nccl = nullptr;

extern "C" int pa_ctx_destroy(pa_ctx *c) {
  if (!c) return PA_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  pa_cg_drop_work(c, nullptr, nullptr, true);
  cudaFree(c->d_cons_ticket);
  if (c->nccl_comm && g_nccl) {
    nccl_destroy_t f = (nccl_destroy_t)dlsym(g_nccl, "ncclCommDestroy");
    if (f) f(c->nccl_comm);
  }
  for (int p = 0; p < c->nparts; ++p)
    if (c->peer_ipc[p] && c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
  for (char *a : c->arena) cudaFree(a);
  cudaFree(c->d_scal);
  cudaFree(c->d_partial);
  cudaFree(c->d_blockpart);
  cudaFree(c->d_ticket);
  cudaFree(c->d_epoch);
  cudaFree(c->d_err);
  cudaFree(c->d_red_epoch);
  cudaFreeHost(c->h_scal);
  cudaFreeHost(c->h_err);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return PA_OK;
}

int pa_check_device_error(pa_ctx *c) {
  PA_CUDA(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  PA_CHECK(*c->h_err == 0, PA_ESTATE,
           "a part waited >100 s for a neighbour's signal: collective calls are out of order across processes");
  return PA_OK;
}

extern "C" int pa_ctx_sync(pa_ctx *c) {
  PA_CHECK(c, PA_EINVAL, "pa_ctx_sync: null context");
  PA_CUDA(cudaSetDevice(c->device));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  return pa_check_device_error(c);
}

extern "C" int pa_ctx_stream(pa_ctx *c, void **out) {
  PA_CHECK(c && out, PA_EINVAL, "pa_ctx_stream: null argument");
  *out = (void *)c->stream;
  return PA_OK;
}

extern "C" int pa_ctx_launch_count(pa_ctx *c, int64_t *out) {
  PA_CHECK(c && out, PA_EINVAL, "pa_ctx_launch_count: null argument");
  *out = c->launches;
  return PA_OK;
}

extern "C" int pa_ctx_arena_export(pa_ctx *c, int32_t k, void *handle64) {
  PA_CHECK(c && handle64 && k >= 0 && k < c->nlocal, PA_EINVAL, "pa_ctx_arena_export: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  PA_CUDA(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  PA_CUDA(cudaIpcGetMemHandle(&h, c->arena[k]));
  memcpy(handle64, &h, 64);
  return PA_OK;
}

extern "C" int pa_ctx_arena_import(pa_ctx *c, int32_t part_id, const void *handle64) {
  PA_CHECK(c && handle64 && part_id >= 1 && part_id <= c->nparts, PA_EINVAL, "pa_ctx_arena_import: bad arguments");
  int p = part_id - 1;
  if (c->local_of_part[p] >= 0) return PA_OK;  // local parts are linked already
  PA_CHECK(c->peer_base[p] == nullptr, PA_ESTATE, "pa_ctx_arena_import: part %d imported twice", part_id);
  PA_CUDA(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void *ptr = nullptr;
  PA_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
  c->peer_base[p] = (char *)ptr;
  c->peer_ipc[p] = true;
  return PA_OK;
}

/* One PROCESS driving several GPUs (SURVEY 8b "Threading": the DebugArray execution model spanning the box): n contexts, one
 * part and one device each, created and peer-linked here — the arenas are mapped with cudaDeviceEnablePeerAccess instead of
 * CUDA IPC, the scalar all-reduce uses the peer-memory kernel (no NCCL).  The caller then drives context k from its own
 * host thread (Julia: Threads.@spawn per part; Python: threading, ctypes releases the GIL): every operation stays the
 * collective it is in the one-process-per-GPU model, so the kernels, the signalling and the results are identical. */
extern "C" int pa_ctx_create_multi(int32_t ndev, const int32_t *devices, uint64_t arena_bytes, pa_ctx **out) {
  PA_CHECK(ndev >= 1 && ndev <= PA_MAX_NBR && devices && out, PA_EINVAL, "pa_ctx_create_multi: bad arguments");
  for (int i = 0; i < ndev; ++i)
    for (int j = 0; j < i; ++j) PA_CHECK(devices[i] != devices[j], PA_EINVAL, "pa_ctx_create_multi: device %d listed twice", devices[i]);
  for (int k = 0; k < ndev; ++k) out[k] = nullptr;
  for (int k = 0; k < ndev; ++k) {
    const int32_t id = k + 1;
    int rc = pa_ctx_create(ndev, 1, &id, devices[k], arena_bytes, nullptr, &out[k]);
    if (rc != PA_OK) {
      for (int q = 0; q < k; ++q) pa_ctx_destroy(out[q]);
      return rc;
    }
    out[k]->rank = k;
    out[k]->world = ndev;
  }
  for (int k = 0; k < ndev; ++k) {
    PA_CUDA(cudaSetDevice(devices[k]));
    for (int q = 0; q < ndev; ++q) {
      if (q == k) continue;
      int can = 0;
      PA_CUDA(cudaDeviceCanAccessPeer(&can, devices[k], devices[q]));
      PA_CHECK(can, PA_ECUDA, "pa_ctx_create_multi: device %d cannot access device %d (no NVLink/PCIe peer path)", devices[k], devices[q]);
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[q], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return pa_cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
      cudaGetLastError();
      out[k]->peer_base[q] = out[q]->arena[0];  // same process: the peer's arena pointer is valid here (UVA)
    }
  }
  return PA_OK;
}

// ------------------------------------------------------------------ NCCL (dlopen, scalars only)
typedef struct { char internal[128]; } nccl_uid;
typedef int (*nccl_getuid_t)(nccl_uid *);
typedef int (*nccl_initrank_t)(void **, int, nccl_uid, int);
typedef int (*nccl_allreduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*nccl_errstr_t)(int);
static nccl_allreduce_t g_allreduce = nullptr;
static nccl_errstr_t g_errstr = nullptr;

static int nccl_load() {
  if (g_nccl) return PA_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; names[i] && !g_nccl; ++i) g_nccl = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  PA_CHECK(g_nccl, PA_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
  g_allreduce = (nccl_allreduce_t)dlsym(g_nccl, "ncclAllReduce");
  g_errstr = (nccl_errstr_t)dlsym(g_nccl, "ncclGetErrorString");
  PA_CHECK(g_allreduce && g_errstr, PA_ENCCL, "libnccl lacks ncclAllReduce");
  return PA_OK;
}

extern "C" int pa_nccl_unique_id(void *id128) {
  PA_CHECK(id128, PA_EINVAL, "pa_nccl_unique_id: null");
  PA_TRY(nccl_load());
  nccl_getuid_t f = (nccl_getuid_t)dlsym(g_nccl, "ncclGetUniqueId");
  PA_CHECK(f, PA_ENCCL, "libnccl lacks ncclGetUniqueId");
  nccl_uid u;
  int r = f(&u);
  PA_CHECK(r == 0, PA_ENCCL, "ncclGetUniqueId: %s", g_errstr(r));
  memcpy(id128, &u, 128);
  return PA_OK;
}

extern "C" int pa_ctx_nccl_init(pa_ctx *c, const void *id128, int32_t rank, int32_t world) {
  PA_CHECK(c && id128 && world >= 1 && rank >= 0 && rank < world, PA_EINVAL, "pa_ctx_nccl_init: bad arguments");
  PA_TRY(nccl_load());
  PA_CUDA(cudaSetDevice(c->device));
  nccl_initrank_t f = (nccl_initrank_t)dlsym(g_nccl, "ncclCommInitRank");
  PA_CHECK(f, PA_ENCCL, "libnccl lacks ncclCommInitRank");
  nccl_uid u;
  memcpy(&u, id128, 128);
  int r = f(&c->nccl_comm, world, u, rank);
  PA_CHECK(r == 0, PA_ENCCL, "ncclCommInitRank: %s", g_errstr(r));
  c->rank = rank;
  c->world = world;
  return PA_OK;
}

// sum of the local parts' partials in part order, then across processes
__global__ void k_sum_partials(const double *partial, int n, double *out) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += partial[i];
  *out = s;
}

// Scalar all-reduce over NVLink peer memory (replaces ncclAllReduce for the CG scalars when every process holds one
// part): each part pushes its value and a sequence flag into a double-buffered slot of every part's arena header with
// system-scope stores, waits for all flags, and adds the P slots in PART ORDER — the order of the reference's sequential
// `sum` over parts (src/primitives.jl:693-698), bitwise identical on all parts, one kernel, ~NVLink store latency.
// Buffer parity e&1 is safe: a part can only reach reduction e+2 after every peer has posted e+1, i.e. finished reading e.
struct RedPeers {
  unsigned long long *flag[PA_MAX_NBR];
  double *val[PA_MAX_NBR];
};

__global__ void k_allreduce_peer(double *inout, unsigned long long *epoch, RedPeers peers, const unsigned long long *my_flag,
                                 const double *my_val, int me, int nparts, int *err) {
  __shared__ double got[PA_MAX_NBR];
  const unsigned long long e = *epoch + 1;
  const int par = (int)(e & 1ull);
  const int q = threadIdx.x;
  __syncthreads();
  if (q < nparts) {
    const double v = *inout;
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(peers.val[q] + par * nparts + me), "d"(v) : "memory");
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.flag[q] + par * nparts + me), "l"(e) : "memory");
    // wait for part q's contribution to arrive in MY header
    const unsigned long long *f = my_flag + par * nparts + q;
    unsigned long long seen;
    long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(f) : "memory");
      if (seen >= e) break;
      if (clock64() - t0 > PA_SPIN_LIMIT) {
        *err = 1;
        break;
      }
    }
    double x;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(x) : "l"(my_val + par * nparts + q) : "memory");
    got[q] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += got[i];  // part order
    *inout = s;
    *epoch = e;
  }
}

static bool peer_allreduce_ok(pa_ctx *c) {
  if (c->world <= 1 || c->nlocal != 1 || c->nparts != c->world || c->nparts > PA_MAX_NBR) return false;
  if (pa_knob(c, "nccl_allreduce", 0)) return false;
  for (int p = 0; p < c->nparts; ++p)
    if (!c->peer_base[p]) return false;
  return true;
}

// Folded reductions / signalling: one local part whose peers are all mapped (one part per process), or a one-part job.
bool pa_fold_ok(pa_ctx *c) {
  if (pa_knob(c, "no_fold", 0)) return false;
  if (c->nlocal != 1 || c->nparts > PA_MAX_NBR) return false;
  if (c->nparts == 1) return true;
  return peer_allreduce_ok(c);
}
RedPush pa_red_push(pa_ctx *c) {
  RedPush rp;
  for (int p = 0; p < PA_MAX_NBR; ++p) {
    unsigned long long *hdr = p < c->nparts ? (unsigned long long *)c->peer_base[p] : nullptr;
    rp.flag[p] = hdr ? hdr + 2 * c->nparts : nullptr;
    rp.val[p] = hdr ? (double *)(hdr + 4 * c->nparts) : nullptr;
  }
  rp.epoch = c->d_red_epoch;
  rp.me = c->part_ids[0];
  rp.nparts = c->nparts;
  return rp;
}
RedWait pa_red_wait(pa_ctx *c) {
  unsigned long long *mine = (unsigned long long *)c->arena[0];
  RedWait w;
  w.flag = mine + 2 * c->nparts;
  w.val = (const double *)(mine + 4 * c->nparts);
  w.epoch = c->d_red_epoch;
  w.nparts = c->nparts;
  w.err = c->d_err;
  return w;
}

int pa_reduce_finish(pa_ctx *c, double *d_out) {
  if (peer_allreduce_ok(c)) {
    RedPeers rp;
    for (int p = 0; p < c->nparts; ++p) {
      unsigned long long *hdr = (unsigned long long *)c->peer_base[p];
      rp.flag[p] = hdr + 2 * c->nparts;
      rp.val[p] = (double *)(hdr + 4 * c->nparts);
    }
    unsigned long long *mine = (unsigned long long *)c->arena[0];
    k_allreduce_peer<<<1, 32, 0, c->stream>>>(d_out, c->d_red_epoch, rp, mine + 2 * c->nparts, (const double *)(mine + 4 * c->nparts),
                                              c->part_ids[0], c->nparts, c->d_err);
    c->launches++;
    PA_CUDA(cudaGetLastError());
    return PA_OK;
  }
  if (c->nlocal > 1) {
    k_sum_partials<<<1, 1, 0, c->stream>>>(c->d_partial, c->nlocal, d_out);
    c->launches++;
    PA_CUDA(cudaGetLastError());
  }  // nlocal == 1: the reduction kernel wrote *d_out directly
  if (c->world > 1) {
    PA_CHECK(c->nccl_comm, PA_ESTATE, "distributed reduction without pa_ctx_nccl_init");
    int r = g_allreduce(d_out, d_out, 1, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->nccl_comm, c->stream);
    PA_CHECK(r == 0, PA_ENCCL, "ncclAllReduce: %s", g_errstr(r));
    c->launches++;
  }
  return PA_OK;
}

int pa_read_scalars(pa_ctx *c, int first, int count, double *out) {
  PA_CUDA(cudaMemcpyAsync(c->h_scal + first, c->d_scal + first, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < count; ++i) out[i] = c->h_scal[first + i];
  return PA_OK;
}

// ------------------------------------------------------------------ collective brackets
static unsigned long long *flag_addr(pa_ctx *c, int owner_part, int from_part, bool done) {
  return (unsigned long long *)c->peer_base[owner_part] + (done ? c->nparts : 0) + from_part;
}

static int signal_all(pa_plan *plan, bool done) {
  pa_ctx *c = plan->ctx;
  for (int k = 0; k < c->nlocal; ++k) {
    const PlanPart &pp = plan->parts[k];
    int n = (int)pp.nbrs.size();
    // a part without neighbours in THIS collective still counts it: the epochs of all parts advance in lockstep
    // (a job of one part has nobody to stay in step with: no launch at all)
    if (!n && (done || c->nparts == 1)) continue;
    FlagPtrs dst;
    for (int i = 0; i < n; ++i) {
      PA_CHECK(c->peer_base[pp.nbrs[i]], PA_ESTATE, "part %d's arena was never imported (pa_ctx_arena_import)", pp.nbrs[i] + 1);
      dst.p[i] = flag_addr(c, pp.nbrs[i], c->part_ids[k], done);
    }
    k_signal<<<1, 32, 0, c->stream>>>(c->d_epoch + k, done ? 0 : 1, dst, n);
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}

static int wait_all(pa_plan *plan, bool done) {
  pa_ctx *c = plan->ctx;
  for (int k = 0; k < c->nlocal; ++k) {
    const PlanPart &pp = plan->parts[k];
    int n = (int)pp.nbrs.size();
    if (!n) continue;
    FlagPtrs src;
    for (int i = 0; i < n; ++i) src.p[i] = flag_addr(c, c->part_ids[k], pp.nbrs[i], done);
    k_wait<<<1, 32, 0, c->stream>>>(c->d_epoch + k, src, n, c->d_err);
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  return PA_OK;
}

// publish + wait in ONE launch (one part per process: nobody else's signal has to be enqueued in between)
__global__ void k_signal_wait(unsigned long long *epoch, FlagPtrs dst, FlagPtrs src, int n, int *err) {
  unsigned long long e = *epoch + 1ull;
  __syncthreads();
  if (threadIdx.x == 0) *epoch = e;
  __threadfence_system();
  if ((int)threadIdx.x < n) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst.p[threadIdx.x]), "l"(e) : "memory");
    const unsigned long long *f = src.p[threadIdx.x];
    unsigned long long got;
    long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(f) : "memory");
      if (got >= e) break;
      if (clock64() - t0 > PA_SPIN_LIMIT) {
        *err = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __threadfence_system();
}

int pa_collective_begin(pa_plan *plan) {
  pa_ctx *c = plan->ctx;
  if (c->nlocal == 1) {
    const PlanPart &pp = plan->parts[0];
    const int n = (int)pp.nbrs.size();
    if (!n && c->nparts == 1) return PA_OK;  // one part in the whole job: nothing to order
    FlagPtrs dst, src;
    for (int i = 0; i < n; ++i) {
      PA_CHECK(c->peer_base[pp.nbrs[i]], PA_ESTATE, "part %d's arena was never imported (pa_ctx_arena_import)", pp.nbrs[i] + 1);
      dst.p[i] = flag_addr(c, pp.nbrs[i], c->part_ids[0], false);
      src.p[i] = flag_addr(c, c->part_ids[0], pp.nbrs[i], false);
    }
    k_signal_wait<<<1, 32, 0, c->stream>>>(c->d_epoch, dst, src, n, c->d_err);
    c->launches++;
    PA_CUDA(cudaGetLastError());
    return PA_OK;
  }
  PA_TRY(signal_all(plan, false));
  return wait_all(plan, false);
}

DoneWait pa_done_wait(pa_plan *plan) {
  pa_ctx *c = plan->ctx;
  const PlanPart &pp = plan->parts[0];
  DoneWait d;
  d.n = (int)pp.nbrs.size();
  for (int i = 0; i < PA_MAX_NBR; ++i) d.flag[i] = i < d.n ? flag_addr(c, c->part_ids[0], pp.nbrs[i], true) : nullptr;
  d.epoch = c->d_epoch;
  d.err = c->d_err;
  return d;
}

// flag addresses of the fused signal + gather + done kernel (pa_vector.cu): arrive/done at the neighbours, arrive here
int pa_sync_flags(pa_plan *plan, FlagPtrs *arrive_dst, FlagPtrs *arrive_src, FlagPtrs *done_dst, int *n) {
  pa_ctx *c = plan->ctx;
  const PlanPart &pp = plan->parts[0];
  *n = (int)pp.nbrs.size();
  for (int i = 0; i < *n; ++i) {
    PA_CHECK(c->peer_base[pp.nbrs[i]], PA_ESTATE, "part %d's arena was never imported (pa_ctx_arena_import)", pp.nbrs[i] + 1);
    arrive_dst->p[i] = flag_addr(c, pp.nbrs[i], c->part_ids[0], false);
    arrive_src->p[i] = flag_addr(c, c->part_ids[0], pp.nbrs[i], false);
    done_dst->p[i] = flag_addr(c, pp.nbrs[i], c->part_ids[0], true);
  }
  return PA_OK;
}

int pa_collective_end(pa_plan *plan) {
  PA_TRY(signal_all(plan, true));
  pa_ctx *c = plan->ctx;
  bool any = false;
  for (auto &pp : plan->parts) any |= !pp.nbrs.empty();
  if (any && std::find(c->pending_done.begin(), c->pending_done.end(), plan) == c->pending_done.end())
    c->pending_done.push_back(plan);
  return PA_OK;
}

void pa_mark_pending_done(pa_plan *plan) {
  pa_ctx *c = plan->ctx;
  bool any = false;
  for (auto &pp : plan->parts) any |= !pp.nbrs.empty();
  if (any && std::find(c->pending_done.begin(), c->pending_done.end(), plan) == c->pending_done.end()) c->pending_done.push_back(plan);
}

int pa_before_write(pa_ctx *c) {
  for (pa_plan *p : c->pending_done) PA_TRY(wait_all(p, true));
  c->pending_done.clear();
  return PA_OK;
}

// ------------------------------------------------------------------ plan
extern "C" int pa_plan_create(pa_ctx *ctx, pa_plan **out) {
  PA_CHECK(ctx && out, PA_EINVAL, "pa_plan_create: null argument");
  pa_plan *p = new pa_plan();
  p->ctx = ctx;
  p->parts.resize(ctx->nlocal);
  *out = p;
  return PA_OK;
}

static void to0(std::vector<int32_t> &dst, const int32_t *src, int64_t n) {
  dst.resize(n);
  for (int64_t i = 0; i < n; ++i) dst[i] = src[i] - 1;
}

extern "C" int pa_plan_set_part(pa_plan *plan, int32_t k, int64_t n_local, int64_t n_own, const int32_t *own_to_local,
                                const int32_t *ghost_to_local, int32_t n_nbr_snd, const int32_t *nbr_snd,
                                const int32_t *snd_ptrs, const int32_t *snd_lids, const int32_t *snd_remote_lids,
                                int32_t n_nbr_rcv, const int32_t *nbr_rcv, const int32_t *rcv_ptrs,
                                const int32_t *rcv_lids, const int32_t *rcv_remote_lids) {
  PA_CHECK(plan && !plan->committed, PA_ESTATE, "pa_plan_set_part: plan missing or already committed");
  pa_ctx *c = plan->ctx;
  PA_CHECK(k >= 0 && k < c->nlocal, PA_EINVAL, "pa_plan_set_part: local part %d out of range", k);
  PA_CHECK(n_local >= 0 && n_own >= 0 && n_own <= n_local && n_local < (1ll << 31), PA_EINVAL,
           "pa_plan_set_part: bad sizes n_local=%lld n_own=%lld", (long long)n_local, (long long)n_own);
  PA_CHECK((own_to_local == nullptr) == (ghost_to_local == nullptr) || n_local == n_own || n_own == 0, PA_EINVAL,
           "pa_plan_set_part: give both own_to_local and ghost_to_local or neither");
  PA_CHECK(n_nbr_snd >= 0 && n_nbr_rcv >= 0, PA_EINVAL, "pa_plan_set_part: negative neighbour count");
  PA_CHECK((n_nbr_snd == 0 || (nbr_snd && snd_ptrs)) && (n_nbr_rcv == 0 || (nbr_rcv && rcv_ptrs)), PA_EINVAL,
           "pa_plan_set_part: null neighbour arrays");
  PlanPart &pp = plan->parts[k];
  pp = PlanPart();
  pp.n_local = n_local;
  pp.n_own = n_own;
  pp.n_ghost = n_local - n_own;
  pp.prefix = (own_to_local == nullptr);
  if (own_to_local) {
    to0(pp.own_to_local, own_to_local, n_own);
    if (ghost_to_local) to0(pp.ghost_to_local, ghost_to_local, pp.n_ghost);
    bool pre = true;
    for (int64_t i = 0; i < n_own && pre; ++i) pre = pp.own_to_local[i] == i;
    for (int64_t i = 0; i < pp.n_ghost && pre && ghost_to_local; ++i) pre = pp.ghost_to_local[i] == n_own + i;
    for (auto v : pp.own_to_local) PA_CHECK(v >= 0 && v < n_local, PA_EINVAL, "own_to_local out of range");
    for (auto v : pp.ghost_to_local) PA_CHECK(v >= 0 && v < n_local, PA_EINVAL, "ghost_to_local out of range");
    if (pre) {
      pp.prefix = true;
      pp.own_to_local.clear();
      pp.ghost_to_local.clear();
    }
  }
  auto ingest = [&](int32_t nn, const int32_t *nbr, const int32_t *ptrs, const int32_t *lids, const int32_t *rl,
                    std::vector<int32_t> &onbr, std::vector<int32_t> &optrs, std::vector<int32_t> &olids,
                    std::vector<int32_t> &orl, bool &has_rl) -> int {
    to0(onbr, nbr, nn);
    optrs.assign(nn + 1, 0);
    for (int i = 0; i <= nn && nn > 0; ++i) optrs[i] = ptrs[i] - 1;
    int64_t tot = nn ? optrs[nn] : 0;
    PA_CHECK(tot >= 0 && (tot == 0 || lids), PA_EINVAL, "pa_plan_set_part: bad ptrs / null lids");
    to0(olids, lids, tot);
    has_rl = rl != nullptr;
    if (rl) to0(orl, rl, tot);
    for (int i = 0; i < nn; ++i) {
      PA_CHECK(onbr[i] >= 0 && onbr[i] < c->nparts && onbr[i] != c->part_ids[k], PA_EINVAL,
               "pa_plan_set_part: neighbour id %d invalid", onbr[i] + 1);
      PA_CHECK(optrs[i + 1] >= optrs[i], PA_EINVAL, "pa_plan_set_part: ptrs not monotone");
    }
    for (auto v : olids) PA_CHECK(v >= 0 && v < n_local, PA_EINVAL, "pa_plan_set_part: local id out of range");
    return PA_OK;
  };
  PA_TRY(ingest(n_nbr_snd, nbr_snd, snd_ptrs, snd_lids, snd_remote_lids, pp.nbr_snd, pp.snd_ptrs, pp.snd_lids,
                pp.snd_rlids, pp.has_snd_rl));
  PA_TRY(ingest(n_nbr_rcv, nbr_rcv, rcv_ptrs, rcv_lids, rcv_remote_lids, pp.nbr_rcv, pp.rcv_ptrs, pp.rcv_lids,
                pp.rcv_rlids, pp.has_rcv_rl));
  pp.nbrs = pp.nbr_snd;
  pp.nbrs.insert(pp.nbrs.end(), pp.nbr_rcv.begin(), pp.nbr_rcv.end());
  std::sort(pp.nbrs.begin(), pp.nbrs.end());
  pp.nbrs.erase(std::unique(pp.nbrs.begin(), pp.nbrs.end()), pp.nbrs.end());
  PA_CHECK(pp.nbrs.size() <= PA_MAX_NBR, PA_EINVAL, "pa_plan_set_part: more than %d neighbours", PA_MAX_NBR);
  pp.set = true;
  return PA_OK;
}

template <typename T>
static int upload(T **dst, const std::vector<T> &src, cudaStream_t s) {
  *dst = nullptr;
  if (src.empty()) return PA_OK;
  PA_CUDA(cudaMalloc((void **)dst, src.size() * sizeof(T)));
  PA_CUDA(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return PA_OK;
}

// Fill the remote-lid lists for neighbours held by this process: part q's matching segment.
static int derive_remote(pa_plan *plan, int k, bool snd_side) {
  pa_ctx *c = plan->ctx;
  PlanPart &pp = plan->parts[k];
  auto &nbr = snd_side ? pp.nbr_snd : pp.nbr_rcv;
  auto &ptrs = snd_side ? pp.snd_ptrs : pp.rcv_ptrs;
  auto &rl = snd_side ? pp.snd_rlids : pp.rcv_rlids;
  rl.assign(ptrs.empty() ? 0 : ptrs.back(), -1);
  for (size_t i = 0; i < nbr.size(); ++i) {
    int kq = c->local_of_part[nbr[i]];
    PA_CHECK(kq >= 0, PA_EINVAL,
             "plan part %d: neighbour %d is remote, the *_remote_lids arrays are required", c->part_ids[k] + 1, nbr[i] + 1);
    PlanPart &q = plan->parts[kq];
    auto &qnbr = snd_side ? q.nbr_rcv : q.nbr_snd;
    auto &qptrs = snd_side ? q.rcv_ptrs : q.snd_ptrs;
    auto &qlids = snd_side ? q.rcv_lids : q.snd_lids;
    auto it = std::find(qnbr.begin(), qnbr.end(), c->part_ids[k]);
    PA_CHECK(it != qnbr.end(), PA_EINVAL, "exchange graph inconsistent between parts %d and %d (is_consistent, src/primitives.jl:861)",
             c->part_ids[k] + 1, nbr[i] + 1);
    size_t j = it - qnbr.begin();
    int len = ptrs[i + 1] - ptrs[i];
    PA_CHECK(qptrs[j + 1] - qptrs[j] == len, PA_EINVAL, "segment length mismatch between parts %d and %d", c->part_ids[k] + 1, nbr[i] + 1);
    for (int t = 0; t < len; ++t) rl[ptrs[i] + t] = qlids[qptrs[j] + t];
  }
  return PA_OK;
}

extern "C" int pa_plan_commit(pa_plan *plan, int64_t sym_n_local) {
  PA_CHECK(plan && !plan->committed, PA_ESTATE, "pa_plan_commit: plan missing or already committed");
  pa_ctx *c = plan->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  int64_t mx = 0;
  for (int k = 0; k < c->nlocal; ++k) {
    PA_CHECK(plan->parts[k].set, PA_ESTATE, "pa_plan_commit: local part %d not set", k);
    mx = std::max(mx, plan->parts[k].n_local);
  }
  if (sym_n_local == 0) {
    PA_CHECK(c->nlocal == c->nparts, PA_EINVAL, "pa_plan_commit: sym_n_local is required when parts are remote");
    sym_n_local = mx;
  }
  PA_CHECK(sym_n_local >= mx, PA_EINVAL, "pa_plan_commit: sym_n_local smaller than a local part");
  plan->sym_n_local = sym_n_local;
  plan->vec_bytes = align_up((uint64_t)std::max<int64_t>(sym_n_local, 1) * sizeof(double), 512);
  for (int k = 0; k < c->nlocal; ++k) {
    PlanPart &pp = plan->parts[k];
    if (!pp.has_snd_rl) PA_TRY(derive_remote(plan, k, true));
    if (!pp.has_rcv_rl) PA_TRY(derive_remote(plan, k, false));
  }
  for (int k = 0; k < c->nlocal; ++k) {
    PlanPart &pp = plan->parts[k];
    auto slot_of = [&](int part) { return (int32_t)(std::lower_bound(pp.nbrs.begin(), pp.nbrs.end(), part) - pp.nbrs.begin()); };
    // consistent! = reversed cache (src/p_vector.jl:427-437,748): receive into the lids of the snd lists
    std::vector<int32_t> lid, slot, rlid;
    for (size_t i = 0; i < pp.nbr_snd.size(); ++i)
      for (int t = pp.snd_ptrs[i]; t < pp.snd_ptrs[i + 1]; ++t) {
        lid.push_back(pp.snd_lids[t]);
        slot.push_back(slot_of(pp.nbr_snd[i]));
        rlid.push_back(pp.snd_rlids[t]);
      }
    pp.n_cons = (int64_t)lid.size();
    PA_TRY(upload(&pp.d_ghost_lid, lid, c->stream));
    PA_TRY(upload(&pp.d_ghost_slot, slot, c->stream));
    PA_TRY(upload(&pp.d_ghost_rlid, rlid, c->stream));
    if (pp.prefix && pp.n_ghost > 0) {
      std::vector<int32_t> gs(pp.n_ghost, -1), gr(pp.n_ghost, -1);
      for (size_t j = 0; j < lid.size(); ++j) {
        int64_t g = lid[j] - pp.n_own;
        if (g >= 0 && g < pp.n_ghost) {
          gs[g] = slot[j];
          gr[g] = rlid[j];
        }
      }
      PA_TRY(upload(&pp.d_gslot_by_gid, gs, c->stream));
      PA_TRY(upload(&pp.d_grlid_by_gid, gr, c->stream));
    }
    // assemble! (src/p_vector.jl:605-609): values[lid] = values[lid] + buf[p] for p in neighbour order;
    // group by destination lid keeping that order so one thread reproduces the sequential sum.
    {
      std::vector<std::pair<int32_t, int64_t>> ord;  // (dst lid, entry index)
      std::vector<int32_t> eslot, erl;
      for (size_t i = 0; i < pp.nbr_rcv.size(); ++i)
        for (int t = pp.rcv_ptrs[i]; t < pp.rcv_ptrs[i + 1]; ++t) {
          ord.push_back({pp.rcv_lids[t], (int64_t)eslot.size()});
          eslot.push_back(slot_of(pp.nbr_rcv[i]));
          erl.push_back(pp.rcv_rlids[t]);
        }
      std::stable_sort(ord.begin(), ord.end(), [](auto &a, auto &b) { return a.first < b.first; });
      std::vector<int32_t> dst, ptr, s2, r2;
      for (size_t j = 0; j < ord.size(); ++j) {
        if (j == 0 || ord[j].first != ord[j - 1].first) {
          dst.push_back(ord[j].first);
          ptr.push_back((int32_t)j);
        }
        s2.push_back(eslot[ord[j].second]);
        r2.push_back(erl[ord[j].second]);
      }
      ptr.push_back((int32_t)ord.size());
      pp.n_asm_dst = (int64_t)dst.size();
      PA_TRY(upload(&pp.d_asm_dst, dst, c->stream));
      if (!dst.empty()) PA_TRY(upload(&pp.d_asm_ptr, ptr, c->stream));
      PA_TRY(upload(&pp.d_asm_slot, s2, c->stream));
      PA_TRY(upload(&pp.d_asm_rlid, r2, c->stream));
      if (pp.prefix && !dst.empty()) {
        std::vector<uint32_t> bm((size_t)(pp.n_local + 31) / 32 + 1, 0u);
        for (int32_t l : dst) bm[(size_t)l >> 5] |= 1u << (l & 31);
        PA_TRY(upload(&pp.d_bnd_bitmap, bm, c->stream));
      }
    }
    if (!pp.prefix) {
      PA_TRY(upload(&pp.d_own_to_local, pp.own_to_local, c->stream));
      PA_TRY(upload(&pp.d_ghost_to_local, pp.ghost_to_local, c->stream));
    }
  }
  PA_CUDA(cudaStreamSynchronize(c->stream));
  for (int k = 0; k < c->nlocal; ++k) {  // layout signature (FNV-1a over every index array of the part)
    PlanPart &pp = plan->parts[k];
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void *p, size_t bytes) {
      const unsigned char *b = (const unsigned char *)p;
      for (size_t i = 0; i < bytes; ++i) h = (h ^ b[i]) * 1099511628211ull;
      h = (h ^ (uint64_t)bytes) * 1099511628211ull;
    };
    auto mixv = [&](const std::vector<int32_t> &v) { mix(v.data(), v.size() * sizeof(int32_t)); };
    const int64_t hdr[4] = {pp.n_local, pp.n_own, pp.prefix ? 1 : 0, (int64_t)c->part_ids[k]};
    mix(hdr, sizeof(hdr));
    mixv(pp.own_to_local); mixv(pp.ghost_to_local);
    mixv(pp.nbr_snd); mixv(pp.snd_ptrs); mixv(pp.snd_lids); mixv(pp.snd_rlids);
    mixv(pp.nbr_rcv); mixv(pp.rcv_ptrs); mixv(pp.rcv_lids); mixv(pp.rcv_rlids);
    pp.signature = h;
  }
  plan->committed = true;
  return PA_OK;
}

extern "C" int pa_plan_destroy(pa_plan *plan) {
  if (!plan) return PA_OK;
  pa_ctx *c = plan->ctx;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  auto &pd = c->pending_done;
  pd.erase(std::remove(pd.begin(), pd.end(), plan), pd.end());
  pa_cg_drop_work(c, nullptr, plan, false);
  for (auto &pp : plan->parts) {
    cudaFree(pp.d_own_to_local);
    cudaFree(pp.d_ghost_to_local);
    cudaFree(pp.d_ghost_lid);
    cudaFree(pp.d_ghost_slot);
    cudaFree(pp.d_ghost_rlid);
    cudaFree(pp.d_gslot_by_gid);
    cudaFree(pp.d_grlid_by_gid);
    cudaFree(pp.d_asm_dst);
    cudaFree(pp.d_asm_ptr);
    cudaFree(pp.d_asm_slot);
    cudaFree(pp.d_asm_rlid);
    cudaFree(pp.d_bnd_bitmap);
  }
  delete plan;
  return PA_OK;
}

extern "C" int pa_host_alloc(void **ptr, size_t bytes) {
  PA_CHECK(ptr, PA_EINVAL, "pa_host_alloc: null");
  PA_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return PA_OK;
}
extern "C" int pa_host_free(void *ptr) {
  if (ptr) PA_CUDA(cudaFreeHost(ptr));
  return PA_OK;
}
