// Backend primitives next to the PVector path: the vector-payload exchange!(rcv, snd, graph) and the per-part
// reductions behind reduce(op, ::PVector) / maximum / minimum / norm(v, p).
// Reference: ExchangeGraph src/primitives.jl:728-741, exchange!/exchange_impl! :992-1042 (DebugArray
// src/debug_array.jl:250-255, MPIArray src/mpi_array.jl:525-614), allocate_exchange :921-947;
// reduce src/p_vector.jl:1178-1187, norm :1201-1206, neutral_element :1170-1175.
//
// exchange!: every part holds a jagged send buffer (one segment per destination) and a jagged receive buffer (one
// segment per source).  The send buffers live in the symmetric arena, so a receiver reads the segment addressed to
// it straight from the sender's HBM (NVLink peer load): there is no message, no pack and no staging copy.  Ordering
// is the same epoch signalling as consistent!/assemble! (pa_collective_begin/end).
#include <algorithm>

#include "pa_internal.h"

struct XchgPart {
  std::vector<int32_t> snd_ids, rcv_ids;    // 0-based part ids
  std::vector<int64_t> snd_ptrs, rcv_ptrs;  // 0-based element offsets, size n+1
  std::vector<int64_t> rcv_src_off;         // offset of the segment addressed to me inside the sender's send buffer
  bool has_src_off = false, set = false;
  int64_t *d_rcv_ptrs = nullptr, *d_src_off = nullptr;
  int32_t *d_slot = nullptr;                // neighbour slot of every source
  unsigned char *d_rcv = nullptr;           // receive buffer (elem bytes per element)
};

struct pa_xchg {
  pa_ctx *ctx = nullptr;
  pa_plan *plan = nullptr;  // neighbour sets for the epoch signalling (no index data)
  std::vector<XchgPart> parts;
  uint64_t snd_off = 0, snd_bytes = 0;
  int elem = 8;  // bytes per element (1, 2, 4 or 8)
  bool committed = false;
};

// one CTA column per source segment (blockIdx.y), grid-stride over its elements
template <typename T>
__global__ void k_exchange(T *rcv, const int64_t *__restrict__ rcv_ptrs, const int32_t *__restrict__ slot,
                           const int64_t *__restrict__ src_off, PeerPtrs peers) {
  const int i = blockIdx.y;
  const int64_t lo = rcv_ptrs[i], n = rcv_ptrs[i + 1] - lo;
  const T *src = reinterpret_cast<const T *>(peers.p[slot[i]]) + src_off[i];
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
    rcv[lo + j] = __ldcg(src + j);
}

extern "C" int pa_xchg_create(pa_ctx *ctx, pa_xchg **out) {
  PA_CHECK(ctx && out, PA_EINVAL, "pa_xchg_create: null argument");
  pa_xchg *x = new pa_xchg();
  x->ctx = ctx;
  x->parts.resize(ctx->nlocal);
  *out = x;
  return PA_OK;
}

/* Element size of the payload in bytes: 8 (default; Float64 / Int64), 4 (Int32 / Float32 — the reference's index lists are
 * JaggedArray{Int32,Int32}, src/p_range.jl:489-531), 2 or 1.  Before pa_xchg_commit; ptrs and offsets count ELEMENTS. */
extern "C" int pa_xchg_set_elem_size(pa_xchg *x, int32_t bytes) {
  PA_CHECK(x && !x->committed, PA_ESTATE, "pa_xchg_set_elem_size: exchange missing or already committed");
  PA_CHECK(bytes == 1 || bytes == 2 || bytes == 4 || bytes == 8, PA_EINVAL, "pa_xchg_set_elem_size: %d bytes per element not supported (1, 2, 4, 8)", bytes);
  x->elem = bytes;
  return PA_OK;
}

extern "C" int pa_xchg_set_part(pa_xchg *x, int32_t k, int32_t n_snd, const int32_t *snd_ids, const int64_t *snd_ptrs, int32_t n_rcv,
                                const int32_t *rcv_ids, const int64_t *rcv_ptrs, const int64_t *rcv_src_offsets) {
  PA_CHECK(x && !x->committed, PA_ESTATE, "pa_xchg_set_part: exchange missing or already committed");
  pa_ctx *c = x->ctx;
  PA_CHECK(k >= 0 && k < c->nlocal && n_snd >= 0 && n_rcv >= 0, PA_EINVAL, "pa_xchg_set_part: bad part or counts");
  PA_CHECK((n_snd == 0 || (snd_ids && snd_ptrs)) && (n_rcv == 0 || (rcv_ids && rcv_ptrs)), PA_EINVAL, "pa_xchg_set_part: null graph arrays");
  XchgPart &p = x->parts[k];
  p = XchgPart();
  auto take = [&](int n, const int32_t *ids, const int64_t *ptrs, std::vector<int32_t> &oi, std::vector<int64_t> &op, const char *what) -> int {
    oi.resize(n);
    op.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) {
      PA_CHECK(ids[i] >= 1 && ids[i] <= c->nparts, PA_EINVAL, "pa_xchg_set_part: %s id %d out of range", what, ids[i]);
      oi[i] = ids[i] - 1;
    }
    for (int i = 0; i <= n; ++i) {
      op[i] = (n ? ptrs[i] : 1) - 1;  // JaggedArray ptrs are 1-based (src/jagged_array.jl)
      PA_CHECK(op[i] >= 0 && (i == 0 || op[i] >= op[i - 1]), PA_EINVAL, "pa_xchg_set_part: %s ptrs not monotone", what);
    }
    PA_CHECK(op[0] == 0, PA_EINVAL, "pa_xchg_set_part: %s ptrs must start at 1", what);
    return PA_OK;
  };
  PA_TRY(take(n_snd, snd_ids, snd_ptrs, p.snd_ids, p.snd_ptrs, "snd"));
  PA_TRY(take(n_rcv, rcv_ids, rcv_ptrs, p.rcv_ids, p.rcv_ptrs, "rcv"));
  if (rcv_src_offsets) {
    p.rcv_src_off.assign(rcv_src_offsets, rcv_src_offsets + n_rcv);
    p.has_src_off = true;
  }
  p.set = true;
  return PA_OK;
}

extern "C" int pa_xchg_commit(pa_xchg *x, int64_t sym_snd_len) {
  PA_CHECK(x && !x->committed, PA_ESTATE, "pa_xchg_commit: exchange missing or already committed");
  pa_ctx *c = x->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  int64_t mx = 0;
  for (int k = 0; k < c->nlocal; ++k) {
    PA_CHECK(x->parts[k].set, PA_ESTATE, "pa_xchg_commit: local part %d not set", k);
    mx = std::max(mx, x->parts[k].snd_ptrs.back());
  }
  if (sym_snd_len == 0) {
    PA_CHECK(c->nlocal == c->nparts, PA_EINVAL, "pa_xchg_commit: sym_snd_len is required when parts are remote");
    sym_snd_len = mx;
  }
  PA_CHECK(sym_snd_len >= mx, PA_EINVAL, "pa_xchg_commit: sym_snd_len smaller than a local send buffer");
  // neighbour sets for the signalling: everybody I read from and everybody who reads from me
  x->plan = new pa_plan();
  x->plan->ctx = c;
  x->plan->parts.resize(c->nlocal);
  for (int k = 0; k < c->nlocal; ++k) {
    XchgPart &p = x->parts[k];
    PlanPart &pp = x->plan->parts[k];
    std::vector<int32_t> nb(p.snd_ids);
    nb.insert(nb.end(), p.rcv_ids.begin(), p.rcv_ids.end());
    std::sort(nb.begin(), nb.end());
    nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
    PA_CHECK((int)nb.size() <= PA_MAX_NBR, PA_EINVAL, "pa_xchg_commit: more than %d neighbours", PA_MAX_NBR);
    for (int32_t q : nb) PA_CHECK(q != c->part_ids[k], PA_EINVAL, "pa_xchg_commit: part %d exchanges with itself", q + 1);
    pp.nbrs = nb;
    pp.set = true;
  }
  x->plan->committed = true;
  // the segment addressed to me inside a local sender's buffer can be derived; remote senders need it from the caller
  for (int k = 0; k < c->nlocal; ++k) {
    XchgPart &p = x->parts[k];
    if (!p.has_src_off) {
      p.rcv_src_off.assign(p.rcv_ids.size(), 0);
      for (size_t i = 0; i < p.rcv_ids.size(); ++i) {
        const int ks = c->local_of_part[p.rcv_ids[i]];
        PA_CHECK(ks >= 0, PA_EINVAL, "pa_xchg_commit: rcv_src_offsets required: part %d is not held by this process", p.rcv_ids[i] + 1);
        const XchgPart &s = x->parts[ks];
        auto it = std::find(s.snd_ids.begin(), s.snd_ids.end(), (int32_t)c->part_ids[k]);
        PA_CHECK(it != s.snd_ids.end(), PA_EINVAL, "pa_xchg_commit: graph not consistent: part %d receives from %d, which does not send to it",
                 c->part_ids[k] + 1, p.rcv_ids[i] + 1);
        const size_t j = (size_t)(it - s.snd_ids.begin());
        PA_CHECK(s.snd_ptrs[j + 1] - s.snd_ptrs[j] == p.rcv_ptrs[i + 1] - p.rcv_ptrs[i], PA_EINVAL,
                 "pa_xchg_commit: segment lengths differ between sender %d and receiver %d", p.rcv_ids[i] + 1, c->part_ids[k] + 1);
        p.rcv_src_off[i] = s.snd_ptrs[j];
      }
    }
    const PlanPart &pp = x->plan->parts[k];
    std::vector<int32_t> slot(p.rcv_ids.size());
    for (size_t i = 0; i < p.rcv_ids.size(); ++i)
      slot[i] = (int32_t)(std::lower_bound(pp.nbrs.begin(), pp.nbrs.end(), p.rcv_ids[i]) - pp.nbrs.begin());
    const size_t nr = p.rcv_ids.size();
    PA_CUDA(cudaMalloc((void **)&p.d_rcv_ptrs, (nr + 1) * sizeof(int64_t)));
    PA_CUDA(cudaMalloc((void **)&p.d_src_off, std::max<size_t>(nr, 1) * sizeof(int64_t)));
    PA_CUDA(cudaMalloc((void **)&p.d_slot, std::max<size_t>(nr, 1) * sizeof(int32_t)));
    PA_CUDA(cudaMalloc((void **)&p.d_rcv, (size_t)std::max<int64_t>(p.rcv_ptrs.back(), 1) * x->elem));
    PA_CUDA(cudaMemcpyAsync(p.d_rcv_ptrs, p.rcv_ptrs.data(), (nr + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    if (nr) {
      PA_CUDA(cudaMemcpyAsync(p.d_src_off, p.rcv_src_off.data(), nr * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
      PA_CUDA(cudaMemcpyAsync(p.d_slot, slot.data(), nr * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
    }
    PA_CUDA(cudaStreamSynchronize(c->stream));
  }
  x->snd_bytes = ((uint64_t)std::max<int64_t>(sym_snd_len, 1) * x->elem + 511) / 512 * 512;
  PA_TRY(pa_arena_alloc(c, x->snd_bytes, &x->snd_off));
  x->committed = true;
  return PA_OK;
}

extern "C" int pa_xchg_destroy(pa_xchg *x) {
  if (!x) return PA_OK;
  pa_ctx *c = x->ctx;
  cudaSetDevice(c->device);
  if (x->committed) {
    pa_before_write(c);  // the slot may be handed out again: every reader must be finished
    cudaStreamSynchronize(c->stream);
    pa_arena_free(c, x->snd_off, x->snd_bytes);
  }
  for (auto &p : x->parts) {
    cudaFree(p.d_rcv_ptrs);
    cudaFree(p.d_src_off);
    cudaFree(p.d_slot);
    cudaFree(p.d_rcv);
  }
  if (x->plan) pa_plan_destroy(x->plan);
  delete x;
  return PA_OK;
}

extern "C" int pa_xchg_upload_snd(pa_xchg *x, int32_t k, const void *data, int64_t n) {
  PA_CHECK(x && x->committed, PA_ESTATE, "pa_xchg_upload_snd: exchange missing or not committed");
  pa_ctx *c = x->ctx;
  PA_CHECK(k >= 0 && k < c->nlocal && n == x->parts[k].snd_ptrs.back() && (n == 0 || data), PA_EINVAL,
           "pa_xchg_upload_snd: %lld elements given, the send buffer holds %lld", (long long)n,
           (long long)(k >= 0 && k < c->nlocal ? x->parts[k].snd_ptrs.back() : -1));
  PA_CUDA(cudaSetDevice(c->device));
  PA_TRY(pa_before_write(c));  // nobody may still be reading the previous contents
  if (n) PA_CUDA(cudaMemcpyAsync(c->arena[k] + x->snd_off, data, (size_t)n * x->elem, cudaMemcpyHostToDevice, c->stream));
  return PA_OK;
}

/* exchange!(rcv, snd, graph): rcv segment i of every part <- the segment its source i addressed to it */
extern "C" int pa_xchg_exchange(pa_xchg *x) {
  PA_CHECK(x && x->committed, PA_ESTATE, "pa_xchg_exchange: exchange missing or not committed");
  pa_ctx *c = x->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  PA_TRY(pa_before_write(c));  // a pending "done" of an earlier collective belongs to an earlier epoch: wait for it first
  PA_TRY(pa_collective_begin(x->plan));
  for (int k = 0; k < c->nlocal; ++k) {
    XchgPart &p = x->parts[k];
    const int nr = (int)p.rcv_ids.size();
    if (!nr || p.rcv_ptrs.back() == 0) continue;
    int64_t longest = 0;
    for (int i = 0; i < nr; ++i) longest = std::max(longest, p.rcv_ptrs[i + 1] - p.rcv_ptrs[i]);
    PeerPtrs peers;
    const PlanPart &pp = x->plan->parts[k];
    for (size_t i = 0; i < PA_MAX_NBR; ++i) peers.p[i] = i < pp.nbrs.size() ? (double *)(c->peer_base[pp.nbrs[i]] + x->snd_off) : nullptr;
    for (size_t i = 0; i < pp.nbrs.size(); ++i)
      PA_CHECK(c->peer_base[pp.nbrs[i]], PA_ESTATE, "part %d's arena was never imported (pa_ctx_arena_import)", pp.nbrs[i] + 1);
    const dim3 grid((unsigned)std::min<int64_t>((longest + 255) / 256, 148 * 4), (unsigned)nr);
    switch (x->elem) {
      case 8: k_exchange<<<grid, 256, 0, c->stream>>>((unsigned long long *)p.d_rcv, p.d_rcv_ptrs, p.d_slot, p.d_src_off, peers); break;
      case 4: k_exchange<<<grid, 256, 0, c->stream>>>((uint32_t *)p.d_rcv, p.d_rcv_ptrs, p.d_slot, p.d_src_off, peers); break;
      case 2: k_exchange<<<grid, 256, 0, c->stream>>>((uint16_t *)p.d_rcv, p.d_rcv_ptrs, p.d_slot, p.d_src_off, peers); break;
      default: k_exchange<<<grid, 256, 0, c->stream>>>((uint8_t *)p.d_rcv, p.d_rcv_ptrs, p.d_slot, p.d_src_off, peers); break;
    }
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  return pa_collective_end(x->plan);
}

extern "C" int pa_xchg_download_rcv(pa_xchg *x, int32_t k, void *data, int64_t n) {
  PA_CHECK(x && x->committed, PA_ESTATE, "pa_xchg_download_rcv: exchange missing or not committed");
  pa_ctx *c = x->ctx;
  PA_CHECK(k >= 0 && k < c->nlocal && n == x->parts[k].rcv_ptrs.back() && (n == 0 || data), PA_EINVAL,
           "pa_xchg_download_rcv: %lld elements asked, the receive buffer holds %lld", (long long)n,
           (long long)(k >= 0 && k < c->nlocal ? x->parts[k].rcv_ptrs.back() : -1));
  PA_CUDA(cudaSetDevice(c->device));
  if (n) PA_CUDA(cudaMemcpyAsync(data, x->parts[k].d_rcv, (size_t)n * x->elem, cudaMemcpyDeviceToHost, c->stream));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  return pa_check_device_error(c);
}

// ------------------------------------------------------------------ reduce(op, own_values) per part
// OP: PA_OP_SUM, PA_OP_MAX, PA_OP_MIN, PA_OP_ABSSUM (norm(.,1)), PA_OP_ABSMAX, PA_OP_ABSPOW (sum |x|^p = norm(.,p)^p)
template <int OP>
__device__ __forceinline__ double red_neutral() {
  return OP == PA_OP_MAX ? -INFINITY : (OP == PA_OP_MIN ? INFINITY : 0.0);  // neutral_element, src/p_vector.jl:1170-1175
}
template <int OP>
__device__ __forceinline__ double red_map(double x, double p) {
  return OP == PA_OP_ABSSUM || OP == PA_OP_ABSMAX ? fabs(x) : (OP == PA_OP_ABSPOW ? pow(fabs(x), p) : x);
}
template <int OP>
__device__ __forceinline__ double red_op(double a, double b) {
  return OP == PA_OP_MAX || OP == PA_OP_ABSMAX ? fmax(a, b) : (OP == PA_OP_MIN ? fmin(a, b) : a + b);
}

template <int OP>
__global__ void __launch_bounds__(PA_RED_THREADS)
    k_reduce_op(const double *__restrict__ x, int64_t n, const int32_t *__restrict__ idx, double p, double *blockpart, unsigned *ticket,
                double *out) {
  __shared__ double sm[PA_RED_THREADS / 32];
  __shared__ bool last;
  double acc = red_neutral<OP>();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc = red_op<OP>(acc, red_map<OP>(x[idx ? idx[i] : i], p));
  auto block_fold = [&](double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = red_op<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : red_neutral<OP>();
    if (threadIdx.x < 32) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) r = red_op<OP>(r, __shfl_xor_sync(0xffffffffu, r, o));
    }
    return r;  // valid in thread 0
  };
  const double s = block_fold(acc);
  if (threadIdx.x == 0) {
    blockpart[blockIdx.x] = s;
    __threadfence();
    const unsigned t = atomicInc(ticket, gridDim.x - 1);  // wraps to 0: self resetting
    last = t == gridDim.x - 1;
  }
  __syncthreads();
  if (last) {  // the last CTA folds the per-CTA partials in a fixed order: deterministic run to run
    __threadfence();
    double a = red_neutral<OP>();
    for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) a = red_op<OP>(a, __ldcg(blockpart + i));
    a = block_fold(a);
    if (threadIdx.x == 0) *out = a;
  }
}

/* out[k] = reduce(op, own_values(x)[k]; init = neutral) for every local part k (the first map of reduce(op, ::PVector),
 * src/p_vector.jl:1178-1183); the reduction over parts is the caller's (reduce(op, b), :1182).  Synchronises. */
extern "C" int pa_vec_reduce_parts(const pa_vec *x, int32_t op, double p, double *out) {
  PA_CHECK(x && out, PA_EINVAL, "pa_vec_reduce_parts: null argument");
  PA_CHECK(op >= PA_OP_SUM && op <= PA_OP_ABSPOW && op != PA_OP_INSERT, PA_EINVAL, "pa_vec_reduce_parts: unknown operation %d", op);
  pa_ctx *c = x->plan->ctx;
  PA_CUDA(cudaSetDevice(c->device));
  for (int k = 0; k < c->nlocal; ++k) {
    const PlanPart &pp = x->plan->parts[k];
    const int64_t n = pp.n_own;
    int grid = (int)std::min<int64_t>(std::max<int64_t>((n + PA_RED_THREADS * 8 - 1) / (PA_RED_THREADS * 8), 1), PA_RED_BLOCKS);
    const int32_t *idx = pp.prefix ? nullptr : pp.d_own_to_local;
    double *bp = c->d_blockpart + (size_t)k * PA_RED_BLOCKS, *o = c->d_partial + k;
    unsigned *tk = c->d_ticket + k;
#define PA_RED_LAUNCH(OP) k_reduce_op<OP><<<grid, PA_RED_THREADS, 0, c->stream>>>(x->d[k], n, idx, p, bp, tk, o)
    switch (op) {
      case PA_OP_SUM: PA_RED_LAUNCH(PA_OP_SUM); break;
      case PA_OP_MAX: PA_RED_LAUNCH(PA_OP_MAX); break;
      case PA_OP_MIN: PA_RED_LAUNCH(PA_OP_MIN); break;
      case PA_OP_ABSSUM: PA_RED_LAUNCH(PA_OP_ABSSUM); break;
      case PA_OP_ABSMAX: PA_RED_LAUNCH(PA_OP_ABSMAX); break;
      default: PA_RED_LAUNCH(PA_OP_ABSPOW); break;
    }
#undef PA_RED_LAUNCH
    c->launches++;
  }
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaMemcpyAsync(out, c->d_partial, c->nlocal * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PA_CUDA(cudaStreamSynchronize(c->stream));
  return pa_check_device_error(c);
}

// ------------------------------------------------------------------ spmv! / spmtv! on ONE local matrix
// The reference's four local kernels (src/sparse_utils.jl:609-690) for every storage it tests
// (test/sparse_utils_tests.jl:14-45,113-118: CSC / CSR{1} / CSR{0} x (Float64,Int) / (Float32,Int32)):
//   kind 0 = spmv_csr!(b,x,ptr,idx,val)   b[i] = sum_p val[p]*x[idx[p]], p in storage order            (:649-669)
//   kind 1 = spmv_csc!(b,x,ptr,idx,val)   b = 0; for j, p: b[idx[p]] += val[p]*x[j]                    (:671-690)
// spmv!(CSR) = kind 0 on (rowptr,colval); spmtv!(CSR) = kind 1 on (rowptr,colval); spmv!(CSC) = kind 1 on (colptr,rowval);
// spmtv!(CSC) = kind 0 on (colptr,rowval) (:617-647).  Kind 1 is executed as a gather: the entries are transposed on the host
// with a stable counting sort, so that every b[i] receives its contributions in the order of the reference's scatter loop
// (ascending j, then storage order) -- same bits.  Separate multiply and add in the value type, like Julia's bi += aij*xj.
template <typename T>
__device__ __forceinline__ T mul_add_seq(T acc, T a, T x);
template <>
__device__ __forceinline__ double mul_add_seq<double>(double acc, double a, double x) { return __dadd_rn(acc, __dmul_rn(a, x)); }
template <>
__device__ __forceinline__ float mul_add_seq<float>(float acc, float a, float x) { return __fadd_rn(acc, __fmul_rn(a, x)); }

template <typename T>
__global__ void k_local_spmv(const int64_t *__restrict__ ptr, const int64_t *__restrict__ idx, const T *__restrict__ val, const T *__restrict__ x,
                             T *__restrict__ b, int64_t nb) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += (int64_t)gridDim.x * blockDim.x) {
    T acc = (T)0;
    for (int64_t p = ptr[i]; p < ptr[i + 1]; ++p) acc = mul_add_seq<T>(acc, val[p], x[idx[p]]);
    b[i] = acc;
  }
}

template <typename T>
static int local_spmv_run(pa_ctx *c, const std::vector<int64_t> &ptr, const std::vector<int64_t> &idx, const T *val, const std::vector<int64_t> &vperm,
                          const T *x, int64_t nx, T *b, int64_t nb) {
  const int64_t nnz = (int64_t)idx.size();
  std::vector<T> v(nnz);
  for (int64_t p = 0; p < nnz; ++p) v[p] = val[vperm.empty() ? p : vperm[p]];
  int64_t *d_ptr = nullptr, *d_idx = nullptr;
  T *d_val = nullptr, *d_x = nullptr, *d_b = nullptr;
  PA_CUDA(cudaMalloc((void **)&d_ptr, (nb + 1) * sizeof(int64_t)));
  PA_CUDA(cudaMalloc((void **)&d_idx, std::max<int64_t>(nnz, 1) * sizeof(int64_t)));
  PA_CUDA(cudaMalloc((void **)&d_val, std::max<int64_t>(nnz, 1) * sizeof(T)));
  PA_CUDA(cudaMalloc((void **)&d_x, std::max<int64_t>(nx, 1) * sizeof(T)));
  PA_CUDA(cudaMalloc((void **)&d_b, std::max<int64_t>(nb, 1) * sizeof(T)));
  PA_CUDA(cudaMemcpyAsync(d_ptr, ptr.data(), (nb + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
  if (nnz) {
    PA_CUDA(cudaMemcpyAsync(d_idx, idx.data(), nnz * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    PA_CUDA(cudaMemcpyAsync(d_val, v.data(), nnz * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  }
  if (nx) PA_CUDA(cudaMemcpyAsync(d_x, x, nx * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  if (nb) {
    k_local_spmv<T><<<(unsigned)std::min<int64_t>((nb + 255) / 256, 148 * 8), 256, 0, c->stream>>>(d_ptr, d_idx, d_val, d_x, d_b, nb);
    c->launches++;
    PA_CUDA(cudaGetLastError());
    PA_CUDA(cudaMemcpyAsync(b, d_b, nb * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
  }
  PA_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_ptr); cudaFree(d_idx); cudaFree(d_val); cudaFree(d_x); cudaFree(d_b);
  return PA_OK;
}

extern "C" int pa_local_spmv(pa_ctx *c, int32_t kind, int32_t index_base, int32_t idx_bits, int32_t val_bits, int64_t ncomp, int64_t nb,
                             const void *ptr, const void *idx, const void *val, const void *x, int64_t nx, void *b) {
  PA_CHECK(c && ptr && b && (kind == 0 || kind == 1) && (index_base == 0 || index_base == 1) && (idx_bits == 32 || idx_bits == 64) &&
               (val_bits == 32 || val_bits == 64) && ncomp >= 0 && nb >= 0 && nx >= 0,
           PA_EINVAL, "pa_local_spmv: bad arguments");
  // dimension asserts of spmv!/spmtv! (src/sparse_utils.jl:618-621): kind 0 writes one entry per compressed row
  PA_CHECK(kind == 1 || nb == ncomp, PA_EINVAL, "pa_local_spmv: length(b) = %lld, the matrix has %lld rows", (long long)nb, (long long)ncomp);
  PA_CHECK(kind == 0 || nx == ncomp, PA_EINVAL, "pa_local_spmv: length(x) = %lld, the matrix has %lld columns", (long long)nx, (long long)ncomp);
  PA_CUDA(cudaSetDevice(c->device));
  auto rd = [&](const void *p, int64_t i) { return idx_bits == 64 ? ((const int64_t *)p)[i] : (int64_t)((const int32_t *)p)[i]; };
  std::vector<int64_t> p0(ncomp + 1);
  for (int64_t i = 0; i <= ncomp; ++i) p0[i] = rd(ptr, i) - index_base;
  PA_CHECK(p0[0] == 0, PA_EINVAL, "pa_local_spmv: ptr does not start at the index base");
  for (int64_t i = 0; i < ncomp; ++i) PA_CHECK(p0[i + 1] >= p0[i], PA_EINVAL, "pa_local_spmv: ptr not monotone");
  const int64_t nnz = p0[ncomp];
  PA_CHECK(nnz == 0 || (idx && val && x), PA_EINVAL, "pa_local_spmv: null array");
  std::vector<int64_t> i0(nnz);
  const int64_t bound = kind == 0 ? nx : nb;
  for (int64_t p = 0; p < nnz; ++p) {
    i0[p] = rd(idx, p) - index_base;
    PA_CHECK(i0[p] >= 0 && i0[p] < bound, PA_EINVAL, "pa_local_spmv: index %lld out of range", (long long)(i0[p] + index_base));
  }
  std::vector<int64_t> gptr, gidx, perm;
  if (kind == 0) {
    gptr = p0;
    gidx = i0;
  } else {  // stable transpose: entries of b[i] in ascending (j, p) = the order of the scatter loop
    gptr.assign(nb + 1, 0);
    for (int64_t p = 0; p < nnz; ++p) gptr[i0[p] + 1]++;
    for (int64_t i = 0; i < nb; ++i) gptr[i + 1] += gptr[i];
    gidx.resize(nnz);
    perm.resize(nnz);
    std::vector<int64_t> at(gptr.begin(), gptr.end() - 1);
    for (int64_t j = 0; j < ncomp; ++j)
      for (int64_t p = p0[j]; p < p0[j + 1]; ++p) {
        const int64_t q = at[i0[p]]++;
        gidx[q] = j;
        perm[q] = p;
      }
  }
  if (val_bits == 64) return local_spmv_run<double>(c, gptr, gidx, (const double *)val, perm, (const double *)x, nx, (double *)b, nb);
  return local_spmv_run<float>(c, gptr, gidx, (const float *)val, perm, (const float *)x, nx, (float *)b, nb);
}
