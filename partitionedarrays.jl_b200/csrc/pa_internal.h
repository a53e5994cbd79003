// Internal structures of libpa_b200.so (not part of the ABI; see include/pa_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/pa_b200.h"

#define PA_MAX_NBR 32   // neighbours per part whose base pointers travel by value in kernel params
#define PA_NSCAL 16     // device scalar slots per context
#define PA_RED_BLOCKS 1184  // 148 SMs x 8 resident CTAs: grid of every BLAS-1 reduction
#define PA_RED_THREADS 256
#define PA_DOT_PARTS 4096  // >= persistent SpMV grid (SMs x resident CTAs)
#define PA_MAT_PAD 48      // entries of tail padding behind colval/nzval: TMA copies are 16-byte granular, the batch
                           // Gauss-Seidel kernel reads 32 entry slots per row without predicates

void pa_set_error(const char *fmt, ...);
int pa_cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define PA_CUDA(call)                                                    \
  do {                                                                   \
    cudaError_t _e = (call);                                             \
    if (_e != cudaSuccess) return pa_cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)
#define PA_CHECK(cond, code, ...) \
  do {                            \
    if (!(cond)) {                \
      pa_set_error(__VA_ARGS__);  \
      return (code);              \
    }                             \
  } while (0)
#define PA_TRY(call)          \
  do {                        \
    int _r = (call);          \
    if (_r != PA_OK) return _r; \
  } while (0)

// scalar slots (device resident; CG never brings them to the host inside the loop)
enum { S_TMP = 0, S_RHO0 = 1, S_RHO1 = 2, S_UC = 3, S_NRM2 = 4, S_NRM0 = 5 };

struct PeerPtrs {
  double *p[PA_MAX_NBR];
};
struct FlagPtrs {
  unsigned long long *p[PA_MAX_NBR];
};

// ---- folded scalar all-reduce (one part per process, or a job of one part): the LAST CTA of the producing kernel
// pushes the part's value + a sequence number into a double-buffered slot of every part's arena header (system-scope
// stores over NVLink); every CTA of the consuming kernel waits for all P flags in its own HBM and adds the P values in
// PART ORDER (the order of the reference's sequential sum over parts, src/primitives.jl:693-698).  No launch of its own.
struct RedPush {
  unsigned long long *flag[PA_MAX_NBR];  // per destination part: red_flag[2][nparts] in its header
  double *val[PA_MAX_NBR];               // per destination part: red_val[2][nparts]
  unsigned long long *epoch;             // reductions posted so far (device resident: graph replay safe)
  int me, nparts;
};
struct RedWait {
  const unsigned long long *flag;  // my header
  const double *val;
  const unsigned long long *epoch;
  int nparts;
  int *err;
};
// neighbours' "done reading your vectors" flags in my header (WAR guard of the next vector write)
struct DoneWait {
  const unsigned long long *flag[PA_MAX_NBR];
  const unsigned long long *epoch;
  int n;
  int *err;
};

struct CgWork;

struct pa_ctx {
  int nparts = 0, nlocal = 0, device = 0;
  std::vector<int> part_ids;       // 0-based global ids of local parts
  std::vector<int> local_of_part;  // global part -> local k or -1
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t side = nullptr;            // consistent!(x) overlapped with the own-block product
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  uint64_t arena_bytes = 0, hdr_bytes = 0, bump = 0;
  std::vector<char *> arena;      // per local part (device memory, cudaMalloc, IPC exportable)
  std::vector<char *> peer_base;  // per global part: arena base in this process' address space
  std::vector<bool> peer_ipc;
  std::map<uint64_t, uint64_t> freeblocks;  // symmetric arena: free blocks (offset -> bytes), coalesced; first fit
  // reductions
  double *d_scal = nullptr;       // [PA_NSCAL]
  double *d_partial = nullptr;    // [nlocal]
  double *d_blockpart = nullptr;  // [nlocal][PA_RED_BLOCKS]
  unsigned *d_ticket = nullptr;   // [nlocal]
  double *h_scal = nullptr;       // pinned [PA_NSCAL]
  unsigned long long *d_epoch = nullptr;  // [nlocal] device-side op counters (graph-replay safe)
  unsigned long long *d_red_epoch = nullptr;  // sequence number of the peer-memory scalar all-reduce
  int *d_err = nullptr;           // device error flag (spin timeout)
  int *h_err = nullptr;           // pinned mirror
  // NCCL (dlopen)
  void *nccl_comm = nullptr;
  int rank = 0, world = 1;
  int64_t launches = 0;
  std::vector<pa_plan *> pending_done;  // plans whose neighbours still have to report "done reading"
  std::map<std::string, int64_t> knobs;
  unsigned *d_cons_ticket = nullptr;    // [2] last-CTA tickets of the fused signal+gather+done kernels
  std::vector<CgWork *> cg_work;        // cached CG workspaces (work vectors, history, captured graph) by (A, x, b)
  uint64_t next_uid = 1;                // identity of matrices / vectors (cache keys survive address reuse)
  double cg_timing[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // last PA_CG_TIMING solve: ddot, waxpby, spmv, precond, total (ms), iterations
};

struct PlanPart {
  int64_t n_local = 0, n_own = 0, n_ghost = 0;
  bool prefix = true;  // own = [0,n_own), ghost = [n_own,n_local)
  std::vector<int32_t> own_to_local, ghost_to_local;  // 0-based (empty when prefix)
  std::vector<int32_t> nbr_snd, snd_ptrs, snd_lids, snd_rlids;  // 0-based
  std::vector<int32_t> nbr_rcv, rcv_ptrs, rcv_lids, rcv_rlids;
  bool has_snd_rl = false, has_rcv_rl = false, set = false;
  uint64_t signature = 0;  // hash of the whole local layout + exchange lists (pa_plan_commit): equal <=> identical partition
  // neighbour union (sorted global part ids) and slot lookup
  std::vector<int32_t> nbrs;
  // device tables
  int32_t *d_own_to_local = nullptr, *d_ghost_to_local = nullptr;
  int32_t *d_ghost_lid = nullptr;   // [n_ghost entries of the snd lists] local id written by consistent!
  int32_t *d_ghost_slot = nullptr;  // neighbour slot (index into nbrs) owning it
  int32_t *d_ghost_rlid = nullptr;  // local id on the owner
  // fused lookup by ghost id (prefix layout): slot / remote lid of ghost g = lid - n_own
  int32_t *d_gslot_by_gid = nullptr, *d_grlid_by_gid = nullptr;
  int64_t n_cons = 0;  // entries of the consistent! gather (= total snd entries)
  // assemble!: destinations (own lids) with their contributions grouped in neighbour order
  int32_t *d_asm_dst = nullptr, *d_asm_ptr = nullptr, *d_asm_slot = nullptr, *d_asm_rlid = nullptr;
  int64_t n_asm_dst = 0;
  // the same destinations as a bitmap over the local ids (bit l set: a neighbour reads/writes local entry l): lets a
  // streaming kernel treat the boundary entries separately (exchange fused into the producer of the vector)
  uint32_t *d_bnd_bitmap = nullptr;
};

struct pa_plan {
  pa_ctx *ctx = nullptr;
  std::vector<PlanPart> parts;
  int64_t sym_n_local = 0;
  uint64_t vec_bytes = 0;
  bool committed = false;
};

struct pa_vec {
  uint64_t uid = 0;
  pa_plan *plan = nullptr;
  uint64_t offset = 0;            // symmetric arena offset
  std::vector<double *> d;        // per local part
};

struct MatPart {
  int64_t nrows = 0, ncols = 0, nnz = 0;
  bool ptr64 = false;
  void *d_rowptr = nullptr;  // int32 or int64, 0-based, nrows+1
  int32_t *d_colval = nullptr;  // 0-based local column ids (column plan numbering)
  double *d_nzval = nullptr;
  bool set = false;
  int rows_per_cta = 256;
  std::map<int, int64_t> tile_nnz;
  // rows that reference ghost columns (ghost block of the split product)
  bool ghost_scanned = false, ghost_tail_ok = true, tma_ok = true;
  int64_t n_grows = 0;
  int32_t *d_grows = nullptr;
  double *d_dotpart = nullptr;  // per-CTA partials of the fused dot epilogue
  unsigned *d_dot_ticket = nullptr;  // last-CTA ticket of the folded dot epilogue
  unsigned long long *d_arrive = nullptr;  // fused consistent! (SpMV MODE 4): CTAs counted in, over all launches
  std::map<int, unsigned char *> tile_ghost;  // MODE 4: per tile size, "tile holds a ghost column" flags
  int64_t arrive_grid = 0;
  // row patterns (column stream compression of the TMA SpMV): pat[row] = id of the row's (length, column - row) tuple in
  // ptab, 255 = the row reads its columns from colval.  0 = not tried, 1 = available, -1 = does not pay for this matrix
  int pat_state = 0;
  unsigned char *d_pat = nullptr;
  int32_t *d_ptab = nullptr;  // [npat][pat_w]
  int npat = 0, pat_w = 0, pat_len0 = 0;  // pat_len0: length of the most frequent pattern (id 0)
  // COO pattern cache (the reference's K of sparse_matrix(...; reuse=true)): sorted permutation + segment starts
  int32_t *d_coo_perm = nullptr, *d_coo_seg = nullptr;
  unsigned char *d_coo_valid = nullptr;
  int64_t n_coo = 0;  // max nnz of a ROWS-row tile, by ROWS (TMA stage sizing)
};

struct pa_mat {
  uint64_t uid = 0;
  pa_ctx *ctx = nullptr;
  pa_plan *rows = nullptr, *cols = nullptr;
  std::vector<MatPart> parts;
  bool committed = false;
  bool subassembled = false;  // some part stores its ghost rows too (n_local rows): mul! ends with assemble!(c)
  pa_mat *T = nullptr;  // lazily built local transposes (transpose mul!)
};

// coefficient of a vector update: immediate, or sign * num/den read from device scalars (the CG
// loop never brings alpha/beta to the host)
struct Coef {
  double imm;
  const double *num, *den;
  double sign;
};
static inline Coef coef_imm(double a) { return Coef{a, nullptr, nullptr, 1.0}; }
static inline Coef coef_ratio(const double *num, const double *den, double sign) { return Coef{0.0, num, den, sign}; }

// ---- runtime helpers shared between translation units ----
bool pa_fold_ok(pa_ctx *ctx);                      // folded reductions / signalling usable (one local part, peers mapped)
RedPush pa_red_push(pa_ctx *ctx);
RedWait pa_red_wait(pa_ctx *ctx);
DoneWait pa_done_wait(pa_plan *plan);              // neighbours of local part 0
int pa_consistent_sync(pa_vec *v);                 // signal + wait + gather + done in ONE kernel (one local part)
void pa_cg_drop_work(pa_ctx *ctx, const pa_mat *A, const pa_plan *plan, bool all);  // invalidate cached CG workspaces
int pa_collective_begin(pa_plan *plan);   // publish my data + wait for the neighbours'
int pa_collective_end(pa_plan *plan);     // tell the neighbours I am done reading theirs
int pa_sync_flags(pa_plan *plan, FlagPtrs *arrive_dst, FlagPtrs *arrive_src, FlagPtrs *done_dst, int *n);
void pa_mark_pending_done(pa_plan *plan);  // the neighbours' "done" for the current epoch has still to be waited for
int pa_before_write(pa_ctx *ctx);         // wait until nobody is still reading my vectors
PeerPtrs pa_peer_ptrs(const pa_vec *v, int k);
int pa_arena_alloc(pa_ctx *ctx, uint64_t bytes, uint64_t *off);  // symmetric offset (same sequence of calls on every process)
void pa_arena_free(pa_ctx *ctx, uint64_t off, uint64_t bytes);
int pa_reduce_finish(pa_ctx *ctx, double *d_out);  // sum local partials (+ NCCL all-reduce) into *d_out
int pa_read_scalars(pa_ctx *ctx, int first, int count, double *out);  // synchronises
int pa_check_device_error(pa_ctx *ctx);
int64_t pa_knob(pa_ctx *ctx, const char *key, int64_t dflt);

// kernels launched from several files
int pa_launch_consistent(pa_vec *v);  // gather kernel only (no signalling)
int pa_waxpby_dev(pa_vec *w, Coef ca, const pa_vec *x, Coef cb, const pa_vec *y);
int pa_reduce_dev_to(const pa_vec *x, const pa_vec *y, int mode, double *d_out);  // mode 0 dot, 1 sumsq, 2 sum
// fold != 0: the dot epilogue's last CTA pushes the part's value to the peers (RedPush) instead of k_sum_parts + all-reduce
int pa_spmv_local(pa_mat *A, pa_vec *x, pa_vec *y, double alpha, double beta, int mode, const pa_vec *dotw, double *d_out, int fold = 0);
int pa_spmv_dot(pa_mat *A, pa_vec *x, pa_vec *y, double alpha, double beta, uint32_t flags, const pa_vec *dotw, double *d_out, int fold = 0);
bool pa_spmv_dot_foldable(pa_mat *A, pa_vec *x);
void pa_mat_drop_transpose(pa_mat *A);
