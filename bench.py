#!/usr/bin/env python
"""bench.py — HPCG-style CG + SpMV benchmark of the PSparseMatrix x PVector hot path on B200.

Workload (BASELINE.json configs[1] at N=1, configs[2] at N=8): gallery 7-pt Laplacian, 512^3 rows per GPU,
fp64 values / int32 columns, weak scaling over a (npx,npy,npz) part grid, one part per GPU / process.
A "step" is one ref_cg!(x,A,b; maxiter=ITERS, Pl=Identity) call (HPCG/src/ref_cg.jl:119-134) from x0=0.

  parity_check = BEFORE anything is timed, at every N: A*ones == rhs exactly; for a hash-valued x every own row that
           touches a ghost column (all faces/edges/corners shared with another part) plus 2000 interior rows is compared
           BIT FOR BIT with the stencil definition evaluated on the host, for every mul! schedule; the ghost slots left by
           consistent! equal the hash of their global id.  No value is printed unless the gate passes on every rank.
  value  = HPCG-model GFLOP/s of the CG loop, whole job, operands resident in HBM
           ((2*nnz + 12*n) flop per iteration — HPCG/src/report_results.jl:27-29 — x iterations / time)
  e2e    = same metric through the public API with HOST buffers: every step uploads b from pinned host memory, sets
           x0 = 0 (fill!), solves, and downloads x and the residual history inside the timed region; the transfers of
           neighbouring steps overlap with the running solve (two buffer pairs, copy streams); the strictly serial
           upload -> solve -> download figure is printed beside it (e2e.serial_value)
  roofline = the SpMV kernel (dominant): algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline = the CPU oracle (C twin of the reference loops, one part per host thread) on a bounded sample
Secondary sections (same JSON line): hpcg27 (27-pt 512^3 per GPU: SpMV, CG, 4-level MG-preconditioned CG — configs[3],
weak, every N), strong (global 512^3 split over the parts, N > 1), fem_c5 (configs[4]: Q1 FEM assembly -> mul! -> CG,
N = 4), reference_ops (the op-for-op CG schedule of ref_cg.jl beside the fused one).

`--impl reference` times the reference's own CPU algorithm (the oracle; the Julia reference cannot run here)."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--grid", dest="n", type=int, default=512, help="grid edge per GPU")
    ap.add_argument("--kind", type=int, default=7, choices=[7, 27])
    ap.add_argument("--iters", type=int, default=50, help="CG iterations per step (HPCG ref_max_iters)")
    ap.add_argument("--spmv-reps", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-rows", type=int, default=1 << 24, help="rows of the CPU sample of the cpu_baseline leg")
    ap.add_argument("--no-hpcg27", action="store_true")
    ap.add_argument("--strong", action="store_true", help="headline in strong-scaling mode: the GLOBAL grid is n^3, split over the parts")
    ap.add_argument("--no-strong", action="store_true", help="skip the secondary strong-scaling section (N > 1)")
    ap.add_argument("--mg", action="store_true", help="(default on) HPCG multigrid-preconditioned CG section (27-pt 512^3, 4 levels)")
    ap.add_argument("--no-mg", action="store_true", help="skip the multigrid-preconditioned CG section")
    ap.add_argument("--fem", action="store_true", help="force the FEM (C5) section (default: only at N = 4)")
    ap.add_argument("--no-fem", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="debugging only: skip the parity gate (the line then says so)")
    return ap.parse_args()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def spmv_bytes(n_rows, nnz, n_cols_local):
    ptr = 4 if nnz < 2 ** 31 else 8
    return nnz * 12 + (n_rows + 1) * ptr + 8 * n_rows + 8 * n_cols_local


def host_threads():
    """Host cores this process may use (torchrun exports OMP_NUM_THREADS=1: the affinity mask is what counts)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed regions."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if not any(a <= t <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except Exception:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU oracle arm
_CPU_CACHE = {}


def cpu_cg_sample(kind, grid, iters, threads=None):
    """The oracle CG (one part per host thread) on z-slabs of the same operator: grid = (nx, ny, nz_total)."""
    key = (kind, tuple(grid), threads)
    if key not in _CPU_CACHE:
        _CPU_CACHE[key] = _cpu_build(kind, grid, threads)
    mats, plan, bvals, gn, P = _CPU_CACHE[key]
    return _cpu_run(kind, mats, plan, bvals, gn, P, iters)


def _cpu_build(kind, grid, threads):
    from oracle import c_oracle, pa_oracle as o

    assert c_oracle.available()
    P = threads or host_threads()
    nx, ny, nzt = (int(g) for g in grid)
    P = max(1, min(P, nzt // 2))
    gn = (nx, ny, nzt)
    part = o.uniform_partition((1, 1, P), gn)
    mats, cols, bvals = [], [], []
    for ind in part:
        lo = [r[0] - 1 for r in ind.box]
        hi = [r[1] for r in ind.box]
        gh = []
        # ghosts of a z-slab = the adjacent planes, in the order union_ghost meets them in the generator's column list
        for zface in sorted({lo[2], hi[2] - 1}):
            ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
            ix, iy = ix.reshape(-1), iy.reshape(-1)
            offs = ([(0, 0, -1), (0, 0, 1)] if kind == 7 else [(sx, sy, sz) for sz in (-1, 0, 1) for sy in (-1, 0, 1) for sx in (-1, 0, 1)])
            cand = []
            for sx, sy, sz in offs:
                cx, cy, cz = ix + sx, iy + sy, zface + sz
                ok = (cx >= 0) & (cx < nx) & (cy >= 0) & (cy < ny) & (cz >= 0) & (cz < gn[2]) & ((cz < lo[2]) | (cz >= hi[2]))
                cand.append(np.where(ok, cx + nx * (cy + ny * cz) + 1, 0))
            gh.append(np.stack(cand, 1).reshape(-1))
        gh = np.concatenate(gh) if gh else np.zeros(0, np.int64)
        gh = gh[gh > 0]
        _, first = np.unique(gh, return_index=True)
        gh = gh[np.sort(first)]
        owners = o.find_owner(part, [gh])[0]
        cols.append(o.LocalIndices(ind.n_global, ind.part, np.concatenate([ind.local_to_global, gh]), np.concatenate([ind.local_to_owner, owners]),
                                   box=ind.box, grid=gn, parts_per_dir=(1, 1, P)))
        rp, cv, nz, b = c_oracle.stencil_csr(kind, gn, lo, hi, gh - 1)
        mats.append((ind.n_own, ind.n_own + len(gh), rp, cv, nz))
        bl = np.zeros(ind.n_own + len(gh)); bl[: ind.n_own] = b
        bvals.append(bl)
    plan = o.assembly_plan(cols)
    prob = c_oracle.CGProblem(mats, plan, bvals, [np.zeros(m[1]) for m in mats])
    prob.cg(2, 0.0)  # warm-up (page faults, thread pool)
    return mats, plan, bvals, gn, P


def _cpu_run(kind, mats, plan, bvals, gn, P, iters):
    from oracle import c_oracle

    prob2 = c_oracle.CGProblem(mats, plan, bvals, [np.zeros(m[1]) for m in mats])
    t0 = time.perf_counter()
    it, hist, tm = prob2.cg(iters, 0.0)
    dt = time.perf_counter() - t0
    n = sum(m[0] for m in mats)
    nnz = sum(len(m[4]) for m in mats)
    flops = (2 * nnz + 12 * n) * it
    t_spmv = prob2.time_spmv(3) / 3
    return {"gflops": flops / dt / 1e9, "iters_per_sec": it / dt, "cores": P, "rows": n, "nnz": nnz, "iters": it, "seconds": dt,
            "spmv_gflops": 2 * nnz / t_spmv / 1e9, "rel_residual": float(hist[-1] / hist[0]),
            "sample": f"{kind}-pt {gn[0]}x{gn[1]}x{gn[2]} ({n} rows, {nnz} nnz) on {P} parts/threads, {it} CG iterations"}


def run_reference(args):
    """The reference's CPU algorithm (C twin of the Julia loops, one part per host thread — the MPIArray model with one
    rank per core) on the GPU arm's single-GPU operator: {kind}-pt n^3, ITERS iterations per step, when host RAM allows;
    else a z-slab sample of it.  Under torchrun rank 0 alone runs, with ALL host cores (torchrun's OMP_NUM_THREADS=1 is
    overridden: the thread count is the affinity mask)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(cores)  # before libgomp is loaded by the oracle library
    os.environ.pop("OMP_THREAD_LIMIT", None)
    n = args.n
    try:
        import psutil

        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    nnz_est = (7 if args.kind == 7 else 27) * n ** 3
    need = nnz_est * 12 + 9 * n ** 3 * 8 * 1.5  # CSR + rowptr + 5 vectors, with head-room for the build
    full = avail > 2.5 * need
    grid = (n, n, n) if full else (n, n, max(2 * cores, int((1 << 24) // (n * n))))
    iters = args.iters
    t_all = time.perf_counter()
    vals, last = [], None
    for s in range(args.warmup + args.steps):
        last = cpu_cg_sample(args.kind, grid, iters, cores)
        if s >= args.warmup:
            vals.append(last)
        elapsed = time.perf_counter() - t_all
        if elapsed > 200 and s < args.warmup:  # slow host: skip the remaining warm-up steps
            args.warmup = s + 1
        if elapsed > 280 and len(vals) >= 1:
            break
    secs = sum(v["seconds"] for v in vals)
    flops = sum(v["gflops"] * v["seconds"] for v in vals)
    g = flops / secs
    same = full and args.gpus == 1
    line = {"impl": "reference", "metric": "hpcg_cg_gflops", "value": g, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": len(vals),
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(vals), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"gallery {args.kind}-pt Laplacian CG (Pl=Identity), CPU oracle: {last['sample']}",
                       "same_operator_as_gpu_arm_at_n1": bool(full),
                       "note": ("the whole single-GPU operator of the GPU arm" if same else
                                ("one GPU's share of the GPU arm's weak-scaled operator (the CPU arm does not grow with N)" if full else
                                 "a z-slab sample of the GPU arm's operator (host RAM)"))},
            "cg_iters_per_sec": last["iters_per_sec"], "spmv_gflops": last["spmv_gflops"],
            "cpu_baseline": {"value": g, "unit": "GFLOP/s", "cores": last["cores"], "kind": "port", "sample": last["sample"]},
            "e2e": {"value": g, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = CPU restatement of the Julia loops (oracle/pa_oracle.c); Julia/MPI are not installable here"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ parity gate (no oracle import here)
def hash_uniform(gid0, seed):
    """splitmix64(gid0 + seed*phi) -> [-1,1): the host restatement of pa_vec_fill_hash_box (csrc/pa_vector.cu)."""
    with np.errstate(over="ignore"):
        z = np.asarray(gid0, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -52) - 1.0


def stencil_rows_expected(kind, gn, box_lo, box_hi, ghost_gid0, rows, seed):
    """(A*x)[rows] for x = hash(gid, seed), evaluated from the stencil DEFINITION (src/gallery.jl:36-78 /
    HPCG/src/sparse_matrix.jl:50-75) with the summation order of the stored CSR row: own columns by ascending local id,
    then ghost columns by ascending ghost id (= the order of spmv_csr! on the reference's local matrix)."""
    gn = [int(g) for g in gn]
    b = [int(h - l) for l, h in zip(box_lo, box_hi)]
    rows = np.asarray(rows, dtype=np.int64)
    ix, iy, iz = rows % b[0], (rows // b[0]) % b[1], rows // (b[0] * b[1])
    gx, gy, gz = ix + box_lo[0], iy + box_lo[1], iz + box_lo[2]
    alpha = float(gn[0] + 1) * float(gn[1] + 1) * float(gn[2] + 1)
    offs = [(sx, sy, sz) for sz in (-1, 0, 1) for sy in (-1, 0, 1) for sx in (-1, 0, 1)]
    if kind == 7:
        offs = [o for o in offs if abs(o[0]) + abs(o[1]) + abs(o[2]) <= 1]
    order = np.argsort(ghost_gid0, kind="stable")
    sg = np.asarray(ghost_gid0, dtype=np.int64)[order]
    want = np.zeros(len(rows))
    gid_g = np.full((len(rows), len(offs)), np.iinfo(np.int64).max, dtype=np.int64)  # ghost id of ghost terms
    val_g = np.zeros((len(rows), len(offs)))
    for t, (sx, sy, sz) in enumerate(offs):  # ascending (sz, sy, sx) == ascending own local id
        cx, cy, cz = gx + sx, gy + sy, gz + sz
        inside = (cx >= 0) & (cx < gn[0]) & (cy >= 0) & (cy < gn[1]) & (cz >= 0) & (cz < gn[2])
        own = inside & (cx >= box_lo[0]) & (cx < box_hi[0]) & (cy >= box_lo[1]) & (cy < box_hi[1]) & (cz >= box_lo[2]) & (cz < box_hi[2])
        gid = np.where(inside, cx + gn[0] * (cy + gn[1] * cz), 0)
        diag = (sx, sy, sz) == (0, 0, 0)
        coef = (6.0 * alpha if diag else -alpha) if kind == 7 else (26.0 if diag else -1.0)
        term = coef * hash_uniform(gid, seed)
        want = np.where(own, want + term, want)
        gh = inside & ~own
        if gh.any():
            pos = np.clip(np.searchsorted(sg, gid[gh]), 0, max(len(sg) - 1, 0))
            if len(sg) == 0 or not np.all(sg[pos] == gid[gh]):
                raise RuntimeError("parity gate: a stencil neighbour outside the own box is not a ghost of this part")
            gid_g[gh, t] = order[pos]
            val_g[gh, t] = term[gh]
    if (gid_g != np.iinfo(np.int64).max).any():
        srt = np.argsort(gid_g, axis=1, kind="stable")
        gs, vs = np.take_along_axis(gid_g, srt, 1), np.take_along_axis(val_g, srt, 1)
        for t in range(gs.shape[1]):
            live = gs[:, t] != np.iinfo(np.int64).max
            if not live.any():
                break
            want = np.where(live, want + vs[:, t], want)
    return want


def parity_rows(gn, box_lo, box_hi, n_interior, rng):
    """Own rows on every face of the box that is shared with another part (they touch ghost columns; edges and corners
    included), the 8 box corners, and n_interior random rows."""
    b = [int(h - l) for l, h in zip(box_lo, box_hi)]
    n = b[0] * b[1] * b[2]
    sel = []
    for d in range(3):
        for side, shared in ((0, box_lo[d] > 0), (b[d] - 1, box_hi[d] < gn[d])):
            if not shared:
                continue
            rng_d = [np.arange(b[0]), np.arange(b[1]), np.arange(b[2])]
            rng_d[d] = np.array([side])
            g = np.meshgrid(*rng_d[::-1], indexing="ij")[::-1]
            sel.append((g[0] + b[0] * (g[1] + b[1] * g[2])).reshape(-1))
    n_boundary = int(len(np.unique(np.concatenate(sel)))) if sel else 0
    corners = np.array([x + b[0] * (y + b[1] * z) for z in (0, b[2] - 1) for y in (0, b[1] - 1) for x in (0, b[0] - 1)], dtype=np.int64)
    sel += [corners, rng.integers(0, n, n_interior)]
    return np.unique(np.concatenate(sel)), n_boundary


def parity_gate(pa, A, rhs, kind, gn, schedules, all_max, all_sum, seed=11):
    """See the module docstring.  Returns the parity_check dict; raises SystemExit if any rank sees a difference."""
    ind = A.cols.indices[0]
    box = ind.block.box
    lo, hi = [r[0] - 1 for r in box], [r[1] for r in box]
    x, y = pa.pones(A.cols), pa.pzeros(A.rows)
    # (i) A*ones == rhs exactly (27 - nnz_row for HPCG; alpha * missing neighbours for the gallery operator)
    pa.mul_(y, A, x)
    y.axpby_(-1.0, rhs, 1.0)
    ones_diff = y.norm()
    # (ii) hash-valued x, every schedule, sampled rows bit for bit against the stencil definition
    rows, n_boundary = parity_rows(gn, lo, hi, 2000, np.random.default_rng(5))
    want = stencil_rows_expected(kind, gn, lo, hi, ind.ghost_to_global - 1, rows, seed)
    ghost_want = hash_uniform(ind.ghost_to_global - 1, seed)
    worst, ghost_worst, names = 0.0, 0.0, []
    for name, flags in schedules:
        pa.fill_hash(x, seed)  # ghost slots zeroed: the schedule has to refresh them itself
        y.fill_(-7.0)
        pa.mul_(y, A, x, flags=flags)
        got = y.local_values()[0][rows]
        bad = got != want
        worst = max(worst, float(np.abs(got - want).max()) if bad.any() else 0.0)
        if not (flags & pa.PA_SPMV_SKIP_GHOST_REFRESH) and ind.n_ghost:
            # (iii) consistent!: the ghost slots of x hold the owners' values
            gv = x.local_values()[0][ind.n_own:]
            ghost_worst = max(ghost_worst, float(np.abs(gv - ghost_want).max()))
        names.append(name)
    for v in (x, y):
        v.free()
    fail = all_max(1.0 if (ones_diff != 0.0 or worst != 0.0 or ghost_worst != 0.0) else 0.0)
    out = {"rows": int(all_sum(len(rows))), "boundary_rows_touching_ghosts": int(all_sum(n_boundary)), "ghost_values": int(all_sum(ind.n_ghost)),
           "max_abs_diff": all_max(worst), "a_times_ones_minus_rhs_norm": ones_diff, "ghost_max_abs_diff": all_max(ghost_worst),
           "schedules": names, "passed": fail == 0.0}
    if fail != 0.0:
        raise SystemExit(f"bench.py: PARITY GATE FAILED ({kind}-pt): {json.dumps(out)} — no value is reported")
    return out


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    meta = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        meta = dist.new_group(backend="gloo")
    import pa_b200 as pa

    N = world
    shape = pa.compute_optimal_shape_xyz(N)
    n = args.n
    gn = (n, n, n) if args.strong else (n * shape[0], n * shape[1], n * shape[2])  # --strong: fixed global grid
    stream = torch.cuda.Stream()
    vec_bytes = (n + 2) ** 3 * 8
    backend = pa.CUDAArray(N, mode="distributed" if world > 1 else "sequential", device=local_rank, arena_bytes=14 * vec_bytes + (64 << 20),
                           stream=stream.cuda_stream, group=meta)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def all_max(v):
        if dist is None:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_sum(v):
        if dist is None:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def timed(fn):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        barrier()
        return all_max(e0.elapsed_time(e1)), (w0, time.time())

    schedules = [("default: gather kernel (signal+wait+gather+done in one launch) then one local SpMV", pa.PA_SPMV_DEFAULT),
                 ("fused: consistent! inside the SpMV kernel", pa.PA_SPMV_FUSED_EXCHANGE),
                 ("overlap: gather on a side stream || own block, then ghost block", pa.PA_SPMV_OVERLAP),
                 ("inline: ghost columns dereference the owner's HBM inside the SpMV", pa.PA_SPMV_INLINE_PEER_LOADS)]
    peak, peak_src = measured_peak()
    windows = []
    sampler = ClockSampler(local_rank) if rank == 0 else None
    L = pa._capi.lib()

    A, b = pa.stencil_matrix(args.kind, gn, shape, backend)
    ind = A.cols.indices[0]
    n_rows, n_local, nnz = ind.n_own, ind.n_local, A.nnz(0)
    nnz_total, rows_total = all_sum(nnz), all_sum(n_rows)

    # --- parity gate: nothing below is reported unless it passes on every rank
    parity = {"passed": None, "skipped": "--no-parity"} if args.no_parity else parity_gate(pa, A, b, args.kind, gn, schedules, all_max, all_sum)

    x = pa.pzeros(A.cols)
    y = pa.pzeros(A.rows)
    u = pa.fill_hash(pa.PVector(A.cols), 1)

    # --- SpMV (the dominant kernel): standalone timed region, inputs >> L2 so no flush is needed
    for _ in range(5):
        pa.mul_(y, A, u, flags=pa.PA_SPMV_SKIP_GHOST_REFRESH)
    l0 = backend.launch_count()
    ms_spmv, w = timed(lambda: [pa.mul_(y, A, u, flags=pa.PA_SPMV_SKIP_GHOST_REFRESH) for _ in range(args.spmv_reps)])
    windows.append(w)
    spmv_launches = backend.launch_count() - l0
    ms_spmv /= args.spmv_reps
    B = spmv_bytes(n_rows, nnz, n_local)
    spmv_gbs = B / (ms_spmv * 1e-3) / 1e9
    spmv_gflops = 2 * nnz_total / (ms_spmv * 1e-3) / 1e9

    # --- the same product with the row patterns switched off (values AND column indices streamed: the CSR byte model's kernel)
    backend.set_knob("spmv_patterns", 0)
    for _ in range(3):
        pa.mul_(y, A, u, flags=pa.PA_SPMV_SKIP_GHOST_REFRESH)
    ms_spmv_plain, w = timed(lambda: [pa.mul_(y, A, u, flags=pa.PA_SPMV_SKIP_GHOST_REFRESH) for _ in range(20)])
    windows.append(w)
    ms_spmv_plain /= 20
    backend.set_knob("spmv_patterns", 1)

    # --- every mul! schedule on the same operands (N > 1: they differ only in how the ghost values travel)
    sched_ms = None
    if N > 1:
        sched_ms = {}
        for name, flags in schedules:
            for _ in range(3):
                pa.mul_(y, A, u, flags=flags)
            ms_s, w = timed(lambda: [pa.mul_(y, A, u, flags=flags) for _ in range(20)])
            windows.append(w)
            sched_ms[name.split(":")[0]] = ms_s / 20

    # --- CG steps, operands resident
    flops_iter = 2 * nnz_total + 12 * rows_total
    def step_resident(flags=0):
        x.fill_(0.0)
        return pa.ref_cg_(x, A, b, tolerance=0.0, maxiter=args.iters, flags=flags)
    for _ in range(args.warmup):
        res = step_resident()
    l0 = backend.launch_count()
    ms_cg, w = timed(lambda: [step_resident() for _ in range(args.steps)])
    windows.append(w)
    launches = backend.launch_count() - l0
    ms_step = ms_cg / args.steps
    value = flops_iter * args.iters / (ms_step * 1e-3) / 1e9
    rel_res = res.residual / res.residual0
    # the op-for-op schedule of ref_cg.jl (copy, dot, waxpby, spmv, dot, 2 x waxpby, norm: 3 reductions, 8 passes) beside it:
    # the HPCG flop model (3 dots) is the model of THIS schedule; the fused default executes 2 reductions
    step_resident(pa.PA_CG_REFERENCE_OPS)
    ms_ref_ops, w = timed(lambda: step_resident(pa.PA_CG_REFERENCE_OPS))
    windows.append(w)

    # --- e2e: host buffers in, host buffers out, every step: b uploaded from pinned host memory, x0 = 0 set on the device (the
    # caller's x0 is the zero vector), x and the residual history downloaded.  Serial: upload -> solve -> download, one after
    # the other.  Pipelined (the headline e2e): two (x, b) buffer pairs; the upload of the NEXT right-hand side and the
    # download of the PREVIOUS solution run on their own copy streams while the current solve occupies the SMs (PCIe is full
    # duplex); every step's copies still happen inside the timed region, the last download included.
    hbuf = [(torch.empty(n_local, dtype=torch.float64).pin_memory(), torch.empty(n_local, dtype=torch.float64).pin_memory()) for _ in range(2)]
    for hb_, _ in hbuf:
        hb_.numpy()[:] = b.local_values()[0]
    def step_e2e_serial():
        hb_, hx_ = hbuf[0]
        pa._capi.check(L.pa_vec_upload(b.h, 0, hb_.data_ptr(), n_local))
        x.fill_(0.0)
        r = pa.ref_cg_(x, A, b, tolerance=0.0, maxiter=args.iters)
        pa._capi.check(L.pa_vec_download(x.h, 0, hx_.data_ptr(), n_local))
        return r
    step_e2e_serial()
    ms_e2e_serial, w = timed(lambda: [step_e2e_serial() for _ in range(args.steps)])
    windows.append(w)
    err = float(np.abs(hbuf[0][1].numpy()[:n_rows] - 1.0).max())
    x2, b2 = pa.pzeros(A.cols), pa.PVector(A.cols)
    pairs = [(x, b), (x2, b2)]
    s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
    def e2e_pipelined(K):
        ev_up, ev_down = [None, None], [None, None]
        def upload(sidx):
            q = sidx % 2
            pa._capi.check(L.pa_vec_upload_async(pairs[q][1].h, 0, hbuf[q][0].data_ptr(), n_local, s_h2d.cuda_stream))
            ev_up[q] = torch.cuda.Event(); ev_up[q].record(s_h2d)
        s_h2d.wait_stream(stream)  # the buffers are idle: everything enqueued so far has to finish first
        s_d2h.wait_stream(stream)
        upload(0)
        for sidx in range(K):
            q = sidx % 2
            if sidx + 1 < K:
                upload(sidx + 1)  # b of the other pair: its last reader (solve sidx-1) has returned; overlaps with the solve below
            stream.wait_event(ev_up[q])
            if ev_down[q] is not None:
                stream.wait_event(ev_down[q])  # x of this pair is free again once its previous solution has been read out
            pairs[q][0].fill_(0.0)
            pa.ref_cg_(pairs[q][0], A, pairs[q][1], tolerance=0.0, maxiter=args.iters)  # returns when the solve is complete
            pa._capi.check(L.pa_vec_download_async(pairs[q][0].h, 0, hbuf[q][1].data_ptr(), n_local, s_d2h.cuda_stream))
            ev_down[q] = torch.cuda.Event(); ev_down[q].record(s_d2h)
        stream.wait_stream(s_d2h)  # the last solution has to be on the host inside the timed region
        stream.wait_stream(s_h2d)
    e2e_pipelined(2)
    ms_e2e, w = timed(lambda: e2e_pipelined(args.steps))
    windows.append(w)
    torch.cuda.synchronize()
    err = max(err, float(np.abs(hbuf[(args.steps - 1) % 2][1].numpy()[:n_rows] - 1.0).max()))
    # the floor the host side sets: the same copies (b up, x down, both directions at once, all ranks at once) with no solve in
    # between.  On a box whose GPUs share one host memory system this grows with N and bounds e2e from below.
    def transfers_only(K):
        s_h2d.wait_stream(stream); s_d2h.wait_stream(stream)
        for sidx in range(K):
            q = sidx % 2
            pa._capi.check(L.pa_vec_upload_async(pairs[q][1].h, 0, hbuf[q][0].data_ptr(), n_local, s_h2d.cuda_stream))
            pa._capi.check(L.pa_vec_download_async(pairs[q][0].h, 0, hbuf[q][1].data_ptr(), n_local, s_d2h.cuda_stream))
        stream.wait_stream(s_d2h); stream.wait_stream(s_h2d)
    transfers_only(1)
    ms_xfer, w = timed(lambda: transfers_only(args.steps))
    windows.append(w)
    torch.cuda.synchronize()
    e2e_value = flops_iter * args.iters / (ms_e2e / args.steps * 1e-3) / 1e9
    e2e_serial_value = flops_iter * args.iters / (ms_e2e_serial / args.steps * 1e-3) / 1e9
    h2d, d2h = n_local * 8, n_local * 8 + (args.iters + 1) * 8
    for v in (x2, b2):
        v.free()
    del hbuf

    extra = {"reference_ops": {"cg_iters_per_sec": args.iters / (ms_ref_ops * 1e-3), "gflops": flops_iter * args.iters / ms_ref_ops / 1e6,
                               "note": "PA_CG_REFERENCE_OPS: one kernel per operation of ref_cg.jl:46-67 (3 reductions, 8 passes); the headline is the fused "
                                       "schedule (2 reductions, 3 passes) scored with the same (2*nnz + 12*n) flop model"}}
    for v in (x, y, u, b):
        v.free()
    A.free()

    want_27 = args.kind == 7 and not args.no_hpcg27 and not args.strong
    if want_27:
        # secondary workload: HPCG 27-pt, n^3 per GPU, weak (configs[3]; 64-bit row pointers at 512^3)
        sh = shape
        A27, b27 = pa.build_p_matrix(backend, n, n, n, *sh)
        gn27 = (n * sh[0], n * sh[1], n * sh[2])
        par27 = None if args.no_parity else parity_gate(pa, A27, b27, 27, gn27, schedules[:2], all_max, all_sum)
        ind27 = A27.cols.indices[0]
        x27, y27 = pa.pzeros(A27.cols), pa.pzeros(A27.rows)
        u27 = pa.fill_hash(pa.PVector(A27.cols), 1)
        nnz27 = A27.nnz(0)
        nnz27_t = all_sum(nnz27)
        for _ in range(3):
            pa.mul_(y27, A27, u27)
        ms27, w = timed(lambda: [pa.mul_(y27, A27, u27) for _ in range(20)])
        windows.append(w)
        ms27 /= 20
        backend.set_knob("spmv_patterns", 0)
        for _ in range(2):
            pa.mul_(y27, A27, u27)
        ms27_plain, w = timed(lambda: [pa.mul_(y27, A27, u27) for _ in range(10)])
        windows.append(w)
        ms27_plain /= 10
        backend.set_knob("spmv_patterns", 1)
        def st27():
            x27.fill_(0.0)
            return pa.ref_cg_(x27, A27, b27, tolerance=0.0, maxiter=args.iters)
        st27()
        mscg27, w = timed(st27)
        windows.append(w)
        B27 = spmv_bytes(ind27.n_own, nnz27, ind27.n_local)
        extra["hpcg27"] = {"workload": f"HPCG 27-pt {n}^3 rows per GPU (global {gn27[0]}x{gn27[1]}x{gn27[2]}), parts {sh}, weak", "parity_check": par27,
                           "spmv_ms": ms27, "spmv_ms_column_stream_kernel": ms27_plain, "spmv_gflops": 2 * nnz27_t / ms27 / 1e6, "spmv_hbm_gbs_per_gpu": B27 / ms27 / 1e6,
                           "spmv_frac_of_peak": B27 / ms27 / 1e6 / peak, "cg_iters_per_sec": args.iters / (mscg27 * 1e-3),
                           "cg_gflops": (2 * nnz27_t + 12 * rows_total) * args.iters / mscg27 / 1e6, "nnz_per_gpu": nnz27}
        for v in (x27, y27, u27, b27):
            v.free()
        A27.free()
        if not args.no_mg:
            # HPCG proper: 4-level multigrid (symmetric Gauss-Seidel) preconditioned CG on the same operator
            extra["hpcg_mg"] = mg_section(pa, backend, n, sh, timed, windows, all_sum, peak)

    if N > 1 and not args.strong and not args.no_strong:
        extra["strong"] = strong_section(pa, backend, n, shape, args, timed, windows, all_sum)

    if (N == 4 or args.fem) and not args.no_fem and not args.strong:
        extra["fem_c5"] = fem_section(pa, backend, N, timed, windows, all_sum, all_max, peak)

    clocks = sampler.stop(windows) if sampler else None
    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        nz = max(2 * host_threads(), int(args.cpu_rows // (256 * 256)))
        c = cpu_cg_sample(args.kind, (256, 256, nz), 10)
        cpu = {"value": c["gflops"], "unit": "GFLOP/s", "cores": c["cores"], "kind": "port", "sample": c["sample"],
               "cg_iters_per_sec_on_sample": c["iters_per_sec"], "spmv_gflops": c["spmv_gflops"]}
    if rank == 0:
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "spmv_traffic.json")) as f:
                traffic = json.load(f).get(f"k{args.kind}_n{n}") if N == 1 else None
            if traffic is not None:
                traffic_src = "constant from the committed ncu --set full capture of this kernel on this config (profiles/), not measured by this run"
        except Exception:
            pass
        op = "gallery 7-pt Laplacian" if args.kind == 7 else "HPCG 27-pt operator"
        per_gpu = f"{n}^3 rows per GPU" if not args.strong else f"global {n}^3 rows split over {N} GPUs"
        line = {
            "metric": "hpcg_cg_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{op} {per_gpu} (global {gn[0]}x{gn[1]}x{gn[2]}), parts {shape}, "
                                   f"CSR fp64/int32, ref_cg! {args.iters} iterations per step, Pl=Identity, x0=0, b=A*ones",
                       "l2_policy": "inputs (matrix 11+ GB, vectors 1 GB each) are far larger than the 126 MB L2; no flush needed",
                       "rows_per_gpu": n_rows, "nnz_per_gpu": nnz, "parallelism": f"row-block partition {shape}, one part per GPU",
                       "mul_schedule": schedules[0][0] if N > 1 else "one part: purely local SpMV",
                       "cg_schedule": "fused: direction (+ ghost exchange of u inside the kernel at N > 1) | SpMV+dot | update+norm; scalar all-reduces and epoch signalling folded into those kernels; CUDA-graph replay"},
            "parity_check": parity,
            "cg_iters_per_sec": args.iters / (ms_step * 1e-3), "cg_rel_residual": rel_res, "cg_max_abs_err_after_iters": err,
            "spmv_gflops": spmv_gflops, "spmv_ms": ms_spmv, "mul_schedules_ms": sched_ms,
            "roofline": {"bound": "hbm", "kernel": "k_spmv_pat (TMA-pipelined CSR SpMV, column stream compressed to one pattern byte per row)",
                         "achieved": spmv_gbs, "peak": peak, "unit": "GB/s", "frac": spmv_gbs / peak,
                         "note": "achieved = ALGORITHMIC bytes of the CSR product (SURVEY 8d: 12 B per entry) / time; the kernel moves fewer bytes than that (traffic), which is how frac exceeds 1; traffic / time is what to compare with the HBM peak",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": B, "traffic": traffic, "traffic_source": traffic_src,
                         "column_stream_kernel": {"kernel": "k_spmv_tma (row patterns off: 12 B per entry streamed)", "ms": ms_spmv_plain,
                                                  "achieved": B / (ms_spmv_plain * 1e-3) / 1e9, "frac": B / (ms_spmv_plain * 1e-3) / 1e9 / peak},
                         "cg_iter_bytes_model": B + 120 * n_rows, "cg_frac_of_peak": (B + 120 * n_rows) * args.iters / (ms_step * 1e-3) / 1e9 / peak},
            "e2e": {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "schedule": "pipelined: upload of the next b and download of the previous x on copy streams, overlapped with the running solve",
                    "serial_value": e2e_serial_value, "serial_ms_per_step": ms_e2e_serial / args.steps,
                    "transfers_only_ms_per_step": ms_xfer / args.steps,
                    "host_link_gbs_per_gpu": (h2d + d2h) / (ms_xfer / args.steps) / 1e6, "host_link_gbs_all_gpus": N * (h2d + d2h) / (ms_xfer / args.steps) / 1e6,
                    "note": "transfers_only = the same H2D + D2H copies of every rank at once with no solve in between: the floor the host memory system sets for e2e at this N"},
            "gpu_launches": int(launches), "launches_per_cg_iteration": launches / (args.steps * args.iters), "spmv_region_launches": int(spmv_launches),
            "clocks": clocks, "cpu_baseline": cpu,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    backend.sync()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def mg_section(pa, backend, n, sh, timed, windows, all_sum, peak):
    """HPCG proper.  Two smoother orders: the reference's (lexicographic: bit-identical iterates, the default) and the opt-in
    multi-colour order (convergence-level parity: it needs 59 instead of 50 iterations for the reference tolerance on the
    reference's own test problem, tests/test_gpu_hpcg_mg.py) — both timed, both reported."""
    P = pa.pc_setup(backend, 4, n, n, n, *sh)
    xm = pa.pzeros(P.A.cols)
    xs = pa.pzeros(P.A.cols)
    mg_iters = 10
    nnz_l = [all_sum(P.A_vec[l].nnz(0)) for l in range(4)]
    rows_t = all_sum(P.A.rows.indices[0].n_own)
    # flop model of the reference report (HPCG/src/report_results.jl:27-40): CG ops + per level 4*nnz pre, 2*nnz residual, 4*nnz post
    mg_flops = sum(10 * z for z in nnz_l[1:]) + 4 * nnz_l[0]
    ind = P.A.cols.indices[0]
    sweep_bytes = 2 * (spmv_bytes(ind.n_own, P.A.nnz(0), ind.n_local) + 8 * ind.n_own)  # forward + backward: matrix, x, b in, x out
    out = {"workload": f"HPCG 27-pt {n}^3 per GPU, parts {sh}, 4-level MG (symmetric Gauss-Seidel), ref_cg! Pl=MG, {mg_iters} iterations",
           "residual_restrict": "the residual between the smoothers is computed at the injection points only (same bits as mul_no_lat! + restrict!)",
           "symgs_bytes_model": "2 x (12*nnz + rowptr + 8*n_cols + 8*n + 8*n) per symmetric application (matrix, x, b in, x out, both sweeps)"}
    for order in ("lexicographic", "multicolor"):
        P.set_order(order)
        pa.ref_cg_pc_(xm, P.A, P.b, P, maxiter=2)
        def stmg():
            xm.fill_(0.0)
            return pa.ref_cg_pc_(xm, P.A, P.b, P, tolerance=0.0, maxiter=mg_iters)
        msmg, w = timed(stmg)
        windows.append(w)
        rmg = stmg()
        # one symmetric Gauss-Seidel application on the finest level, timed alone (the kernel furthest from its roofline)
        gs = P.gs[P.l - 1]
        gs.smooth_(xs, P.b, False)
        msgs, w = timed(lambda: [gs.smooth_(xs, P.b, False) for _ in range(3)])
        windows.append(w)
        msgs /= 3
        out[order] = {"pcg_iters_per_sec": mg_iters / (msmg * 1e-3), "ms_per_iter": msmg / mg_iters,
                      "gflops": (2 * nnz_l[3] + 12 * rows_t + mg_flops) * mg_iters / msmg / 1e6,
                      "scaled_residual_after_10": rmg.residual / rmg.residual0,
                      "symgs_finest_ms": msgs, "symgs_finest_hbm_gbs": sweep_bytes / msgs / 1e6, "symgs_finest_frac_of_peak": sweep_bytes / msgs / 1e6 / peak,
                      "smoother": pa.hpcg.smoother_name(P)}
    out["pcg_iters_per_sec"] = out["lexicographic"]["pcg_iters_per_sec"]  # the default (bit-exact) order
    xs.free(); xm.free()
    P.free()
    return out


def strong_section(pa, backend, n, shape, args, timed, windows, all_sum):
    """configs[3] strong scaling: the GLOBAL grid is n^3, split over the parts (7-pt gallery CG and 27-pt HPCG CG)."""
    out = {"workload": f"global {n}^3 rows split over parts {shape}: ref_cg! {args.iters} iterations, Pl=Identity"}
    for kind in (7, 27):
        A, b = pa.stencil_matrix(kind, (n, n, n), shape, backend)
        x = pa.pzeros(A.cols)
        nnz_t, rows_t = all_sum(A.nnz(0)), all_sum(A.rows.indices[0].n_own)
        def st():
            x.fill_(0.0)
            return pa.ref_cg_(x, A, b, tolerance=0.0, maxiter=args.iters)
        st(); st()
        ms, w = timed(lambda: [st() for _ in range(3)])
        windows.append(w)
        ms /= 3
        out[f"k{kind}"] = {"cg_iters_per_sec": args.iters / (ms * 1e-3), "cg_gflops": (2 * nnz_t + 12 * rows_t) * args.iters / ms / 1e6,
                           "ms_per_iter": ms / args.iters, "rows_per_gpu": A.rows.indices[0].n_own}
        x.free(); b.free(); A.free()
    return out


def fem_section(pa, backend, N, timed, windows, all_sum, all_max, peak):
    """configs[4]: fem_example.jl — disassembled Q1 triplets -> psparse (device compression, ghost rows shipped on the
    device) -> pvector -> mul! -> CG, ~10 M dofs (3162^2), rows of 4/6/9 entries, (2,2) parts at N = 4."""
    from pa_b200 import fem_example as fe

    parts = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}.get(N, (N, 1))
    nd = 3162
    t0 = time.perf_counter()
    lay = fe.Q1Layout(parts, (nd + 1, nd + 1), (2.0, 2.0))
    trip = [fe.q1_part(lay, p) for p in backend.parts]
    t_gen = time.perf_counter() - t0
    rows = pa.variable_partition(backend, lay.n_own_dofs, lay.n_global_dofs)
    backend.sync()
    t0 = time.perf_counter()
    A = pa.psparse([t[0] for t in trip], [t[1] for t in trip], [t[2] for t in trip], rows, rows, assembled=False, local_format="csr",
                   compress="device", ship="device")
    backend.sync()
    t_asm = time.perf_counter() - t0
    rhs = pa.pvector_from_triplets([t[3] for t in trip], [t[4] for t in trip], rows)
    nnz_t = all_sum(A.nnz(0))
    ind = A.cols.indices[0]
    # psparse!-style refresh of the values of the final (owner-side) compression stage: one gather-sum kernel per part
    vals = [A.coo_values(k) for k in range(len(backend.parts))]
    A.update_coo_values_(vals)
    refresh_ms, w = timed(lambda: A.update_coo_values_(vals))
    windows.append(w)
    # parity: A * u_exact == rhs at rounding level (Q1 reproduces x1 + x2 exactly)
    xe = pa.PVector(A.cols).set_local_values([np.concatenate([lay.exact_own(p), np.zeros(i.n_ghost)]) for p, i in zip(backend.parts, A.cols.indices)])
    y = pa.pzeros(A.rows)
    pa.mul_(y, A, xe)
    scale = float(np.abs(lay.Ae).max() * 4.0)
    resid = all_max(max(float(np.abs(a - c).max()) for a, c in zip(y.own_values(), rhs.own_values())))
    if not (resid < 64 * np.finfo(float).eps * scale):
        raise SystemExit(f"bench.py: PARITY GATE FAILED (FEM C5): |A*u_exact - rhs|_inf = {resid}")
    for _ in range(5):
        pa.mul_(y, A, xe)
    reps = 100
    ms, w = timed(lambda: [pa.mul_(y, A, xe) for _ in range(reps)])
    windows.append(w)
    ms /= reps
    Bf = spmv_bytes(ind.n_own, A.nnz(0), ind.n_local)
    bc = pa.pzeros(A.cols)
    bc.copy_(rhs)
    xs = pa.pzeros(A.cols)
    def st():
        xs.fill_(0.0)
        return pa.ref_cg_(xs, A, bc, tolerance=0.0, maxiter=200)
    st()
    mscg, w = timed(st)
    windows.append(w)
    res = st()
    out = {"workload": f"Q1 FEM (test/fem_example.jl) {nd}^2 = {lay.n_global_dofs} dofs, parts {parts}, nnz {int(nnz_t)}, rows of 4/6/9 entries",
           "triplet_generation_s_host": t_gen, "psparse_assembly_s": t_asm,
           "coo_value_refresh_ms": refresh_ms, "coo_value_refresh_note": "sparse_matrix!(A,V,K) of the owner-side compression stage (host->device copy of V included)",
           "parity_check": {"a_times_u_exact_minus_rhs_inf": resid, "bound": 64 * np.finfo(float).eps * scale, "passed": True},
           "spmv_ms": ms, "spmv_gflops": 2 * nnz_t / ms / 1e6, "spmv_hbm_gbs_per_gpu": Bf / ms / 1e6, "spmv_frac_of_peak": Bf / ms / 1e6 / peak,
           "cg_iters_per_sec": 200 / (mscg * 1e-3), "cg_rel_residual_after_200": res.residual / res.residual0}
    for v in (xe, y, bc, xs, rhs):
        v.free()
    A.free()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
