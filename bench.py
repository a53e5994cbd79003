#!/usr/bin/env python
"""bench.py — HPCG-style CG + SpMV benchmark of the PSparseMatrix x PVector hot path on B200.

Workload (BASELINE.json configs[1] at N=1, configs[2] at N=8): gallery 7-pt Laplacian, 512^3 rows per GPU,
fp64 values / int32 columns, weak scaling over a (npx,npy,npz) part grid, one part per GPU / process.
A "step" is one ref_cg!(x,A,b; maxiter=ITERS, Pl=Identity) call (HPCG/src/ref_cg.jl:119-134) from x0=0.

  value  = HPCG-model GFLOP/s of the CG loop, whole job, operands resident in HBM
           ((2*nnz + 12*n) flop per iteration — HPCG/src/report_results.jl:27-29 — x iterations / time)
  e2e    = same metric through the public API with HOST buffers: every step uploads b and x0 from pinned host
           memory and downloads x and the residual history inside the timed region
  roofline = the SpMV kernel (dominant): algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline = the CPU oracle (C twin of the reference loops, one part per host thread) on a bounded sample

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
    ap.add_argument("--cpu-rows", type=int, default=1 << 24, help="rows of the CPU sample")
    ap.add_argument("--no-hpcg27", action="store_true")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the GLOBAL grid is n^3, split over the parts")
    ap.add_argument("--mg", action="store_true", help="(default on) HPCG multigrid-preconditioned CG section (27-pt 512^3, 4 levels)")
    ap.add_argument("--no-mg", action="store_true", help="skip the multigrid-preconditioned CG section")
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


def cpu_cg_sample(kind, rows_target, iters, threads=None):
    """The oracle CG (one part per host thread) on a z-slab sample of the same operator.  Returns a dict."""
    from oracle import c_oracle, pa_oracle as o

    key = (kind, rows_target, threads)
    if key not in _CPU_CACHE:
        _CPU_CACHE[key] = _cpu_build(kind, rows_target, threads)
    mats, plan, bvals, gn, P = _CPU_CACHE[key]
    return _cpu_run(kind, mats, plan, bvals, gn, P, iters)


def _cpu_build(kind, rows_target, threads):
    from oracle import c_oracle, pa_oracle as o

    assert c_oracle.available()
    P = threads or min(c_oracle.max_threads(), os.cpu_count() or 1)
    nx = ny = 256
    nz_part = max(2, int(rows_target // (nx * ny * P)))
    gn = (nx, ny, nz_part * P)
    part = o.uniform_partition((1, 1, P), gn)
    # ghosts of a z-slab: the adjacent planes, in first-appearance order (7-pt: -z plane then +z interleaved per row;
    # computed by the oracle's own union_ghost on the boundary rows)
    mats, cols, bvals = [], [], []
    for ind in part:
        lo = [r[0] - 1 for r in ind.box]
        hi = [r[1] for r in ind.box]
        gh = []
        # boundary planes only: emit neighbour columns of the two z-faces in the generator's order
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
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    vals, last = [], None
    for s in range(args.warmup + args.steps):
        last = cpu_cg_sample(args.kind, args.cpu_rows, min(args.iters, 10))
        if s >= args.warmup:
            vals.append(last)
        if time.perf_counter() - t_all > 240 and len(vals) >= 1:
            break
    secs = sum(v["seconds"] for v in vals)
    flops = sum(v["gflops"] * v["seconds"] for v in vals)
    g = flops / secs
    line = {"impl": "reference", "metric": "hpcg_cg_gflops", "value": g, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": len(vals),
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / len(vals), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"gallery {args.kind}-pt Laplacian CG (Pl=Identity), CPU oracle on a bounded sample: {last['sample']}"},
            "cg_iters_per_sec": last["iters_per_sec"], "spmv_gflops": last["spmv_gflops"],
            "cpu_baseline": {"value": g, "unit": "GFLOP/s", "cores": last["cores"], "kind": "port", "sample": last["sample"]},
            "e2e": {"value": g, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = CPU restatement of the Julia loops (oracle/pa_oracle.c); Julia/MPI are not installable here"}
    print(json.dumps(line), flush=True)


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
    backend = pa.CUDAArray(N, mode="distributed" if world > 1 else "sequential", device=local_rank, arena_bytes=8 * vec_bytes + (64 << 20),
                           stream=stream.cuda_stream, group=meta)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
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
        return max_over_ranks(e0.elapsed_time(e1)), (w0, time.time())

    A, b = pa.stencil_matrix(args.kind, gn, shape, backend)
    ind = A.cols.indices[0]
    n_rows, n_local, nnz = ind.n_own, ind.n_local, A.nnz(0)
    x = pa.pzeros(A.cols)
    y = pa.pzeros(A.rows)
    u = pa.fill_hash(pa.PVector(A.cols), 1)
    windows = []
    sampler = ClockSampler(local_rank) if rank == 0 else None

    # --- SpMV (the dominant kernel): standalone timed region, inputs >> L2 so no flush is needed
    for _ in range(5):
        pa.mul_(y, A, u, flags=pa.PA_SPMV_SKIP_GHOST_REFRESH)
    l0 = backend.launch_count()
    ms_spmv, w = timed(lambda: [pa.mul_(y, A, u, flags=pa.PA_SPMV_SKIP_GHOST_REFRESH) for _ in range(args.spmv_reps)])
    windows.append(w)
    spmv_launches = backend.launch_count() - l0
    ms_spmv /= args.spmv_reps
    B = spmv_bytes(n_rows, nnz, n_local)
    peak, peak_src = measured_peak()
    spmv_gbs = B / (ms_spmv * 1e-3) / 1e9
    spmv_gflops = 2 * nnz / (ms_spmv * 1e-3) / 1e9 * N

    # --- CG steps, operands resident
    flops_iter = (2 * nnz + 12 * n_rows) * N
    def step_resident():
        x.fill_(0.0)
        return pa.ref_cg_(x, A, b, tolerance=0.0, maxiter=args.iters)
    for _ in range(args.warmup):
        res = step_resident()
    l0 = backend.launch_count()
    ms_cg, w = timed(lambda: [step_resident() for _ in range(args.steps)])
    windows.append(w)
    launches = backend.launch_count() - l0
    ms_step = ms_cg / args.steps
    value = flops_iter * args.iters / (ms_step * 1e-3) / 1e9
    rel_res = res.residual / res.residual0

    # --- e2e: host buffers in, host buffers out, every step
    hb = torch.empty(n_local, dtype=torch.float64).pin_memory()
    hx = torch.zeros(n_local, dtype=torch.float64).pin_memory()
    hb.numpy()[:] = b.local_values()[0]
    L = pa._capi.lib()
    def step_e2e():
        pa._capi.check(L.pa_vec_upload(b.h, 0, hb.data_ptr(), n_local))
        hx.zero_()
        pa._capi.check(L.pa_vec_upload(x.h, 0, hx.data_ptr(), n_local))
        r = pa.ref_cg_(x, A, b, tolerance=0.0, maxiter=args.iters)
        pa._capi.check(L.pa_vec_download(x.h, 0, hx.data_ptr(), n_local))
        return r
    step_e2e()
    ms_e2e, w = timed(lambda: [step_e2e() for _ in range(args.steps)])
    windows.append(w)
    e2e_value = flops_iter * args.iters / (ms_e2e / args.steps * 1e-3) / 1e9
    h2d, d2h = 2 * n_local * 8, n_local * 8 + (args.iters + 1) * 8
    err = float(np.abs(hx.numpy()[:n_rows] - 1.0).max())

    extra = {}
    if args.kind == 7 and not args.no_hpcg27 and N == 1 and n == 512:
        # secondary workload: HPCG 27-pt 512^3 (configs[3] at 1 GPU; 64-bit row pointers)
        for v in (x, y, u, b):
            v.free()
        A.free()
        A27, b27 = pa.build_p_matrix(backend, n, n, n, 1, 1, 1)
        x27, y27 = pa.pzeros(A27.cols), pa.pzeros(A27.rows)
        u27 = pa.fill_hash(pa.PVector(A27.cols), 1)
        nnz27 = A27.nnz(0)
        for _ in range(3):
            pa.mul_(y27, A27, u27)
        ms27, w = timed(lambda: [pa.mul_(y27, A27, u27) for _ in range(20)])
        windows.append(w)
        ms27 /= 20
        pa.ref_cg_(x27, A27, b27, maxiter=5)
        def st27():
            x27.fill_(0.0)
            return pa.ref_cg_(x27, A27, b27, tolerance=0.0, maxiter=args.iters)
        mscg27, w = timed(st27)
        windows.append(w)
        B27 = spmv_bytes(n_rows, nnz27, n_rows)
        extra["hpcg27_512"] = {"spmv_ms": ms27, "spmv_gflops": 2 * nnz27 / ms27 / 1e6, "spmv_hbm_gbs": B27 / ms27 / 1e6,
                               "spmv_frac_of_peak": B27 / ms27 / 1e6 / peak, "cg_iters_per_sec": args.iters / (mscg27 * 1e-3),
                               "cg_gflops": (2 * nnz27 + 12 * n_rows) * args.iters / mscg27 / 1e6, "nnz": nnz27}
        if not args.no_mg:
            # HPCG proper: 4-level multigrid (symmetric Gauss-Seidel) preconditioned CG on the same operator
            for v in (x27, y27, u27, b27):
                v.free()
            A27.free()
            P = pa.pc_setup(backend, 4, n, n, n, 1, 1, 1)
            xm = pa.pzeros(P.A.cols)
            pa.ref_cg_pc_(xm, P.A, P.b, P, maxiter=2)
            mg_iters = 10
            def stmg():
                xm.fill_(0.0)
                return pa.ref_cg_pc_(xm, P.A, P.b, P, tolerance=0.0, maxiter=mg_iters)
            msmg, w = timed(stmg)
            windows.append(w)
            rmg = stmg()
            # flop model of the reference report (HPCG/src/report_results.jl:27-40): CG ops + per level 4*nnz pre, 2*nnz residual, 4*nnz post
            nnz_l = [P.A_vec[l].nnz(0) for l in range(4)]
            mg_flops = sum(10 * z for z in nnz_l[1:]) + 4 * nnz_l[0]
            extra["hpcg_mg_512"] = {"pcg_iters_per_sec": mg_iters / (msmg * 1e-3), "ms_per_iter": msmg / mg_iters,
                                    "gflops": (2 * nnz27 + 12 * n_rows + mg_flops) * mg_iters / msmg / 1e6,
                                    "scaled_residual_after_10": rmg.residual / rmg.residual0,
                                    "note": "bit-exact wavefront Gauss-Seidel (same iterates as the reference's sequential sweeps)"}

    clocks = sampler.stop(windows) if sampler else None
    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        c = cpu_cg_sample(args.kind, args.cpu_rows, 10)
        cpu = {"value": c["gflops"], "unit": "GFLOP/s", "cores": c["cores"], "kind": "port", "sample": c["sample"],
               "cg_iters_per_sec_on_sample": c["iters_per_sec"], "spmv_gflops": c["spmv_gflops"]}
    if rank == 0:
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "spmv_traffic.json")) as f:
                traffic = json.load(f).get(f"k{args.kind}_n{n}")
        except Exception:
            pass
        line = {
            "metric": "hpcg_cg_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": N, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{("gallery 7-pt Laplacian" if args.kind == 7 else "HPCG 27-pt operator")} {n}^3 rows per GPU (global {gn[0]}x{gn[1]}x{gn[2]}), parts {shape}, "
                                   f"CSR fp64/int32, ref_cg! {args.iters} iterations per step, Pl=Identity, x0=0, b=A*ones",
                       "l2_policy": "inputs (matrix 11+ GB, vectors 1 GB each) are far larger than the 126 MB L2; no flush needed",
                       "rows_per_gpu": n_rows, "nnz_per_gpu": nnz, "parallelism": f"row-block partition {shape}, one part per GPU"},
            "cg_iters_per_sec": args.iters / (ms_step * 1e-3), "cg_rel_residual": rel_res, "cg_max_abs_err_after_iters": err,
            "spmv_gflops": spmv_gflops, "spmv_ms": ms_spmv,
            "roofline": {"bound": "hbm", "kernel": "k_spmv_tma", "achieved": spmv_gbs, "peak": peak, "unit": "GB/s", "frac": spmv_gbs / peak,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": B, "traffic": traffic,
                         "cg_iter_bytes_model": B + 120 * n_rows, "cg_frac_of_peak": (B + 120 * n_rows) * args.iters / (ms_step * 1e-3) / 1e9 / peak},
            "e2e": {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "spmv_region_launches": int(spmv_launches), "clocks": clocks, "cpu_baseline": cpu,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    backend.sync()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
