"""Time the HPCG multigrid pieces on one GPU: symmetric GS per level, V-cycle, preconditioned CG."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pa_b200 as pa  # noqa: E402


def timed(stream, backend, fn, reps):
    fn()
    backend.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    levels = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    stream = torch.cuda.Stream()
    b = pa.CUDAArray(1, arena_bytes=14 * (n + 2) ** 3 * 8, stream=stream.cuda_stream)
    t0 = time.time()
    order = os.environ.get("MG_ORDER", "lexicographic")
    P = pa.pc_setup(b, levels, n, n, n, 1, 1, 1, order=order)
    b.sync()
    print(f"order {order}, gs_kernel {os.environ.get('PA_GS_KERNEL', 'default')}", flush=True)
    print(f"setup {time.time() - t0:.2f} s", flush=True)
    for lev in reversed(range(levels)):
        A = P.A_vec[lev]
        x, rhs = pa.pzeros(A.cols), P.b_vec[lev]
        nnz, nr = A.nnz(0), A.rows.indices[0].n_own
        ms = timed(stream, b, lambda: P.gs[lev].smooth_(x, rhs, False), 3)
        B = 2 * (nnz * 12 + nr * (8 if nnz >= 2 ** 31 else 4) + 24 * nr)
        print(f"level {lev} ({nr} rows): symmetric GS {ms:9.3f} ms  ~{B / ms / 1e6:7.1f} GB/s  ({4 * nnz / ms / 1e6:7.1f} GFLOP/s)", flush=True)
        x.free()
    if os.environ.get("MG_QUICK"):
        b.close()
        return
    A = P.A
    x, c = pa.pzeros(A.cols), pa.pzeros(A.cols)
    ms = timed(stream, b, lambda: P.ldiv_(c, P.b), 3)
    print(f"V-cycle (ldiv!): {ms:9.3f} ms", flush=True)
    x.fill_(0.0)
    res = pa.ref_cg_pc_(x, A, P.b, P, maxiter=5)
    its = 25
    def run():
        x.fill_(0.0)
        return pa.ref_cg_pc_(x, A, P.b, P, tolerance=0.0, maxiter=its)
    ms = timed(stream, b, run, 1)
    res = run()
    print(f"MG-preconditioned CG: {its / ms * 1e3:7.2f} iters/s  ({ms / its:8.3f} ms/iter), scaled residual after {its}: {res.residual / res.residual0:.3e}", flush=True)
    b.close()


if __name__ == "__main__":
    main()
