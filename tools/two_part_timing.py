"""Two parts on ONE GPU (x-split; every line of the grid has a row with a ghost column): time mul! and one symmetric multi-colour
Gauss-Seidel application — the row patterns have to cope with rows that read their columns from colval."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pa_b200 as pa  # noqa: E402


def timed(stream, b, fn, reps):
    fn(); b.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream); e1.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 320
    stream = torch.cuda.Stream()
    for kind in (7, 27):
        b = pa.CUDAArray(2, arena_bytes=10 * (n + 2) ** 3 * 8, stream=stream.cuda_stream)
        A, rhs = pa.stencil_matrix(kind, (2 * n, n, n), (2, 1, 1), b)
        x = pa.fill_hash(pa.PVector(A.cols), 1)
        y = pa.pzeros(A.rows)
        ms = timed(stream, b, lambda: pa.mul_(y, A, x), 10)
        nnz = A.nnz(0) + A.nnz(1)
        print(f"{kind}-pt 2 x {n}^3: mul! {ms:.3f} ms = {12 * nnz / ms / 1e6:.0f} GB/s of CSR bytes (both parts, one GPU)", flush=True)
        if kind == 27:
            gs = pa.GaussSeidel(A, kind=27).set_order("multicolor")
            ms = timed(stream, b, lambda: gs.smooth_(x, rhs, False), 3)
            print(f"27-pt 2 x {n}^3: symmetric multi-colour GS {ms:.3f} ms", flush=True)
        b.close()


if __name__ == "__main__":
    main()
