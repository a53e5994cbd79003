"""Sweep SpMV kernel configurations on one GPU (tuning aid; prints ms / GB/s / fraction of measured peak)."""
import itertools
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pa_b200 as pa  # noqa: E402
from bench import measured_peak, spmv_bytes  # noqa: E402


def main():
    kind = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    reps = 20
    stream = torch.cuda.Stream()
    b = pa.CUDAArray(1, arena_bytes=4 * (n + 2) ** 3 * 8, stream=stream.cuda_stream)
    A, _ = pa.stencil_matrix(kind, (n, n, n), (1, 1, 1), b, with_rhs=False)
    x = pa.fill_hash(pa.PVector(A.cols), 1)
    y = pa.pzeros(A.rows)
    nnz, nr = A.nnz(0), A.rows.indices[0].n_own
    B = spmv_bytes(nr, nnz, nr)
    peak, _ = measured_peak()
    ref = None

    def run(label, **knobs):
        nonlocal ref
        for k, v in knobs.items():
            b.set_knob(k, v)
        try:
            for _ in range(3):
                pa.mul_(y, A, x)
            b.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(reps):
                pa.mul_(y, A, x)
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1) / reps
            s = y.norm()
            if ref is None:
                ref = s
            print(f"{label:48s} {ms:8.4f} ms  {B / ms / 1e6:8.1f} GB/s  frac {B / ms / 1e6 / peak:.3f}  {'OK' if s == ref else 'MISMATCH'}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{label:48s} failed: {e}", flush=True)

    # streaming references measured the same way (CUDA events): what pure reads / read+write reach on this GPU
    def stream_ref(label, fn, nbytes):
        for _ in range(3):
            fn()
        b.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{label:48s} {ms:8.4f} ms  {nbytes / ms / 1e6:8.1f} GB/s  frac {nbytes / ms / 1e6 / peak:.3f}", flush=True)

    stream_ref("ref: dot(x,y)   (16n read)", lambda: x.dot(y), 16 * nr)
    stream_ref("ref: norm(x)    (8n read)", lambda: x.norm(), 8 * nr)
    stream_ref("ref: y=2x+y     (16n read + 8n write)", lambda: y.axpby_(2.0, x, 1.0), 24 * nr)
    stream_ref("ref: copy y<-x  (8n read + 8n write)", lambda: y.copy_(x), 16 * nr)
    if os.environ.get("SWEEP_QUICK"):
        run("tma default", spmv_kernel=3)
        b.close()
        return
    if os.environ.get("SWEEP_COMBOS"):  # "rows:stages:ctas:batch,..." with the current PA_SPMV_PATTERNS setting
        for item in os.environ["SWEEP_COMBOS"].split(","):
            rows, stages, ctas, batch = (int(q) for q in item.split(":"))
            run(f"tma rows={rows} stages={stages} ctas={ctas} batch={batch}", spmv_kernel=3, tma_rows=rows, tma_stages=stages, tma_ctas=ctas, tma_batch=batch)
        b.close()
        return
    run("v1 stream kernel", spmv_kernel=1)
    if kind == 7:
        combos = [(256, 2, 0, 8), (256, 2, 0, 16), (128, 2, 0, 8), (256, 3, 0, 8), (192, 2, 0, 8), (224, 2, 0, 8), (256, 2, 4, 8), (256, 2, 3, 8)]
    else:
        combos = [(r, s_, 0, bt) for r in (32, 64, 96, 128) for s_ in (2, 3) for bt in (16, 32)]
    for rows, stages, ctas, batch in combos:
        run(f"tma rows={rows} stages={stages} ctas={ctas} batch={batch}", spmv_kernel=3, tma_rows=rows, tma_stages=stages, tma_ctas=ctas, tma_batch=batch)
    b.close()


if __name__ == "__main__":
    main()
