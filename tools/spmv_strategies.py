"""Time the mul! schedules on ONE GPU with several parts in one process (the DebugArray execution model): a proxy for the
kernel-side cost of each schedule (the peers are the same GPU, so no NVLink latency is involved)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pa_b200 as pa  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    npd = (2, 1, 1)
    stream = torch.cuda.Stream()
    b = pa.CUDAArray(2, arena_bytes=6 * (n + 2) ** 3 * 8, stream=stream.cuda_stream)
    A, _ = pa.stencil_matrix(7, (2 * n, n, n), npd, b)
    x = pa.fill_hash(pa.PVector(A.cols), 1)
    y = pa.pzeros(A.rows)
    for name, flags in (("explicit", pa.PA_SPMV_DEFAULT), ("fused", pa.PA_SPMV_FUSED_EXCHANGE), ("overlap", pa.PA_SPMV_OVERLAP),
                        ("inline", pa.PA_SPMV_INLINE_PEER_LOADS), ("explicit", pa.PA_SPMV_DEFAULT), ("fused", pa.PA_SPMV_FUSED_EXCHANGE)):
        for _ in range(5):
            pa.mul_(y, A, x, flags=flags)
        b.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(30):
            pa.mul_(y, A, x, flags=flags)
        e1.record(stream)
        e1.synchronize()
        print(f"{name:9s} {e0.elapsed_time(e1) / 30:8.4f} ms per mul! (2 parts of {n}^3 rows on one GPU)", flush=True)
    b.close()


if __name__ == "__main__":
    main()
