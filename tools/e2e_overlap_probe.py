"""Why does the pipelined e2e of bench.py not reach solve time + (first upload + last download)/K ?  Times one CG solve alone and
with a 1 GiB upload + a 1 GiB download running on copy streams, with and without CUDA-graph replay."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pa_b200 as pa  # noqa: E402

n = 512
stream = torch.cuda.Stream()
b = pa.CUDAArray(1, arena_bytes=14 * (n + 2) ** 3 * 8, stream=stream.cuda_stream)
A, rhs = pa.stencil_matrix(7, (n, n, n), (1, 1, 1), b)
x = pa.pzeros(A.cols)
y = pa.pzeros(A.cols)
nl = A.cols.indices[0].n_local
h_in = torch.ones(nl, dtype=torch.float64).pin_memory()
h_out = torch.empty(nl, dtype=torch.float64).pin_memory()
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
L = pa._capi.lib()


def ev():
    return torch.cuda.Event(enable_timing=True)


def run(label, copies):
    torch.cuda.synchronize()
    k0, k1, c0, c1, c2, c3 = ev(), ev(), ev(), ev(), ev(), ev()
    if copies:
        c0.record(s_up); pa._capi.check(L.pa_vec_upload_async(y.h, 0, h_in.data_ptr(), nl, s_up.cuda_stream)); c1.record(s_up)
        c2.record(s_dn); pa._capi.check(L.pa_vec_download_async(y.h, 0, h_out.data_ptr(), nl, s_dn.cuda_stream)); c3.record(s_dn)
    k0.record(stream)
    x.fill_(0.0)
    pa.ref_cg_(x, A, rhs, tolerance=0.0, maxiter=50)
    k1.record(stream)
    torch.cuda.synchronize()
    msg = f"{label}: solve {k0.elapsed_time(k1):7.1f} ms"
    if copies:
        msg += f" | H2D {c0.elapsed_time(c1):6.1f} ms, D2H {c2.elapsed_time(c3):6.1f} ms (issued before the solve)"
    print(msg, flush=True)


for graph in (1, 0):
    b.set_knob("cg_graph", graph)
    pa.ref_cg_(x, A, rhs, tolerance=0.0, maxiter=50)
    for _ in range(2):
        run(f"cg_graph={graph} alone      ", False)
        run(f"cg_graph={graph} with copies", True)
b.close()
