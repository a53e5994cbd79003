"""Short workload for ncu: a few CG iterations (7-pt 512^3), the unfused BLAS-1 ops, a 2-part consistent!/assemble!
on one GPU, and one symmetric Gauss-Seidel smooth (27-pt 256^3)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pa_b200 as pa  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
b = pa.CUDAArray(1, arena_bytes=10 * (n + 2) ** 3 * 8)
A, rhs = pa.stencil_matrix(7, (n, n, n), (1, 1, 1), b)
x = pa.pzeros(A.cols)
pa.ref_cg_(x, A, rhs, maxiter=3)
pa.ref_cg_(x, A, rhs, maxiter=2, flags=pa.PA_CG_REFERENCE_OPS)
y = pa.pzeros(A.cols)
y.axpby_(2.0, x, 1.0)
x.dot(y); x.norm(); x.sum()
y.copy_(x); y.rmul_(0.5); y.fill_(1.0)
for v in (x, y, rhs):
    v.free()
A.free()
b.close()
# ghost exchange kernels: 2 parts on one GPU (z-split 256x256x512 -> two 256x256x256 parts)
b2 = pa.CUDAArray(2, arena_bytes=6 * (258 ** 3) * 8)
A2, r2 = pa.stencil_matrix(7, (256, 256, 512), (1, 1, 2), b2)
v = pa.pones(A2.cols)
v.consistent_().wait()
v.assemble_().wait()
w = pa.pzeros(A2.rows)
for fl in (pa.PA_SPMV_DEFAULT, pa.PA_SPMV_OVERLAP, pa.PA_SPMV_INLINE_PEER_LOADS):
    pa.mul_(w, A2, v, flags=fl)
b2.sync()
b2.close()
# Gauss-Seidel
b3 = pa.CUDAArray(1, arena_bytes=8 * (258 ** 3) * 8)
A3, r3 = pa.stencil_matrix(27, (256, 256, 256), (1, 1, 1), b3)
gs = pa.GaussSeidel(A3, kind=27)
x3 = pa.pzeros(A3.cols)
gs.smooth_(x3, r3, True)
gs.smooth_(x3, r3, False)
b3.sync()
print("done")
