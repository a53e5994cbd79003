import os, sys
sys.path.insert(0, os.getcwd())
import torch
import pa_b200 as pa
n = int(sys.argv[1])
b = pa.CUDAArray(1, arena_bytes=14 * (n + 2) ** 3 * 8)
P = pa.pc_setup(b, 1, n, n, n, 1, 1, 1)
A = P.A_vec[0]
x, rhs = pa.pzeros(A.cols), P.b_vec[0]
P.gs[0].smooth_(x, rhs, False)
b.sync()
b.set_knob("gs_trace", int(sys.argv[2]) + 1)
P.gs[0].smooth_(x, rhs, False)
b.sync()
b.close()
