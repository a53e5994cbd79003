"""Short workload for ncu (round 2): the kernels added or changed this round, few launches each.
  1. 2 parts x-split on ONE GPU (400^3 rows each): mul! default (k_consistent_sync is not used with 2 local parts: k_consistent
     + k_spmv_tma MODE 0) and PA_SPMV_FUSED_EXCHANGE (k_spmv_tma MODE 4, per-tile ghost gating)
  2. 7-pt 512^3, 1 part: 3 iterations of the folded CG (k_cg_direction, k_spmv_tma with the folded dot epilogue, k_cg_update_fold)
  3. 27-pt 256^3: one symmetric Gauss-Seidel application in both orders (k_gs_flow_pipe; k_gs_sell<27,0> per colour)
  4. "mg": 27-pt GS_N^3 (default 512), 2 levels, multi-colour order: one V-cycle (k_gs_color_tma per colour, k_residual_restrict,
     k_prolong) — the kernels of the second half of round 2"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pa_b200 as pa  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "spmv"):
    n = 400
    b = pa.CUDAArray(2, arena_bytes=6 * (n + 2) ** 3 * 8)
    A, _ = pa.stencil_matrix(7, (2 * n, n, n), (2, 1, 1), b)
    x = pa.fill_hash(pa.PVector(A.cols), 1)
    y = pa.pzeros(A.rows)
    for flags in (pa.PA_SPMV_DEFAULT, pa.PA_SPMV_FUSED_EXCHANGE, pa.PA_SPMV_DEFAULT, pa.PA_SPMV_FUSED_EXCHANGE):
        pa.mul_(y, A, x, flags=flags)
    b.sync()
    for v in (x, y):
        v.free()
    A.free()
    b.close()
if which in ("all", "cg"):
    n = 512
    b = pa.CUDAArray(1, arena_bytes=8 * (n + 2) ** 3 * 8)
    A, rhs = pa.stencil_matrix(7, (n, n, n), (1, 1, 1), b)
    x = pa.pzeros(A.cols)
    pa.ref_cg_(x, A, rhs, maxiter=3)
    for v in (x, rhs):
        v.free()
    A.free()
    b.close()
if which in ("all", "gs"):
    n = int(os.environ.get("GS_N", "256"))
    b = pa.CUDAArray(1, arena_bytes=8 * (n + 2) ** 3 * 8)
    A, rhs = pa.stencil_matrix(27, (n, n, n), (1, 1, 1), b)
    gs = pa.GaussSeidel(A, kind=27)
    x = pa.pzeros(A.cols)
    gs.smooth_(x, rhs, False)
    gs.set_order("multicolor")
    gs.smooth_(x, rhs, False)
    b.sync()
    b.close()
if which == "mg":
    n = int(os.environ.get("GS_N", "512"))
    b = pa.CUDAArray(1, arena_bytes=14 * (n + 2) ** 3 * 8)
    P = pa.pc_setup(b, 2, n, n, n, 1, 1, 1, order="multicolor")
    c = pa.pzeros(P.A.cols)
    P.ldiv_(c, P.b)
    b.sync()
    b.close()
