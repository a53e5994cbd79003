"""A small 2-part job (one process, one GPU) that touches every kernel family of the hot path — all four mul! schedules, the
folded/fused CG loop, consistent!/assemble!, reductions, the Gauss-Seidel dataflow kernel and the multi-colour kernel — for
compute-sanitizer (racecheck: shared-memory hazards of the TMA ring / staging buffers; memcheck: out-of-bounds accesses;
synccheck: barrier misuse).  Run:  compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pa_b200 as pa  # noqa: E402


def main():
    for nparts, npd in ((2, (2, 1, 1)), (1, (1, 1, 1))):
        b = pa.CUDAArray(nparts, arena_bytes=16 << 20)
        # the row-pattern kernels (k_spmv_pat, k_gs_color_tma<PAT>) are built for parts of >= 4096 rows by default: force them here
        b.set_knob("spmv_pattern_min_rows", 1)
        b.set_knob("gs_pattern_min_rows", 1)
        for kind in (7, 27):
            A, rhs = pa.stencil_matrix(kind, (12 * npd[0], 10, 8), npd, b)
            x = pa.fill_hash(pa.PVector(A.cols), 3)
            y = pa.pzeros(A.rows)
            for flags in (pa.PA_SPMV_DEFAULT, pa.PA_SPMV_FUSED_EXCHANGE, pa.PA_SPMV_OVERLAP, pa.PA_SPMV_INLINE_PEER_LOADS):
                pa.mul_(y, A, x, flags=flags)
            x.consistent_().wait()
            x.assemble_().wait()
            assert np.isfinite(x.dot(y)) and np.isfinite(y.norm())
            xs = pa.pzeros(A.cols)
            res = pa.ref_cg_(xs, A, rhs, tolerance=0.0, maxiter=12)
            res = pa.ref_cg_(xs, A, rhs, tolerance=0.0, maxiter=6, flags=pa.PA_CG_REFERENCE_OPS)
            gs = pa.GaussSeidel(A, kind=kind)
            gs.smooth_(xs, rhs, False)
            gs.set_order("multicolor")
            gs.smooth_(xs, rhs, False)
            gs.smooth_(xs, rhs, True)
            gs.free()
            if kind == 27 and nparts == 1:  # V-cycle: k_residual_restrict, k_prolong
                P = pa.pc_setup(b, 2, 8, 8, 8, 1, 1, 1, order="multicolor")
                cvec = pa.pzeros(P.A.cols)
                P.ldiv_(cvec, P.b)
                cvec.free()
                P.free()
            for v in (x, y, xs, rhs):
                v.free()
            A.free()
        b.close()
    print("SANITIZE_SMALL_OK")


if __name__ == "__main__":
    main()
