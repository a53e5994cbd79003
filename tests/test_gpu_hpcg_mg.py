"""GPU parity for the HPCG multigrid preconditioner: Gauss-Seidel sweeps, V-cycle and preconditioned
CG through the C ABI against the oracle (oracle/hpcg_mg.py) and the reference's known answer."""
import numpy as np
import pytest

from oracle import hpcg_mg
from oracle import pa_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pa():
    import pa_b200

    return pa_b200


def _vec(pa, rows, vals):
    return pa.PVector(rows).set_local_values(vals)


# every variant of the sweep kernel: default choice, warp-per-row dataflow kernel with 0 (any row length) / 4 / 8 / 16 / 32 lanes per
# row, and the batch kernel (32 rows of one level per warp) with and without the L2 prefetch
@pytest.mark.parametrize("lanes", [None, "strip", "sell-flow", "sell-rows", "sell-gate", 0, 4, 8, 16, 32, "batch", "batch-noprefetch"])
@pytest.mark.parametrize("npd,nloc,hint", [((2, 2, 1), (8, 6, 4), True), ((1, 2, 2), (4, 4, 6), False), ((1, 1, 1), (10, 9, 8), True),
                                           ((2, 1, 1), (40, 32, 24), True)])
def test_symmetric_gauss_seidel_is_bit_exact(pa, npd, nloc, hint, lanes):
    """smooth! (smoothers.jl:98-125): wavefront sweeps == the reference's sequential per-part sweeps, bit for bit."""
    if nloc[0] >= 40 and lanes not in (None, "strip", "sell-flow", "sell-rows", "sell-gate", 16, "batch", "batch-noprefetch"):
        pytest.skip("the large case runs the default, the 16-lane and the batch kernels")
    lev = hpcg_mg.Level(*nloc, npd)
    P = len(lev.part)
    b = pa.CUDAArray(P, arena_bytes=32 << 20)
    if lanes == "strip":  # a warp owns 32 grid lines of a plane (k_gs_strip); without the box hint the default kernel runs
        b.set_knob("gs_kernel", 3)
    elif lanes in ("sell-flow", "sell-rows", "sell-gate"):  # thread-per-row kernels on the sweep-ordered SELL copy: staged dataflow / per-row pairs / fenced gate
        b.set_knob("gs_kernel", 2)
        b.set_knob("gs_sell_mode", {"sell-flow": 3, "sell-rows": 1, "sell-gate": 2}[lanes])
    elif isinstance(lanes, str):
        b.set_knob("gs_kernel", 1)
        b.set_knob("gs_prefetch", 0 if lanes.endswith("noprefetch") else 1)
    elif lanes is not None:
        b.set_knob("gs_kernel", 0)
        b.set_knob("gs_lanes", lanes)
    gn = tuple(a * c for a, c in zip(npd, nloc))
    A, rhs = pa.stencil_matrix(27, gn, npd, b)
    gs = pa.GaussSeidel(A, kind=27 if hint else None)  # hint=False exercises the generic (host) level schedule
    rng = np.random.default_rng(1)
    bvals = [rng.standard_normal(i.n_local) for i in lev.part]
    for zero_guess in (True, False):
        x0 = [np.zeros(i.n_local) if zero_guess else rng.standard_normal(i.n_local) for i in lev.part]
        xo = [v.copy() for v in x0]
        for _ in range(2):  # two symmetric iterations (the second always with a non-zero guess)
            hpcg_mg.smooth(lev, xo, bvals, zero_guess)
            zg2 = False
        x = _vec(pa, A.cols, x0)
        bv = _vec(pa, A.cols, bvals)
        gs.smooth_(x, bv, zero_guess)
        gs.smooth_(x, bv, zero_guess)
        got = x.local_values()
        # the oracle applied smooth twice with the same flag; mirror exactly
        for k in range(P):
            n = lev.part[k].n_own
            assert np.array_equal(got[k][:n], xo[k][:n]), (zero_guess, k)
        x.free(); bv.free()
    gs.free()
    b.close()


@pytest.mark.parametrize("fused_restrict", [1, 0])
def test_v_cycle_is_bit_exact(pa, fused_restrict):
    """ldiv!(x, P, b) (mg_preconditioner.jl:202-206,314-328): 3 levels, 4 parts, all pieces bit-exact => V-cycle bit-exact.
    fused_restrict = 1 (default): the residual is computed at the injection points only (same bits as mul_no_lat! + restrict!)."""
    npd, n, levels = (2, 2, 1), 16, 3
    mg = hpcg_mg.MG(npd, levels, n, n, n)
    L = mg.levels[levels - 1]
    rng = np.random.default_rng(3)
    r = [rng.standard_normal(i.n_local) for i in L.part]
    o.consistent(r, L.plan)
    want = [np.zeros(i.n_local) for i in L.part]
    mg.ldiv(want, [v.copy() for v in r])
    b = pa.CUDAArray(4, arena_bytes=64 << 20)
    b.set_knob("mg_fused_restrict", fused_restrict)
    P = pa.pc_setup(b, levels, n, n, n, *npd)
    rv = _vec(pa, P.A.cols, r)
    x = pa.pzeros(P.A.cols)
    P.ldiv_(x, rv)
    got = x.local_values()
    for k, ind in enumerate(L.part):
        assert np.array_equal(got[k][: ind.n_own], want[k][: ind.n_own]), k
    P.free()
    b.close()


def test_hpcg_preconditioned_cg_matches_reference_constant(pa):
    """HPCG/test/hpcg_benchmark_tests.jl:31-41: np=4, 32^3 per part, 4 levels, 50 iterations of MG-preconditioned CG:
    ||r||/||r0|| = 2.877476184683206e-13 in the reference.  Residual history vs the oracle rel 1e-8."""
    npd, n, levels = (2, 2, 1), 32, 4
    mg = hpcg_mg.MG(npd, levels, n, n, n)
    L = mg.levels[levels - 1]
    xo, r0o, ro, ito, histo = hpcg_mg.pcg(mg, [v.copy() for v in L.r], [np.zeros(i.n_local) for i in L.part], 50, 0.0)
    b = pa.CUDAArray(4, arena_bytes=256 << 20)
    P = pa.pc_setup(b, levels, n, n, n, *npd)
    x = pa.pzeros(P.A.cols)
    res = pa.ref_cg_pc_(x, P.A, P.b, P, tolerance=0.0, maxiter=50)
    assert res.iters == 50
    scaled = res.residual / res.residual0
    assert scaled < 1e-12  # the reference's own assertion
    assert abs(scaled - 2.877476184683206e-13) <= 1e-6 * 2.877476184683206e-13
    np.testing.assert_allclose(res.history, histo, rtol=1e-8, atol=1e-15 * histo[0])
    got = x.local_values()
    for k, ind in enumerate(L.part):
        np.testing.assert_allclose(got[k][: ind.n_own], xo[k][: ind.n_own], rtol=1e-10, atol=1e-12)
    P.free()
    b.close()


def _smooth_and_compare(pa, lev, A, gs, seed):
    rng = np.random.default_rng(seed)
    bvals = [rng.standard_normal(i.n_local) for i in lev.part]
    for zero_guess in (True, False):
        x0 = [np.zeros(i.n_local) if zero_guess else rng.standard_normal(i.n_local) for i in lev.part]
        xo = [v.copy() for v in x0]
        x, bv = _vec(pa, A.cols, x0), _vec(pa, A.cols, bvals)
        for _ in range(2):
            hpcg_mg.smooth(lev, xo, bvals, zero_guess)
            gs.smooth_(x, bv, zero_guess)
        got = x.local_values()
        for k, ind in enumerate(lev.part):
            assert np.array_equal(got[k][: ind.n_own], xo[k][: ind.n_own]), (zero_guess, k)
        x.free(); bv.free()


@pytest.mark.parametrize("kernel", [0, 1, 2, 3])  # 0 = warp-per-row dataflow kernel (default), 1 = batch kernel, 2 = SELL thread-per-row kernel, 3 = strip kernel
@pytest.mark.parametrize("case", ["fdm7-box", "fdm7-generic", "fem9"])
def test_gauss_seidel_on_short_rows_is_bit_exact(pa, case, kernel):
    """Rows of <= 8 entries (7-pt gallery operator, closed-form and host-computed levels) and <= 16 entries (the Q1 FEM
    operator of test/fem_example.jl, levels from the sparsity pattern): same sweeps as the reference, bit for bit."""
    from oracle import fem_q1

    if case.startswith("fdm7"):
        gn, npd = (14, 12, 10), (2, 1, 2)
        b = pa.CUDAArray(4, arena_bytes=16 << 20)
        b.set_knob("gs_kernel", kernel)
        A, _ = pa.stencil_matrix(7, gn, npd, b)
        I, J, V, rows, cols = o.laplacian_fdm(gn, npd)
        Ao = o.psparse(I, J, V, rows, cols, assembled=True, local_format="csr")
        gs = pa.GaussSeidel(A, kind=7 if case.endswith("box") else None)
    else:
        prob = fem_q1.Q1Problem((2, 2), (13, 9), (2.0, 2.0 * 9 / 13))
        b = pa.CUDAArray(4, arena_bytes=16 << 20)
        b.set_knob("gs_kernel", kernel)
        rows = pa.variable_partition(b, prob.n_own_dofs, prob.n_global_dofs)
        A = pa.psparse(prob.I, prob.J, prob.V, rows, rows, assembled=False, local_format="csr")
        orows = o.variable_partition(prob.n_own_dofs, prob.n_global_dofs)
        Ao = o.psparse(prob.I, prob.J, prob.V, orows, orows, assembled=False, local_format="csr")
        gs = pa.GaussSeidel(A)
    lev = hpcg_mg.Level.from_psparse(Ao)
    _smooth_and_compare(pa, lev, A, gs, 5)
    gs.free()
    b.close()


# colour kernel: "lanes" = every lane loads its own entries (k_gs_sell<W,0>); (slices per tile, stages, batch) = the slices
# stream through a TMA ring (k_gs_color_tma, the default: 2 slices, 2 stages, batches of 14)
# a fourth entry 0 switches the row patterns off (column words streamed instead of one byte per row)
@pytest.mark.parametrize("ckern", ["lanes", (2, 2, 14), (1, 3, 9), (4, 2, 27), (3, 4, 14), (2, 2, 14, 0), (4, 2, 9, 0)])
@pytest.mark.parametrize("npd,nloc", [((2, 2, 1), (8, 6, 4)), ((1, 1, 1), (10, 9, 8)), ((2, 1, 1), (40, 32, 24))])
def test_multicolor_gauss_seidel_matches_the_oracle_order(pa, npd, nloc, ckern):
    """The opt-in multi-colour order (8 colours, one launch per colour): same per-row arithmetic as the reference's
    sweep, rows visited colour by colour — bit-identical to the oracle's sweep in that order, and switching back to the
    lexicographic order gives the reference's iterates again."""
    lev = hpcg_mg.Level(*nloc, npd)
    lev.order, lev.kind = "multicolor", 27
    P = len(lev.part)
    b = pa.CUDAArray(P, arena_bytes=32 << 20)
    _set_color_kernel(b, ckern)
    gn = tuple(a * c for a, c in zip(npd, nloc))
    A, rhs = pa.stencil_matrix(27, gn, npd, b)
    gs = pa.GaussSeidel(A, kind=27).set_order("multicolor")
    _smooth_and_compare(pa, lev, A, gs, 7)
    gs.set_order("lexicographic")
    lev.order = "lexicographic"
    _smooth_and_compare(pa, lev, A, gs, 8)
    gs.free()
    b.close()


def _set_color_kernel(b, ckern):
    if ckern == "lanes":
        b.set_knob("gs_color_kernel", 0)
    else:
        b.set_knob("gs_color_kernel", 1)
        for key, val in zip(("gs_color_slices", "gs_color_stages", "gs_color_batch", "gs_color_patterns"), ckern):
            b.set_knob(key, val)


@pytest.mark.parametrize("ckern", ["lanes", (2, 2, 14), (8, 3, 14), (2, 2, 14, 0)])
@pytest.mark.parametrize("case", ["fdm7-redblack", "fdm27-thin"])
def test_multicolor_gauss_seidel_other_row_widths(pa, case, ckern):
    """Red/black order of the 7-pt gallery operator (rows of <= 7 entries) and the 8-colour order on a box one cell thick
    (rows of <= 9 entries: the generic-width instantiation of the colour kernels), against the oracle's sweep in that order."""
    if case == "fdm7-redblack":
        gn, npd, kind = (14, 12, 10), (2, 1, 2), 7
        I, J, V, rows, cols = o.laplacian_fdm(gn, npd)
        Ao = o.psparse(I, J, V, rows, cols, assembled=True, local_format="csr")
        lev = hpcg_mg.Level.from_psparse(Ao)
        b = pa.CUDAArray(4, arena_bytes=16 << 20)
        A, _ = pa.stencil_matrix(7, gn, npd, b)
    else:
        npd, nloc, kind = (2, 1, 1), (12, 10, 1), 27
        lev = hpcg_mg.Level(*nloc, npd)
        b = pa.CUDAArray(2, arena_bytes=16 << 20)
        A, _ = pa.stencil_matrix(27, tuple(a * c for a, c in zip(npd, nloc)), npd, b)
    _set_color_kernel(b, ckern)
    lev.order, lev.kind = "multicolor", kind
    gs = pa.GaussSeidel(A, kind=kind).set_order("multicolor")
    _smooth_and_compare(pa, lev, A, gs, 9)
    gs.free()
    b.close()


def test_multicolor_preconditioned_cg_convergence_level_parity(pa):
    """HPCG/test/hpcg_benchmark_tests.jl:31-41 with the multi-colour smoother: history equal to the oracle's run in the same
    order (rel 1e-8).  The reference's constant (2.88e-13 after 50 iterations, asserted < 1e-12) belongs to the lexicographic
    order; the 8-colour order reaches 1.9e-11 after 50 iterations and 1e-12 a few iterations later — stated, not hidden."""
    npd, n, levels = (2, 2, 1), 32, 4
    mg = hpcg_mg.MG(npd, levels, n, n, n, order="multicolor")
    L = mg.levels[levels - 1]
    xo, r0o, ro, ito, histo = hpcg_mg.pcg(mg, [v.copy() for v in L.r], [np.zeros(i.n_local) for i in L.part], 60, 0.0)
    b = pa.CUDAArray(4, arena_bytes=256 << 20)
    P = pa.pc_setup(b, levels, n, n, n, *npd, order="multicolor")
    x = pa.pzeros(P.A.cols)
    res = pa.ref_cg_pc_(x, P.A, P.b, P, tolerance=0.0, maxiter=60)
    np.testing.assert_allclose(res.history, histo, rtol=1e-8, atol=1e-15 * histo[0])
    scaled = res.history / res.history[0]
    assert scaled[50] < 1e-10 and scaled[60] < 1e-12  # slower than lexicographic (2.88e-13 at 50), same fixed point
    assert np.abs(np.concatenate(x.own_values()) - 1.0).max() < 1e-9  # exact solution = ones
    P.free()
    b.close()


@pytest.mark.parametrize("order", ["lexicographic", "multicolor"])
def test_hpcg_benchmark_driver_and_report(pa, order):
    """hpcg_benchmark(distribute, np, nx, ny, nz; total_runtime) (HPCG/src/hpcg_benchmark.jl:26-100, called by
    HPCG/test/hpcg_benchmark_tests.jl:43): reference phase, optimised setup phase, timing phase, report with the reference's
    fields (report_results.jl:89-152).  np=4, 32^3 per part: the reference tolerance is the 2.88e-13 of the known answer; the
    multi-colour smoother needs more than 50 iterations to reach it and the driver finds that count."""
    b = pa.CUDAArray(4, arena_bytes=256 << 20)
    rep = pa.hpcg_benchmark(b, 32, 32, 32, 2, 2, 1, total_runtime=0.2, order=order, max_sets=3)
    assert rep["procs"] == 4 and rep["nr_equations"] == 4 * 32 ** 3
    assert rep["geometry"]["gnx"] == 64 and rep["geometry"]["gnz"] == 32
    assert set(rep["main_times"]) == {"setup", "total", "DDOT", "WAXPBY", "SPMV", "allreduce", "MG", "halo_time", "opt_time", "ref_time"}
    t = rep["main_times"]
    assert t["MG"] > 0 and t["SPMV"] > 0 and t["DDOT"] > 0 and t["WAXPBY"] > 0
    assert abs(t["MG"] + t["SPMV"] + t["DDOT"] + t["WAXPBY"] - t["total"]) <= 1e-6 * t["total"] + 1e-9
    assert abs(rep["reference_tolerance"] - 2.877476184683206e-13) <= 1e-6 * 2.877476184683206e-13  # reference phase = reference sweeps
    if order == "lexicographic":
        assert rep["iter_data"]["opt_iters_set"] == 50
    else:
        assert 55 <= rep["iter_data"]["opt_iters_set"] <= 70  # the multi-colour order pays for its parallelism in iterations
    assert rep["reproducibility_data"]["var"] == 0.0          # every set reproduces the same residual bit for bit
    assert rep["GFLOP/s"]["Total"] > 0 and rep["flops"]["MG"] > rep["flops"]["SpMV"]
    b.close()
