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


@pytest.mark.parametrize("lanes", [None, 0, 8, 16, 32])  # every variant of the sweep kernel (lanes per row; 0 = any row length)
@pytest.mark.parametrize("npd,nloc,hint", [((2, 2, 1), (8, 6, 4), True), ((1, 2, 2), (4, 4, 6), False), ((1, 1, 1), (10, 9, 8), True)])
def test_symmetric_gauss_seidel_is_bit_exact(pa, npd, nloc, hint, lanes):
    """smooth! (smoothers.jl:98-125): wavefront sweeps == the reference's sequential per-part sweeps, bit for bit."""
    lev = hpcg_mg.Level(*nloc, npd)
    P = len(lev.part)
    b = pa.CUDAArray(P, arena_bytes=32 << 20)
    if lanes is not None:
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


def test_v_cycle_is_bit_exact(pa):
    """ldiv!(x, P, b) (mg_preconditioner.jl:202-206,314-328): 3 levels, 4 parts, all pieces bit-exact => V-cycle bit-exact."""
    npd, n, levels = (2, 2, 1), 16, 3
    mg = hpcg_mg.MG(npd, levels, n, n, n)
    L = mg.levels[levels - 1]
    rng = np.random.default_rng(3)
    r = [rng.standard_normal(i.n_local) for i in L.part]
    o.consistent(r, L.plan)
    want = [np.zeros(i.n_local) for i in L.part]
    mg.ldiv(want, [v.copy() for v in r])
    b = pa.CUDAArray(4, arena_bytes=64 << 20)
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
