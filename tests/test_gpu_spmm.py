"""spmm / spmtm / rap on the device (src/p_sparse_matrix.jl:2212-2307) against dense products of the centralised matrices —
the reference's own check: `B = A*A; centralize(B) == centralize(A)*centralize(A)` (test/p_sparse_matrix_tests.jl:131-164)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def centralize(M):
    """Dense global matrix of an assembled PSparseMatrix (test helper: the reference's `centralize`)."""
    out = np.zeros((len(M.rows), len(M.cols)))
    for k, (ir, ic) in enumerate(zip(M.rows.indices, M.cols.indices)):
        rp, cv, nz = M.download_csr(k)
        rowid = np.repeat(np.arange(ir.n_own), np.diff(rp))
        np.add.at(out, (ir.own_to_global[rowid] - 1, ic.local_to_global[cv] - 1), nz)
    return out


def test_reference_golden_a_times_a():
    import pa_b200 as pa

    I = [[1, 2, 1, 2, 2], [3, 3, 4, 6], [5, 5, 6, 7], [9, 9, 8, 10, 6]]
    J = [[2, 6, 1, 2, 1], [3, 9, 4, 2], [5, 6, 6, 7], [9, 3, 8, 10, 5]]
    V = [[1.0, 2.0, 30.0, 10.0, 1.0], [10.0, 2.0, 30.0, 2.0], [10.0, 2.0, 30.0, 1.0], [10.0, 2.0, 30.0, 50.0, 2.0]]
    for compress in ("host", "device"):
        b = pa.CUDAArray(4, arena_bytes=8 << 20)
        rows = pa.uniform_partition(b, 4, 10)
        A = pa.psparse(I, J, V, rows, rows, assembled=False, compress=compress)  # the reference's default: disassembled input
        Ad = centralize(A)
        B = pa.spmm(A, A)
        assert np.array_equal(centralize(B), Ad @ Ad)  # small integers: exact in any order, like the reference's == test
        # the product is a usable PSparseMatrix: mul! on it equals A*(A*x)
        x = pa.pvector_from_global(np.arange(1.0, 11.0), B.cols)
        y = pa.pzeros(B.rows)
        pa.mul_(y, B, x)
        assert np.array_equal(y.collect(), Ad @ (Ad @ np.arange(1.0, 11.0)))
        C = pa.spmtm(A, A)
        assert np.array_equal(centralize(C), Ad.T @ Ad)
        R = pa.rap(A, A, A)
        assert np.array_equal(centralize(R), Ad @ Ad @ Ad)
        b.close()


@pytest.mark.parametrize("P", [1, 3, 4])
def test_random_rectangular_products(P):
    import pa_b200 as pa
    from oracle import pa_oracle as o

    rng = np.random.default_rng(4 + P)
    n, m = 61, 23  # fine / coarse sizes: R is m x n, A is n x n, Pm is n x m  (rap = Galerkin product of AMG)
    b = pa.CUDAArray(P, arena_bytes=8 << 20)
    rn, rm = pa.uniform_partition(b, P, n), pa.uniform_partition(b, P, m)
    tn, tm = o.global_to_owner_table(o.uniform_partition(P, n)), o.global_to_owner_table(o.uniform_partition(P, m))

    def rand_matrix(nr, nc, per_row, rows_pr, cols_pr, tab):
        I = np.repeat(np.arange(1, nr + 1), per_row)
        J = rng.integers(1, nc + 1, size=len(I))
        V = rng.integers(-4, 5, size=len(I)).astype(float)
        parts = [(I[tab[I] == p + 1], J[tab[I] == p + 1], V[tab[I] == p + 1]) for p in range(P)]
        M = pa.psparse([q[0] for q in parts], [q[1] for q in parts], [q[2] for q in parts], rows_pr, cols_pr, assembled=True)
        D = np.zeros((nr, nc))
        np.add.at(D, (I - 1, J - 1), V)
        return M, D

    A, Ad = rand_matrix(n, n, 4, rn, rn, tn)
    R, Rd = rand_matrix(m, n, 5, rm, rn, tm)
    Pm, Pd = rand_matrix(n, m, 2, rn, rm, tn)
    assert np.array_equal(centralize(A), Ad) and np.array_equal(centralize(R), Rd)
    assert np.array_equal(centralize(pa.spmm(A, A)), Ad @ Ad)
    assert np.array_equal(centralize(pa.spmm(R, A)), Rd @ Ad)
    assert np.array_equal(centralize(pa.spmm(A, Pm)), Ad @ Pd)
    assert np.array_equal(centralize(pa.rap(R, A, Pm)), Rd @ Ad @ Pd)
    assert np.array_equal(centralize(pa.spmtm(Pm, A)), Pd.T @ Ad)       # P^T A
    assert np.array_equal(centralize(pa.spmtm(A, Pm)), Ad.T @ Pd)
    # non-integer values: rounding-level agreement (the summation order inside an entry is ascending local k)
    A2, A2d = rand_matrix(n, n, 4, rn, rn, tn)
    A2.fillstored_(0.1)
    A2d = (A2d != 0) * 0.0 + centralize(A2)
    np.testing.assert_allclose(centralize(pa.spmm(A2, A2)), A2d @ A2d, rtol=1e-14, atol=1e-15)
    b.close()
