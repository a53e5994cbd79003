"""mul!(c, transpose(A), b[, alpha, beta]) (src/p_sparse_matrix.jl:2144-2162) against the oracle and a dense product."""
import numpy as np
import pytest

from oracle import pa_oracle as o

pytestmark = pytest.mark.gpu


def test_transpose_mul_matches_oracle_and_dense():
    import pa_b200 as pa

    rng = np.random.default_rng(11)
    n, P = 400, 4
    orows = o.uniform_partition(P, n)
    tab = o.global_to_owner_table(orows)
    lens = rng.choice([1, 3, 6], size=n)
    I = np.repeat(np.arange(1, n + 1), lens)
    J = rng.integers(1, n + 1, size=len(I))
    V = rng.standard_normal(len(I))
    Is = [I[tab[I] == p + 1] for p in range(P)]
    Js = [J[tab[I] == p + 1] for p in range(P)]
    Vs = [V[tab[I] == p + 1] for p in range(P)]
    Ao = o.psparse(Is, Js, Vs, orows, orows, assembled=True)
    dense = np.zeros((n, n))
    np.add.at(dense, (I - 1, J - 1), V)
    bg, cg = rng.standard_normal(n), rng.standard_normal(n)
    bk = pa.CUDAArray(P, arena_bytes=16 << 20)
    rows = pa.uniform_partition(bk, P, n)
    A = pa.psparse(Is, Js, Vs, rows, rows, assembled=True)
    for alpha, beta in ((1.0, 0.0), (0.5, -2.0)):
        b = pa.pvector_from_global(bg, A.rows)
        c = pa.pvector_from_global(cg, A.cols, ghosts=True)
        pa.mul_transpose_(c, A, b, alpha, beta)
        bo = o.pvector_from_global(bg, Ao.row_partition)
        co = o.pvector_from_global(cg, Ao.col_partition, ghosts=True)
        o.pmul_transpose(Ao, bo, co, alpha, beta)
        got = c.local_values()
        np.testing.assert_allclose(c.collect(), alpha * (dense.T @ bg) + beta * cg, rtol=1e-12, atol=1e-12)
        for k, ind in enumerate(Ao.col_partition):
            if (alpha, beta) == (1.0, 0.0):
                assert np.array_equal(got[k][: ind.n_own], co[k][: ind.n_own])  # same summation order as the oracle
            else:
                np.testing.assert_allclose(got[k][: ind.n_own], co[k][: ind.n_own], rtol=1e-13, atol=1e-13)
            assert np.all(got[k][ind.n_own:] == 0.0)  # ghosts zeroed by assemble!
        b.free(); c.free()
    # (A^T)^T x == A x on the symmetric 7-pt operator: transpose product equals the forward product
    S, rhs = pa.stencil_matrix(7, (6, 5, 4), (2, 2, 1), bk)
    x = pa.fill_hash(pa.PVector(S.cols), 2)
    y1, y2 = pa.pzeros(S.rows), pa.pzeros(S.cols)
    pa.mul_(y1, S, x)
    pa.mul_transpose_(y2, S, x)
    np.testing.assert_allclose(y2.collect(), y1.collect(), rtol=1e-13, atol=1e-9)
    bk.close()
