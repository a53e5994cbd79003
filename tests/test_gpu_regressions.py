"""Regression tests for defects found by review."""
import numpy as np
import pytest

from oracle import pa_oracle as o

pytestmark = pytest.mark.gpu


def test_cg_default_arguments_stop_at_exact_zero_residual():
    """ref_cg!(x, A, b) with its defaults (tolerance = 0.0, maxiter = length(b)): the reference stops as soon as
    residual/residual0 <= 0, i.e. at an exactly zero residual (HPCG/src/ref_cg.jl:19-26).  A = 2I converges exactly
    in one iteration; the device-resident loop must not run on into 0/0."""
    import pa_b200 as pa

    for nparts in (1, 4):
        bk = pa.CUDAArray(nparts, arena_bytes=8 << 20)
        rows = pa.uniform_partition(bk, nparts, 10 * nparts)
        I = [ind.own_to_global.copy() for ind in rows.indices]
        A = pa.psparse(I, I, [np.full(len(i), 2.0) for i in I], rows, rows)
        for flags in (0, pa.PA_CG_REFERENCE_OPS):
            b = pa.pfill(6.0, A.cols)
            x = pa.pzeros(A.cols)
            res = pa.ref_cg_(x, A, b, flags=flags)  # defaults: tolerance 0.0, maxiter = len(cols)
            assert res.converged and res.iters == 1, (res.iters, res.history)
            assert res.residual == 0.0
            assert np.array_equal(x.collect(), np.full(10 * nparts, 3.0))
        bk.close()


def test_transpose_product_sees_refreshed_values():
    """mul!(c, transpose(A), b) after psparse!(A, V, cache) / fillstored!(A, a): the cached local transposes copy the
    values and must be rebuilt (src/p_sparse_matrix.jl:1291-1305, :2144-2162)."""
    import pa_b200 as pa

    rng = np.random.default_rng(3)
    n, P = 60, 3
    bk = pa.CUDAArray(P, arena_bytes=8 << 20)
    rows = pa.uniform_partition(bk, P, n)
    tab = o.global_to_owner_table(o.uniform_partition(P, n))
    I = np.repeat(np.arange(1, n + 1), 3)
    J = rng.integers(1, n + 1, size=len(I))
    V = rng.standard_normal(len(I))
    Is, Js, Vs = ([a[tab[I] == p + 1] for p in range(P)] for a in (I, J, V))
    A = pa.psparse(Is, Js, Vs, rows, rows, assembled=True, compress="device")
    bg = rng.standard_normal(n)

    def dense(vals):
        d = np.zeros((n, n))
        np.add.at(d, (I - 1, J - 1), vals)
        return d

    def tmul():
        b = pa.pvector_from_global(bg, A.rows)
        c = pa.pzeros(A.cols)
        pa.mul_transpose_(c, A, b)
        out = c.collect()
        b.free(); c.free()
        return out

    np.testing.assert_allclose(tmul(), dense(V).T @ bg, rtol=1e-12, atol=1e-12)
    V2 = rng.standard_normal(len(I))
    A.update_coo_values_([V2[tab[I] == p + 1] for p in range(P)])
    np.testing.assert_allclose(tmul(), dense(V2).T @ bg, rtol=1e-12, atol=1e-12)
    A.fillstored_(1.0)
    pattern = (dense(np.ones(len(I))) != 0).astype(float)
    np.testing.assert_allclose(tmul(), pattern.T @ bg, rtol=1e-12, atol=1e-12)
    bk.close()


def test_periodic_single_part_direction_consistent_and_reductions():
    """uniform_partition with periodic ghosts and ONE part in a direction: the wrapped layer is owned by the part itself
    and is still a ghost layer (src/p_range.jl:620-671); reductions count own entries once and consistent! leaves the
    self-owned ghosts alone (compute_assembly_neighbors skips owner == rank, :436-450)."""
    import pa_b200 as pa

    npd, n = (1, 2), (4, 4)
    parts = o.uniform_partition(npd, n, (True, True), (True, True))
    plan = o.assembly_plan(parts)
    bk = pa.CUDAArray(2, arena_bytes=8 << 20)
    rows = pa.uniform_partition(bk, npd, n, (True, True), (True, True))
    assert [i.n_own for i in rows.indices] == [8, 8] and [i.n_ghost for i in rows.indices] == [16, 16]
    v = pa.pvector(lambda ind: np.where(np.isin(np.arange(1, ind.n_local + 1), ind.own_to_local), 10.0 * ind.part, -1.0), rows)
    vo = [np.where(p.own_mask, 10.0 * p.part, -1.0) for p in parts]
    assert v.sum() == 8 * 10.0 + 8 * 20.0
    v.consistent_().wait()
    o.consistent(vo, plan)
    for got, want in zip(v.local_values(), vo):
        assert np.array_equal(got, want)
    v.assemble_().wait()
    o.assemble(vo, parts, plan)
    for got, want in zip(v.local_values(), vo):
        assert np.array_equal(got, want)
    bk.close()


def test_subassembled_mul_matches_the_oracle_and_the_assembled_product():
    """psparse(I,J,V,rows,cols; assemble=false) keeps the rows a part does not own as ghost rows; mul!(c,A,b) multiplies own and
    ghost rows and finishes with assemble!(c) (src/p_sparse_matrix.jl:2105-2142).  The 4-part matrix of
    test/p_sparse_matrix_tests.jl:131-160 (exact in any order: small integers) and a random one against the oracle."""
    import pa_b200 as pa

    I = [[1, 2, 1, 2, 2], [3, 3, 4, 6], [5, 5, 6, 7], [9, 9, 8, 10, 6]]
    J = [[2, 6, 1, 2, 1], [3, 9, 4, 2], [5, 6, 6, 7], [9, 3, 8, 10, 5]]
    V = [[1.0, 2.0, 30.0, 10.0, 1.0], [10.0, 2.0, 30.0, 2.0], [10.0, 2.0, 30.0, 1.0], [10.0, 2.0, 30.0, 50.0, 2.0]]
    n, P = 10, 4
    rng = np.random.default_rng(2)
    cases = [(I, J, V, n)]
    n2 = 57
    tab = o.global_to_owner_table(o.uniform_partition(P, n2))
    I2 = [rng.integers(1, n2 + 1, 40) for _ in range(P)]   # rows owned anywhere: plenty of ghost rows
    J2 = [rng.integers(1, n2 + 1, 40) for _ in range(P)]
    V2 = [rng.standard_normal(40) for _ in range(P)]
    cases.append((I2, J2, V2, n2))
    for Ic, Jc, Vc, nn in cases:
        bk = pa.CUDAArray(P, arena_bytes=8 << 20)
        rows = pa.uniform_partition(bk, P, nn)
        A = pa.psparse(Ic, Jc, Vc, rows, rows, assembled=False, assemble=False)
        assert A.assembled is False
        orows = o.uniform_partition(P, nn)
        Ao = o.psparse_subassembled(Ic, Jc, Vc, orows, orows, local_format="csr")
        dense = np.zeros((nn, nn))
        for i, j, v in zip(Ic, Jc, Vc):
            np.add.at(dense, (np.asarray(i) - 1, np.asarray(j) - 1), v)
        xg = rng.integers(-3, 4, nn).astype(float)
        for alpha, beta in ((1.0, 0.0), (2.0, -1.0)):
            x = pa.pvector_from_global(xg, A.cols)
            c = pa.pfill(1.0, A.rows)
            pa.mul_(c, A, x, alpha, beta)
            xo = o.pvector_from_global(xg, Ao.col_partition, ghosts=False)
            co = [np.ones(r.n_local) for r in Ao.row_partition]
            o.pmul_subassembled(Ao, xo, co, alpha, beta)
            got = c.local_values()
            for k, r in enumerate(Ao.row_partition):
                if (alpha, beta) == (1.0, 0.0):
                    assert np.array_equal(got[k], co[k]), k      # own entries summed, ghost entries zeroed by assemble!
                else:
                    np.testing.assert_allclose(got[k], co[k], rtol=1e-13, atol=1e-13)
            # ghost rows carry beta*c_ghost too before assemble! (every ghost copy of c was 1): compare with the oracle only
            if (alpha, beta) == (1.0, 0.0):
                np.testing.assert_allclose(c.collect(), dense @ xg, rtol=1e-13, atol=1e-12)
            x.free(); c.free()
        bk.close()
