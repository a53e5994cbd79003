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
