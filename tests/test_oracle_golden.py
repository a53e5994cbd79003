"""Pins the CPU oracle against the reference's own golden vectors / known-answer tests.
Citations: path:line in /root/reference (not read at run time — values are transcribed)."""
import numpy as np
import pytest

from oracle import pa_oracle as o
from oracle import c_oracle


def test_local_range_goldens():
    # test/p_range_tests.jl:7-15
    assert o.local_range(1, 3, 10) == (1, 3)
    assert o.local_range(2, 3, 10) == (4, 6)
    assert o.local_range(3, 3, 10) == (7, 10)
    assert o.local_range(1, 3, 10, True) == (1, 4)
    assert o.local_range(2, 3, 10, True) == (3, 7)
    assert o.local_range(3, 3, 10, True) == (6, 10)
    assert o.local_range(1, 3, 10, True, True) == (0, 4)
    assert o.local_range(2, 3, 10, True, True) == (3, 7)
    assert o.local_range(3, 3, 10, True, True) == (6, 11)


def _l2g(parts):
    return [p.local_to_global.tolist() for p in parts]


def test_uniform_partition_goldens():
    # src/p_range.jl:562-582 docstring
    assert _l2g(o.uniform_partition(4, 10)) == [[1, 2], [3, 4], [5, 6, 7], [8, 9, 10]]
    assert _l2g(o.uniform_partition((2, 2), (4, 4))) == [[1, 2, 5, 6], [3, 4, 7, 8], [9, 10, 13, 14], [11, 12, 15, 16]]
    # test/p_range_tests.jl:210-263
    assert _l2g(o.uniform_partition((2, 2), (5, 4))) == [[1, 2, 6, 7], [3, 4, 5, 8, 9, 10], [11, 12, 16, 17], [13, 14, 15, 18, 19, 20]]
    assert _l2g(o.uniform_partition((2, 2), (5, 4), (True, True))) == [
        [1, 2, 3, 6, 7, 8, 11, 12, 13],
        [2, 3, 4, 5, 7, 8, 9, 10, 12, 13, 14, 15],
        [6, 7, 8, 11, 12, 13, 16, 17, 18],
        [7, 8, 9, 10, 12, 13, 14, 15, 17, 18, 19, 20],
    ]
    assert _l2g(o.uniform_partition((2, 2), (4, 4), (True, True), (True, True))) == [
        [16, 13, 14, 15, 4, 1, 2, 3, 8, 5, 6, 7, 12, 9, 10, 11],
        [14, 15, 16, 13, 2, 3, 4, 1, 6, 7, 8, 5, 10, 11, 12, 9],
        [8, 5, 6, 7, 12, 9, 10, 11, 16, 13, 14, 15, 4, 1, 2, 3],
        [6, 7, 8, 5, 10, 11, 12, 9, 14, 15, 16, 13, 2, 3, 4, 1],
    ]
    assert _l2g(o.uniform_partition((2, 2), (4, 4), (True, True), (False, True))) == [
        [13, 14, 15, 1, 2, 3, 5, 6, 7, 9, 10, 11],
        [14, 15, 16, 2, 3, 4, 6, 7, 8, 10, 11, 12],
        [5, 6, 7, 9, 10, 11, 13, 14, 15, 1, 2, 3],
        [6, 7, 8, 10, 11, 12, 14, 15, 16, 2, 3, 4],
    ]


def test_find_owner_golden():
    # src/p_range.jl:317-344 docstring
    part = o.uniform_partition(4, 10)
    got = o.find_owner(part, [[3], [4, 5], [7, 2], [9, 10, 1]])
    assert [g.tolist() for g in got] == [[2], [2, 3], [3, 1], [4, 4, 1]]


def test_exchange_golden():
    # src/primitives.jl:889-919 docstring
    snd_ids = [[3, 4], [1, 3], [1, 4], [2]]
    rcv_ids = o.find_rcv_ids(snd_ids)
    assert rcv_ids == [[2, 3], [4], [1, 2], [1, 3]]
    snd = [o.jagged_from_lists([[v] for v in row], np.int64) for row in [[10, 10], [20, 20], [30, 30], [40]]]
    rcv = o.exchange(snd, snd_ids, rcv_ids)
    assert [r.data.tolist() for r in rcv] == [[20, 30], [40], [10, 20], [10, 30]]


def test_consistent_assemble_docstring_goldens():
    # src/p_vector.jl:666-693 (assemble!) and :719-745 (consistent!) : uniform_partition(rank,6,true)
    part = o.uniform_partition((2,), (6,), (True,))
    assert _l2g(part) == [[1, 2, 3, 4], [3, 4, 5, 6]]
    plan = o.assembly_plan(part)
    a = [np.ones(4), np.ones(4)]
    o.assemble(a, part, plan)
    assert a[0].tolist() == [1.0, 1.0, 2.0, 0.0] and a[1].tolist() == [0.0, 2.0, 1.0, 1.0]
    a = [np.full(4, 1.0), np.full(4, 2.0)]
    o.consistent(a, plan)
    assert a[0].tolist() == [1, 1, 1, 2] and a[1].tolist() == [1, 2, 2, 2]


def irregular_partition():
    # test/p_vector_tests.jl:95-107
    n = 10
    return [
        o.LocalIndices(n, 1, [1, 2, 3, 5, 7, 8], [1, 1, 1, 2, 3, 3]),
        o.LocalIndices(n, 2, [2, 4, 5, 10], [1, 2, 2, 4]),
        o.LocalIndices(n, 3, [6, 7, 8, 5, 4, 10], [3, 3, 3, 2, 2, 4]),
        o.LocalIndices(n, 4, [1, 3, 7, 9, 10], [1, 1, 3, 4, 4]),
    ]


def test_irregular_consistent_assemble_goldens():
    # test/p_vector_tests.jl:93-142
    part = irregular_partition()
    plan = o.assembly_plan(part)
    v = [np.where(ind.local_to_owner == ind.part, 10.0 * ind.part, 0.0) for ind in part]
    o.consistent(v, plan)
    for vals, ind in zip(v, part):
        assert vals.tolist() == (10.0 * ind.local_to_owner).tolist()
    v = [np.full(ind.n_local, 10.0) for ind in part]
    o.assemble(v, part, plan)
    assert v[0].tolist() == [20.0, 20.0, 20.0, 0.0, 0.0, 0.0]
    assert v[1].tolist() == [0.0, 20.0, 30.0, 0.0]
    assert v[2].tolist() == [10.0, 30.0, 20.0, 0.0, 0.0, 0.0]
    assert v[3].tolist() == [0.0, 0.0, 0.0, 10.0, 30.0]
    assert o.collect(v, part).tolist() == [20.0, 20.0, 20.0, 20.0, 30.0, 10.0, 30.0, 20.0, 10.0, 30.0]


def test_spmv_golden_7x6():
    # test/sparse_utils_tests.jl:14-45 : spmv! == mul! for I=[1,2,5,4,1],J=[3,6,1,1,3],V=[4,5,3,2,5]
    I, J, V = [1, 2, 5, 4, 1], [3, 6, 1, 1, 3], [4.0, 5.0, 3.0, 2.0, 5.0]
    A = o.sparse_matrix_csr(I, J, V, 7, 6)
    assert A.rowptr.tolist() == [1, 2, 3, 3, 4, 5, 5, 5]
    assert A.colval.tolist() == [3, 6, 1, 1] and A.nzval.tolist() == [9.0, 5.0, 2.0, 3.0]
    x = np.arange(1.0, 7.0)
    dense = np.zeros((7, 6))
    for i, j, v in zip(I, J, V):
        dense[i - 1, j - 1] += v
    want = dense @ x
    assert o.spmv_csr_py(A, x).tolist() == want.tolist()
    assert c_oracle.spmv_csr(A, x).tolist() == want.tolist()
    assert (A.to_scipy() @ x).tolist() == want.tolist()
    # spmtv! via the CSC kernel on the same arrays (src/sparse_utils.jl:625-631)
    xt = np.arange(1.0, 8.0)
    got = o.spmv_csc_py(6, A.rowptr, A.colval, A.nzval, xt)
    assert got.tolist() == (dense.T @ xt).tolist()


def test_skip_out_of_range_ids():
    # src/sparse_utils.jl:370-390: ids < 1 become a stored (1,1,0)
    A = o.sparse_matrix_csr([0, 2, -1], [1, 2, 3], [5.0, 7.0, 9.0], 3, 3)
    assert A.to_scipy().toarray().tolist() == [[0, 0, 0], [0, 7, 0], [0, 0, 0]]
    assert A.nnz == 2


def test_mul_known_answers_2I():
    # test/p_sparse_matrix_tests.jl:207-248: A = 2I (n=10, 4 parts), x=3 -> own values 6; after
    # consistent! all local 6.  fillstored!(A,1) -> 3 (:285-291)
    rows = o.uniform_partition(4, 10)
    I = [ind.own_to_global.copy() for ind in rows]
    V = [np.full(len(i), 2.0) for i in I]
    A = o.psparse(I, [i.copy() for i in I], V, rows, rows)
    plan = o.assembly_plan(A.col_partition)
    x = [np.full(ind.n_local, 3.0) for ind in A.col_partition]
    b = [np.zeros(ind.n_local) for ind in A.row_partition]
    o.pmul(A, x, plan, b)
    for vals, ind in zip(b, A.row_partition):
        assert np.all(o.own_values(vals, ind) == 6.0)
    o.consistent(b, o.assembly_plan(A.row_partition))
    assert all(np.all(v == 6.0) for v in b)
    for blk in A.own_own + A.own_ghost + A.local:
        blk.nzval[:] = 1.0
    o.pmul(A, x, plan, b)
    assert all(np.all(o.own_values(v, ind) == 3.0) for v, ind in zip(b, A.row_partition))


def irregular_coo():
    # test/p_sparse_matrix_tests.jl:306-316 (ids < 1 are skipped, rows owned elsewhere are shipped)
    return (
        [[1, 2, 1, 2, 2], [3, 3, 4, 6, 0], [5, 5, 6, 7], [9, 9, 8, 10, 6, -1]],
        [[2, 6, 1, 2, 1], [3, 9, 4, 2, 0], [5, 6, 6, 7], [9, 3, 8, 10, 5, 1]],
        [[1.0, 2.0, 30.0, 10.0, 1.0], [10.0, 2.0, 30.0, 2.0, 2.0], [10.0, 2.0, 30.0, 1.0], [10.0, 2.0, 30.0, 50.0, 2.0, 1.0]],
    )


def test_irregular_psparse_mul_matches_dense():
    rows = o.uniform_partition(4, 10)
    I, J, V = irregular_coo()
    A = o.psparse(I, J, V, rows, rows, assembled=False)
    dense = np.zeros((10, 10))
    for Ip, Jp, Vp in zip(I, J, V):
        for i, j, v in zip(Ip, Jp, Vp):
            if i >= 1 and j >= 1:
                dense[i - 1, j - 1] += v
    xg = np.arange(1.0, 11.0)
    plan = o.assembly_plan(A.col_partition)
    x = o.pvector_from_global(xg, A.col_partition, ghosts=False)
    b = [np.zeros(ind.n_local) for ind in A.row_partition]
    o.pmul(A, x, plan, b)
    np.testing.assert_allclose(o.collect(b, A.row_partition), dense @ xg, rtol=0, atol=1e-12)
    # CG residual < 1e-9 as in :334-345 (A is SPD-like diagonally dominant here? use the reference's check
    # on A*x_exact instead): solve with CG on the symmetrised system is not what the reference does;
    # it checks norm(A*x - b) after cg!; we check the oracle CG on A'A-free SPD generator below.


def test_hpcg_b_equals_collect_pb():
    # HPCG/test/hpcg_benchmark_tests.jl:20-28 : b (sequential 32x32x16) == collect(pb) on 2x2x1 parts of 16^3
    _, _, _, b_seq, _ = o.hpcg_build_matrix(32, 32, 16, 32, 32, 16, 1, 1, 1)
    A, pb = o.hpcg_build_p_matrix(16, 16, 16, 2, 2, 1)
    assert np.array_equal(o.collect(pb, A.col_partition), b_seq)
    # exact solution = ones: A*1 == b  (HPCG/src/sparse_matrix.jl:60-75)
    plan = o.assembly_plan(A.col_partition)
    x = [np.ones(ind.n_local) for ind in A.col_partition]
    c = [np.zeros(ind.n_local) for ind in A.col_partition]
    o.mul_no_lat(A, x, plan, c)
    assert np.array_equal(o.collect(c, A.col_partition), b_seq)


def test_laplacian_fdm_known_answer_and_split_vs_unsplit():
    # src/gallery.jl:36,65,75 (values pinned by source): (A*1)_i = alpha * (#missing neighbours)
    n = (6, 5, 4)
    I, J, V, rows, cols = o.laplacian_fdm(n, (2, 1, 2))
    A = o.psparse(I, J, V, rows, cols, assembled=True)
    alpha = 7 * 6 * 5
    plan = o.assembly_plan(A.col_partition)
    x = [np.ones(ind.n_local) for ind in A.col_partition]
    c = [np.zeros(ind.n_local) for ind in A.row_partition]
    o.pmul(A, x, plan, c)
    y = o.collect(c, A.row_partition).reshape(n[::-1])  # [z,y,x]
    ix, iy, iz = np.meshgrid(np.arange(6), np.arange(5), np.arange(4), indexing="ij")
    missing = sum(m.astype(float) for m in (ix == 0, ix == 5, iy == 0, iy == 4, iz == 0, iz == 3))
    assert np.array_equal(y, (alpha * missing).transpose(2, 1, 0))
    # split mul! == unsplit mul_no_lat! bitwise (sequential term order, own cols first)
    rng = np.random.default_rng(0)
    xg = rng.standard_normal(int(np.prod(n)))
    x1 = o.pvector_from_global(xg, A.col_partition, ghosts=False)
    x2 = [v.copy() for v in x1]
    c1 = [np.zeros(ind.n_local) for ind in A.row_partition]
    c2 = [np.zeros(ind.n_local) for ind in A.row_partition]
    o.pmul(A, x1, plan, c1)
    o.mul_no_lat(A, x2, plan, c2)
    assert np.array_equal(o.collect(c1, A.row_partition), o.collect(c2, A.row_partition))
    # vs scipy on the centralised operator
    import scipy.sparse as sp
    Ig, Jg, Vg = np.concatenate(I), np.concatenate(J), np.concatenate(V)
    Ag = sp.csr_matrix((Vg, (Ig - 1, Jg - 1)), shape=(len(xg), len(xg)))
    np.testing.assert_allclose(o.collect(c1, A.row_partition), Ag @ xg, rtol=1e-13, atol=1e-13 * alpha * 12)


def test_ref_cg_converges_and_c_twin_matches():
    # fdm_example-like: test/fdm_example.jl:128 (norm(x-x_exact) < 1e-5) on 9^3, parts (2,1,2)
    n = (9, 9, 9)
    I, J, V, rows, cols = o.laplacian_fdm(n, (2, 1, 2))
    A = o.psparse(I, J, V, rows, cols, assembled=True)
    part = A.col_partition
    plan = o.assembly_plan(part)
    xe = [np.ones(ind.n_local) for ind in part]
    b = [np.zeros(ind.n_local) for ind in part]
    o.pmul(A, xe, plan, b)
    x0 = [np.zeros(ind.n_local) for ind in part]
    x, r0, r, it, hist = o.ref_cg(A, b, x0, maxiter=200, tolerance=1e-12)
    err = np.sqrt(sum(np.sum((o.own_values(xv, ind) - 1.0) ** 2) for xv, ind in zip(x, part)))
    assert err < 1e-5 and r / r0 <= 1e-12
    # C twin: same residual history to round-off (different dot summation order than numpy's BLAS)
    mats = []
    for p, ind in enumerate(part):
        L = A.local[p]
        mats.append((ind.n_own, ind.n_local, L.rowptr.astype(np.int64) - 1, L.colval.astype(np.int32) - 1, L.nzval))
    prob = c_oracle.CGProblem(mats, plan, b, [np.zeros(ind.n_local) for ind in part])
    it2, hist2, _ = prob.cg(200, 1e-12)
    assert it2 == it
    np.testing.assert_allclose(hist2, hist, rtol=1e-8, atol=1e-13 * hist[0])


def test_c_stencil_generator_matches_psparse():
    for kind, npd, nloc in ((7, (2, 1, 2), (4, 5, 3)), (27, (2, 2, 1), (4, 3, 5)), (27, (2, 2, 2), (3, 3, 3))):
        gn = tuple(a * b for a, b in zip(npd, nloc))
        if kind == 7:
            I, J, V, rows, cols = o.laplacian_fdm(gn, npd)
            A = o.psparse(I, J, V, rows, cols, assembled=True)
        else:
            A, _ = o.hpcg_build_p_matrix(*nloc, *npd)
        for p, ind in enumerate(A.col_partition):
            lo = [r[0] - 1 for r in ind.box]
            hi = [r[1] for r in ind.box]
            rp, cv, nz, _ = c_oracle.stencil_csr(kind, gn, lo, hi, ind.ghost_to_global - 1)
            L = A.local[p]
            assert np.array_equal(rp, L.rowptr.astype(np.int64) - 1)
            assert np.array_equal(cv, L.colval - 1)
            assert np.array_equal(nz, L.nzval)


def test_hash_uniform_range_and_determinism():
    g = np.arange(1, 100001)
    a = o.hash_uniform(g, 7)
    assert a.min() >= -1.0 and a.max() < 1.0 and abs(a.mean()) < 0.02
    assert np.array_equal(a, o.hash_uniform(g, 7)) and not np.array_equal(a, o.hash_uniform(g, 8))


def test_periodic_single_part_ghosts_stay_ghosts():
    # block_with_constant_size (src/p_range.jl:620-671) decides own vs ghost by POSITION in the own ranges: with a periodic
    # ghost layer and one part in a direction the wrapped layer is owned by the part itself and is still a ghost layer.
    # Values derived from that source (np=1, n=6: local range 0:7 wraps to 6,1..6,1); no reference test holds them.
    (ind,) = o.uniform_partition((1,), (6,), (True,), (True,))
    assert ind.local_to_global.tolist() == [6, 1, 2, 3, 4, 5, 6, 1]
    assert (ind.n_own, ind.n_ghost) == (6, 2)
    assert ind.own_to_global.tolist() == [1, 2, 3, 4, 5, 6] and ind.ghost_to_owner.tolist() == [1, 1]
    assert ind.global_to_local([1, 6]).tolist() == [2, 7]  # the own id wins over the ghost copy
    snd, rcv = o.assembly_neighbors([ind])
    assert snd[0].tolist() == [] and rcv[0].tolist() == []  # owner == rank is skipped (src/p_range.jl:436-450)
    parts = o.uniform_partition((1, 2), (4, 4), (True, True), (True, True))
    assert [p.n_own for p in parts] == [8, 8] and [p.n_ghost for p in parts] == [16, 16]
    # the product's host mirror builds the same index sets and plans
    from pa_b200 import prange as pr

    for rank, po in enumerate(parts, start=1):
        pp = pr.uniform_partition_part(rank, (1, 2), (4, 4), (True, True), (True, True))
        assert pp.local_to_global.tolist() == po.local_to_global.tolist()
        assert pp.own_to_local.tolist() == po.own_to_local.tolist() and pp.ghost_to_owner.tolist() == po.ghost_to_owner.tolist()
        q = np.arange(0, 18)
        assert pp.global_to_local(q).tolist() == po.global_to_local(q).tolist()
    mine = [pr.uniform_partition_part(r, (1, 2), (4, 4), (True, True), (True, True)) for r in (1, 2)]
    plans = pr.build_plans(mine, lambda x: x)
    plan_o = o.assembly_plan(parts)
    for k in range(2):
        assert plans[k].nbr_snd.tolist() == plan_o.neighbors_snd[k].tolist() == [2 - k]
        assert plans[k].snd_lids.tolist() == plan_o.local_indices_snd[k].data.tolist()
        assert plans[k].rcv_lids.tolist() == plan_o.local_indices_rcv[k].data.tolist()
