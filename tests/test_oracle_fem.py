"""Oracle check of the disassembled assembly path (psparse / pvector defaults) on the reference's FEM example
(BASELINE config C5): pinned by the reference's own known answer norm(x - x_exact) < 1e-5 (test/fem_example.jl:289),
plus hand-checkable cases for the two things the assembly stages decide: how sums associate and how ghost columns are
numbered."""
import numpy as np
import pytest

from oracle import fem_q1
from oracle import pa_oracle as o


def _global_dense(A: o.PSparse, n):
    d = np.zeros((n, n))
    for p, (oo, og) in enumerate(zip(A.own_own, A.own_ghost)):
        r, c = A.row_partition[p], A.col_partition[p]
        for blk, col_g in ((oo, c.own_to_global), (og, c.ghost_to_global)):
            rows = np.repeat(np.arange(blk.m), np.diff(blk.rowptr.astype(np.int64)))
            d[r.own_to_global[rows] - 1, col_g[blk.colval - 1] - 1] += blk.nzval
    return d


@pytest.mark.parametrize("fmt", ["csc", "csr"])
def test_fem_example_known_answer(fmt):
    prob = fem_q1.Q1Problem((2, 2), (10, 10))
    assert prob.n_global_dofs == 81 and sorted(prob.n_own_dofs) == [16, 20, 20, 25]
    rows = o.variable_partition(prob.n_own_dofs, prob.n_global_dofs)
    A = o.psparse(prob.I, prob.J, prob.V, rows, rows, assembled=False, local_format=fmt)
    # same operator as adding every triplet into a dense matrix
    dense = np.zeros((81, 81))
    for I, J, V in zip(prob.I, prob.J, prob.V):
        np.add.at(dense, (I - 1, J - 1), V)
    np.testing.assert_allclose(_global_dense(A, 81), dense, rtol=0, atol=1e-15)
    # Q1 stiffness stencil: 9 entries per interior dof, 6 next to an edge, 4 in a corner
    lens = np.concatenate([np.diff(a.rowptr.astype(np.int64)) + np.diff(b.rowptr.astype(np.int64)) for a, b in zip(A.own_own, A.own_ghost)])
    assert sorted(np.unique(lens).tolist()) == [4, 6, 9]
    b_own = o.pvector_disassembled(prob.II, prob.VV, rows)
    bg = np.zeros(81)
    for II, VV in zip(prob.II, prob.VV):
        np.add.at(bg, II - 1, VV)
    np.testing.assert_allclose(np.concatenate(b_own), bg, rtol=0, atol=1e-15)
    # cg, then the reference's check against u(x) = x1 + x2 (fem_example.jl:284-289)
    cols = A.col_partition
    b_vals = [np.concatenate([bo, np.zeros(c.n_ghost)]) for bo, c in zip(b_own, cols)]
    x_vals = [np.zeros(c.n_local) for c in cols]
    x_vals, r0, r, iters, _ = o.ref_cg(A, b_vals, x_vals, maxiter=200, tolerance=1e-10, mul=o.pmul)
    x = np.concatenate([o.own_values(v, c) for v, c in zip(x_vals, cols)])
    assert np.linalg.norm(x - prob.exact_solution()) < 1.0e-5
    assert iters < 60


def test_sender_contributions_are_combined_before_they_are_added():
    """(own + own) + (sender + sender), not ((own + own) + sender) + sender: src/p_sparse_matrix.jl:1196 compresses each part's
    triplets before assemble (:1651-1703) appends what the owners receive."""
    rows = o.uniform_partition(2, 4)
    tiny = 1.0e-16
    I = [[1], [1, 1]]
    J = [[1], [1, 1]]
    V = [[1.0], [tiny, tiny]]
    A = o.psparse(I, J, V, rows, rows, assembled=False)
    assert A.own_own[0].nzval.tolist() == [1.0 + 2 * tiny] != [(1.0 + tiny) + tiny]
    v = o.pvector_disassembled([[1], [1, 1]], [[1.0], [tiny, tiny]], rows)
    assert v[0][0] == 1.0 + 2 * tiny


def test_ghost_columns_are_numbered_by_storage_order_then_by_sender():
    """cols_fa = union_ghost(own cols, columns of [findnz(own_ghost) ++ received]) (src/p_sparse_matrix.jl:1667-1676,1739):
    the own part lists its ghost columns in the STORAGE order of its sub-assembled own_ghost block, which differs
    between SparseMatrixCSC (column-major over first-appearance ghost ids) and SparseMatrixCSR (row-major)."""
    rows = o.uniform_partition(2, 6)  # part 1 owns 1:3, part 2 owns 4:6
    # part 1: row 2 meets ghost col 6 first, then row 1 meets ghost col 5, row 1 col 6; part 2 contributes (3,4) to part 1
    I = [[2, 1, 1], [3, 4]]
    J = [[6, 5, 6], [4, 4]]
    V = [[1.0, 2.0, 3.0], [4.0, 5.0]]
    csc = o.psparse(I, J, V, rows, rows, assembled=False, local_format="csc")
    csr = o.psparse(I, J, V, rows, rows, assembled=False, local_format="csr")
    # sub-assembled ghost ids of part 1: 6 -> 1, 5 -> 2.  CSC walks columns 6 (rows 1,2) then 5; CSR walks row 1 (cols by
    # ghost id: 6 then 5), then row 2 -> both start with 6; the received column 4 comes last
    assert csc.col_partition[0].ghost_to_global.tolist() == [6, 5, 4]
    assert csr.col_partition[0].ghost_to_global.tolist() == [6, 5, 4]
    # a case where they differ: first appearance (ghost ids) 5 -> 1, 6 -> 2, but row 1 only has col 6 and row 2 only col 5
    I = [[3, 1, 2], [4]]
    J = [[5, 6, 5], [4]]
    V = [[9.0, 1.0, 2.0], [1.0]]
    # row 3 col 5 is an own row too: rows 1..3 are own. CSC: col 5 (rows 2,3), col 6 (row 1) -> [5, 6]; CSR: row 1 (6), row 2 (5) -> [6, 5]
    csc = o.psparse(I, J, V, rows, rows, assembled=False, local_format="csc")
    csr = o.psparse(I, J, V, rows, rows, assembled=False, local_format="csr")
    assert csc.col_partition[0].ghost_to_global.tolist() == [5, 6]
    assert csr.col_partition[0].ghost_to_global.tolist() == [6, 5]
    x = np.arange(1.0, 7.0)
    for A in (csc, csr):
        np.testing.assert_array_equal(_global_dense(A, 6) @ x, np.array([6.0, 10.0, 45.0, 4.0, 0.0, 0.0]))


def test_product_side_example_driver_emits_the_oracle_triplets():
    """pa_b200.fem_example (the per-part driver bench.py uses for config C5) against oracle/fem_q1.py."""
    from pa_b200 import fem_example as fe

    for parts, cells in (((2, 2), (10, 10)), ((3, 2), (13, 9)), ((1, 4), (5, 17))):
        lengths = (2.0, 2.0 * cells[1] / cells[0])
        prob = fem_q1.Q1Problem(parts, cells, lengths)
        lay = fe.Q1Layout(parts, cells, lengths)
        assert lay.n_own_dofs == prob.n_own_dofs and lay.n_global_dofs == prob.n_global_dofs
        xe = prob.exact_solution()
        for rank in range(1, len(prob.I) + 1):
            got = fe.q1_part(lay, rank)
            want = (prob.I[rank - 1], prob.J[rank - 1], prob.V[rank - 1], prob.II[rank - 1], prob.VV[rank - 1])
            for g, w in zip(got, want):
                assert np.array_equal(g, w)
            off = int(lay.offset[rank - 1])
            assert np.array_equal(lay.exact_own(rank), xe[off : off + lay.n_own_dofs[rank - 1]])
