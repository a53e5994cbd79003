"""BASELINE config C5: the reference's FEM example (test/fem_example.jl) -- disassembled triplets -> psparse -> pvector
-> cg -- through the product against the oracle, then at full size (3162^2 = 9 998 244 dofs, 4 parts) through
size-independent properties.  The triplets come from oracle/fem_q1.py (the restated example DRIVER, user-side code in
the reference); what is compared is the library path: assembly, ghost numbering, mul!, rhs assembly, CG."""
import numpy as np
import pytest

from oracle import fem_q1
from oracle import pa_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pa():
    import pa_b200

    return pa_b200


def _compare_with_oracle(pa, b, prob, fmt, compress, ship="host"):
    P = len(prob.I)
    rows = pa.variable_partition(b, prob.n_own_dofs, prob.n_global_dofs)
    A = pa.psparse(prob.I, prob.J, prob.V, rows, rows, assembled=False, local_format=fmt, compress=compress, ship=ship)
    orows = o.variable_partition(prob.n_own_dofs, prob.n_global_dofs)
    Ao = o.psparse(prob.I, prob.J, prob.V, orows, orows, assembled=False, local_format=fmt)
    for k in range(P):
        # same ghost columns in the same order => same term order in every row of the ghost block
        assert A.cols.indices[k].ghost_to_global.tolist() == Ao.col_partition[k].ghost_to_global.tolist()
        assert A.cols.indices[k].ghost_to_owner.tolist() == Ao.col_partition[k].ghost_to_owner.tolist()
        rp, cv, nz = A.download_csr(k)
        L = Ao.local[k]
        assert np.array_equal(rp, L.rowptr.astype(np.int64) - 1) and np.array_equal(cv, L.colval - 1)
        assert np.array_equal(nz, L.nzval)  # own sum, then each sender's sum: bit-identical
    # mul!
    rng = np.random.default_rng(3)
    xg = rng.standard_normal(prob.n_global_dofs)
    x = pa.pvector_from_global(xg, A.cols)
    y = pa.pzeros(A.rows)
    pa.mul_(y, A, x)
    xo = o.pvector_from_global(xg, Ao.col_partition, ghosts=False)
    yo = [np.zeros(r.n_local) for r in Ao.row_partition]
    o.pmul(Ao, xo, o.assembly_plan(Ao.col_partition), yo)
    for k, (got, want) in enumerate(zip(y.own_values(), yo)):
        assert np.array_equal(got, o.own_values(want, Ao.row_partition[k])), k
    # pvector(II,VV,rows)
    rhs = pa.pvector_from_triplets(prob.II, prob.VV, rows)
    rhs_o = o.pvector_disassembled(prob.II, prob.VV, orows)
    for got, want in zip(rhs.own_values(), rhs_o):
        assert np.array_equal(got, want)
    return A, rhs


@pytest.mark.parametrize("compress", ["host", "device"])
@pytest.mark.parametrize("fmt", ["csc", "csr"])
@pytest.mark.parametrize("parts,cells", [((2, 2), (10, 10)), ((3, 2), (13, 9)), ((1, 4), (5, 17))])
def test_fem_example_matches_oracle_and_known_answer(pa, parts, cells, fmt, compress):
    prob = fem_q1.Q1Problem(parts, cells, (2.0, 2.0 * cells[1] / cells[0]))
    b = pa.CUDAArray(len(prob.I), arena_bytes=16 << 20)
    A, rhs = _compare_with_oracle(pa, b, prob, fmt, compress)
    # x = cg(A,b); norm(x - x_exact) < 1e-5 (test/fem_example.jl:284-289)
    bc = pa.pzeros(A.cols)
    bc.copy_(rhs)
    x = pa.pzeros(A.cols)
    res = pa.ref_cg_(x, A, bc, tolerance=1e-10, maxiter=500)
    assert res.converged
    assert np.linalg.norm(x.collect() - prob.exact_solution()) < 1.0e-5
    b.close()


def test_fem_assembly_with_device_shipping_is_bit_identical(pa):
    """Same assembly with the ghost-row entries pulled by their owners through exchange! on the device (psparse(...; ship=
    "device")): identical CSR, ghost numbering and products as the oracle's stage-by-stage assembly."""
    prob = fem_q1.Q1Problem((3, 2), (13, 9), (2.0, 2.0 * 9 / 13))
    b = pa.CUDAArray(len(prob.I), arena_bytes=16 << 20)
    _compare_with_oracle(pa, b, prob, "csr", "device", ship="device")
    b.close()


def test_sender_sums_and_ghost_numbering_through_the_product(pa):
    """the two hand-checkable cases of tests/test_oracle_fem.py through the product"""
    b = pa.CUDAArray(2, arena_bytes=8 << 20)
    tiny = 1.0e-16
    rows = pa.uniform_partition(b, 2, 4)
    for compress in ("host", "device"):
        A = pa.psparse([[1], [1, 1]], [[1], [1, 1]], [[1.0], [tiny, tiny]], rows, rows, assembled=False, compress=compress)
        assert A.download_csr(0)[2].tolist() == [1.0 + 2 * tiny]
    v = pa.pvector_from_triplets([[1], [1, 1]], [[1.0], [tiny, tiny]], rows)
    assert v.own_values()[0][0] == 1.0 + 2 * tiny
    rows = pa.uniform_partition(b, 2, 6)
    I, J, V = [[3, 1, 2], [4]], [[5, 6, 5], [4]], [[9.0, 1.0, 2.0], [1.0]]
    csc = pa.psparse(I, J, V, rows, rows, assembled=False, local_format="csc")
    csr = pa.psparse(I, J, V, rows, rows, assembled=False, local_format="csr")
    assert csc.cols.indices[0].ghost_to_global.tolist() == [5, 6]
    assert csr.cols.indices[0].ghost_to_global.tolist() == [6, 5]
    b.close()


def test_full_size_c5_properties(pa):
    """3163^2 cells -> 3162^2 = 9 998 244 dofs on (2,2) parts, nnz = (3*3162-2)^2 = 89 946 256, rows of 4/6/9 entries."""
    n = 3162
    prob = fem_q1.Q1Problem((2, 2), (n + 1, n + 1), (2.0, 2.0))
    assert prob.n_global_dofs == n * n
    b = pa.CUDAArray(4, arena_bytes=2 << 30)
    rows = pa.variable_partition(b, prob.n_own_dofs, prob.n_global_dofs)
    A = pa.psparse(prob.I, prob.J, prob.V, rows, rows, assembled=False, local_format="csc", compress="device")
    assert sum(A.nnz(k) for k in range(4)) == (3 * n - 2) ** 2
    lens = np.concatenate([np.diff(A.download_csr(k)[0]) for k in range(4)])
    assert sorted(np.unique(lens).tolist()) == [4, 6, 9]
    assert np.count_nonzero(lens == 4) == 4 and np.count_nonzero(lens == 6) == 4 * (n - 2)
    # A * u_exact = rhs exactly in exact arithmetic (Q1 reproduces x1 + x2): the discrete residual is at rounding level
    rhs = pa.pvector_from_triplets(prob.II, prob.VV, rows)
    xe = prob.exact_solution()
    x = pa.pvector_from_global(xe, A.cols)
    y = pa.pzeros(A.rows)
    pa.mul_(y, A, x)
    r = y.collect() - rhs.collect()
    scale = np.abs(prob.Ae).max() * np.abs(xe).max()
    assert np.abs(r).max() < 64 * np.finfo(float).eps * scale
    # linearity + symmetry of the assembled operator: <A u, v> == <u, A v>
    rng = np.random.default_rng(0)
    u = pa.pvector_from_global(rng.standard_normal(n * n), A.cols)
    v = pa.pvector_from_global(rng.standard_normal(n * n), A.cols)
    Au, Av = pa.pzeros(A.rows), pa.pzeros(A.rows)
    pa.mul_(Au, A, u); pa.mul_(Av, A, v)
    uc, vc = pa.pzeros(A.rows), pa.pzeros(A.rows)
    uc.copy_(u); vc.copy_(v)
    d1, d2 = Au.dot(vc), Av.dot(uc)
    assert abs(d1 - d2) <= 1e-12 * max(abs(d1), abs(d2), 1e-300) + 1e-9 * np.abs(prob.Ae).max()
    # CG makes progress on the 10M-dof system (a full solve needs O(n) iterations: not a unit test)
    bc = pa.pzeros(A.cols)
    bc.copy_(rhs)
    xs = pa.pzeros(A.cols)
    res = pa.ref_cg_(xs, A, bc, tolerance=0.0, maxiter=200)
    assert res.iters == 200 and np.isfinite(res.residual) and res.residual < res.residual0
    b.close()
