"""psparse with disassembled input (the reference's default mode): rows owned by other parts are shipped to their owner.
Mirrors test/p_sparse_matrix_tests.jl:306-345 (irregular COO with out-of-range ids, then CG with ||A x - y|| < 1e-9)."""
import numpy as np
import pytest

from oracle import pa_oracle as o

pytestmark = pytest.mark.gpu

I = [[1, 2, 1, 2, 2], [3, 3, 4, 6, 0], [5, 5, 6, 7], [9, 9, 8, 10, 6, -1]]
J = [[2, 6, 1, 2, 1], [3, 9, 4, 2, 0], [5, 6, 6, 7], [9, 3, 8, 10, 5, 1]]
V = [[1.0, 2.0, 30.0, 10.0, 1.0], [10.0, 2.0, 30.0, 2.0, 2.0], [10.0, 2.0, 30.0, 1.0], [10.0, 2.0, 30.0, 50.0, 2.0, 1.0]]


@pytest.mark.parametrize("ship", ["host", "device"])  # ghost-row entries through the metadata channel / the device exchange!
@pytest.mark.parametrize("fmt", ["csr", "csc"])
def test_disassembled_psparse_mul_and_cg(fmt, ship):
    import pa_b200 as pa

    b = pa.CUDAArray(4, arena_bytes=16 << 20)
    rows = pa.uniform_partition(b, 4, 10)
    A = pa.psparse(I, J, V, rows, rows, assembled=False, local_format=fmt, ship=ship)
    dense = np.zeros((10, 10))
    for Ip, Jp, Vp in zip(I, J, V):
        for i, j, v in zip(Ip, Jp, Vp):
            if i >= 1 and j >= 1:
                dense[i - 1, j - 1] += v
    # same operator as the oracle's semantic psparse (by global id)
    orows = o.uniform_partition(4, 10)
    Ao = o.psparse(I, J, V, orows, orows, assembled=False)
    for k in range(4):
        assert sorted(A.cols.indices[k].ghost_to_global.tolist()) == sorted(Ao.col_partition[k].ghost_to_global.tolist())
    x = pa.pones(A.cols)
    y = pa.pzeros(A.rows)
    pa.mul_(y, A, x)                                   # y = A*x
    np.testing.assert_allclose(y.collect(), dense @ np.ones(10), rtol=0, atol=1e-12)
    yc = pa.pvector_from_global(y.collect(), A.cols)   # rhs on the column partition (square operator)
    xs = pa.pzeros(A.cols)
    res = pa.ref_cg_(xs, A, yc, tolerance=1e-14, maxiter=100)
    r = pa.pzeros(A.rows)
    pa.mul_(r, A, xs)
    assert np.linalg.norm(r.collect() - y.collect()) < 1e-9   # test/p_sparse_matrix_tests.jl:334-336
    np.testing.assert_allclose(xs.collect(), np.ones(10), atol=1e-9)
    b.close()
