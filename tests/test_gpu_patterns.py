"""Row-pattern compression of the SpMV column stream (csrc/pa_spmv.cu, build_patterns / k_spmv_tma<..., PAT>): one byte per row
instead of four bytes per entry where rows repeat a few (length, column - row) tuples.  Same products in the same order: the
results must equal the plain kernel's and the oracle's (spmv_csr! order, src/sparse_utils.jl:649-669) bit for bit."""
import numpy as np
import pytest

from oracle import pa_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pa():
    import pa_b200

    return pa_b200


def _mul_all(pa, A, seed, flags=0):
    x = pa.fill_hash(pa.PVector(A.cols), seed)
    y = pa.pzeros(A.rows)
    y.fill_(-3.0)
    pa.mul_(y, A, x, flags=flags)
    out = y.collect()
    x.free(); y.free()
    return out


@pytest.mark.parametrize("kind,gn,npd", [(7, (24, 20, 18), (1, 1, 1)), (7, (24, 20, 18), (2, 1, 2)), (27, (16, 18, 14), (2, 2, 1)),
                                         (27, (12, 12, 12), (2, 2, 2)), (7, (64, 64, 64), (2, 1, 1))])
def test_pattern_kernel_equals_plain_kernel_and_oracle(pa, kind, gn, npd):
    P = int(np.prod(npd))
    outs = {}
    for mode in ("patterns", "plain"):
        b = pa.CUDAArray(P, arena_bytes=64 << 20)
        b.set_knob("spmv_patterns", 1 if mode == "patterns" else 0)
        b.set_knob("spmv_pattern_min_rows", 1)
        A, rhs = pa.stencil_matrix(kind, gn, npd, b)
        outs[mode] = [_mul_all(pa, A, 5), _mul_all(pa, A, 6, pa.PA_SPMV_OVERLAP)]
        # 5-argument mul! and the CG loop (SpMV with the fused dot epilogue) on the same path
        x = pa.fill_hash(pa.PVector(A.cols), 9)
        y = pa.fill_hash(pa.PVector(A.rows), 10)
        pa.mul_(y, A, x, 0.5, -2.0)
        outs[mode].append(y.collect())
        xs = pa.pzeros(A.cols)
        res = pa.ref_cg_(xs, A, rhs, tolerance=0.0, maxiter=12)
        outs[mode].append(np.asarray(res.history))
        outs[mode].append(xs.collect())
        b.close()
    for got, want in zip(outs["patterns"][:3], outs["plain"][:3]):  # mul! (3- and 5-argument, two schedules): bit for bit
        assert np.array_equal(got, want)
    # CG: the fused dot adds per-tile partial sums, and the pattern kernel may use other tiles (two rows per thread): the
    # histories agree to rounding, not to the bit
    np.testing.assert_allclose(outs["patterns"][3], outs["plain"][3], rtol=1e-12)
    np.testing.assert_allclose(outs["patterns"][4], outs["plain"][4], rtol=1e-9, atol=1e-12)
    # and against the oracle's sequential loops
    if kind == 7:
        I, J, V, rows, cols = o.laplacian_fdm(gn, npd)
        Ao = o.psparse(I, J, V, rows, cols, assembled=True, local_format="csr")
    else:
        nloc = tuple(g // q for g, q in zip(gn, npd))
        Ao, _ = o.hpcg_build_p_matrix(*nloc, *npd)
    xg = o.hash_uniform(np.arange(1, int(np.prod(gn)) + 1), 5)
    xo = o.pvector_from_global(xg, Ao.col_partition, ghosts=False)
    co = [np.zeros(ind.n_local) for ind in Ao.row_partition]
    o.mul_no_lat(Ao, xo, o.assembly_plan(Ao.col_partition), co)
    assert np.array_equal(outs["patterns"][0], o.collect(co, Ao.row_partition))


def test_irregular_rows_fall_back_to_the_column_stream(pa):
    """Random sparsity: (almost) every row is its own pattern, the matrix keeps the plain column stream (or marks the rows
    as escapes) — same bits either way; a banded matrix with a few odd rows mixes table rows and escape rows in one tile."""
    rng = np.random.default_rng(3)
    n = 6000
    b = pa.CUDAArray(1, arena_bytes=32 << 20)
    b.set_knob("spmv_pattern_min_rows", 1)
    rows = pa.uniform_partition(b, 1, n)
    # banded part: tridiagonal everywhere ...
    I = np.concatenate([np.arange(1, n + 1), np.arange(2, n + 1), np.arange(1, n)])
    J = np.concatenate([np.arange(1, n + 1), np.arange(1, n), np.arange(2, n + 1)])
    V = rng.standard_normal(len(I))
    # ... plus 150 rows with extra random entries (their patterns are not in the table)
    odd = rng.choice(np.arange(10, n - 10), size=150, replace=False) + 1
    I = np.concatenate([I, np.repeat(odd, 3)])
    J = np.concatenate([J, rng.integers(1, n + 1, size=3 * len(odd))])
    V = np.concatenate([V, rng.standard_normal(3 * len(odd))])
    b.set_knob("spmv_patterns", 1)
    A = pa.psparse([I], [J], [V], rows, rows, assembled=True, local_format="csr")
    x = rng.standard_normal(n)
    xv = pa.PVector(A.cols).set_local_values([x])
    y = pa.pzeros(A.rows)
    pa.mul_(y, A, xv)
    got = y.collect()
    rp, cv, nz = A.download_csr(0)
    want = np.zeros(n)
    for i in range(n):  # spmv_csr!: sequential, separate multiply and add
        acc = 0.0
        for p in range(rp[i], rp[i + 1]):
            acc = acc + nz[p] * x[cv[p]]
        want[i] = acc
    assert np.array_equal(got, want)
    b.close()
