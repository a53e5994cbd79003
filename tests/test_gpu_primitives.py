"""GPU parity for the backend primitives next to the PVector path: exchange!, assemble! with an
operation, reduce / maximum / minimum / norm(v, p) -- against the reference's golden vectors and the oracle."""
import numpy as np
import pytest

from oracle import pa_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pa():
    import pa_b200

    return pa_b200


def test_exchange_docstring_and_reference_tests(pa):
    b = pa.CUDAArray(4, arena_bytes=8 << 20)
    snd_ids = [[3, 4], [1, 3], [1, 4], [2]]
    # src/primitives.jl:889-919 (docstring): scalars
    graph = pa.ExchangeGraph(b, snd_ids)
    assert graph.rcv == [[2, 3], [4], [1, 2], [1, 3]]  # default_find_rcv_ids, test/primitives_tests.jl:164-174
    rcv = pa.exchange([[10, 10], [20, 20], [30, 30], [40]], graph)
    assert [[int(v[0]) for v in r] for r in rcv] == [[20, 30], [40], [10, 20], [10, 30]]
    assert all(v.dtype == np.int64 for r in rcv for v in r)
    # test/primitives_tests.jl:190-205: snd = 10*snd_ids with an explicit receive side
    graph = pa.ExchangeGraph(b, snd_ids, [[2, 3], [4], [1, 2], [1, 3]])
    rcv = pa.exchange([[10 * q for q in s] for s in snd_ids], graph)
    assert [[int(v[0]) for v in r] for r in rcv] == [[10, 10], [20], [30, 30], [40, 40]]
    # test/primitives_tests.jl:221-234: vector payloads collect(1:j)
    rcv = pa.exchange([[np.arange(1, j + 1) for j in s] for s in snd_ids], graph)
    assert [[v.tolist() for v in r] for r in rcv] == [[[1], [1]], [[1, 2]], [[1, 2, 3], [1, 2, 3]], [[1, 2, 3, 4], [1, 2, 3, 4]]]
    # Float64 payloads travel bit for bit
    rcv = pa.exchange([[np.array([0.1 * q, -1e300]) for q in s] for s in snd_ids], graph)
    assert rcv[0][0].dtype == np.float64 and rcv[2][1].tolist() == [0.1 * 3, -1e300]
    b.close()


@pytest.mark.parametrize("dtype", [np.int32, np.float32, np.int16, np.uint8, np.int64])
def test_exchange_payload_element_types(pa, dtype):
    """exchange! with Int32 payloads (the element type of the reference's index lists, src/p_range.jl:489-531 — odd segment
    lengths and offsets, so 4-byte segments start off 8-byte boundaries), Float32 and narrower types: same values, same type."""
    rng = np.random.default_rng(11)
    P = 5
    b = pa.CUDAArray(P, arena_bytes=16 << 20)
    snd_ids = [[2, 3, 5], [1], [1, 2, 4, 5], [], [3]]
    hi = 100 if np.dtype(dtype).itemsize == 1 else 30000
    segs = [[(rng.integers(0, hi, size=int(rng.integers(1, 700)) | 1)).astype(dtype) for _ in s] for s in snd_ids]
    graph = pa.ExchangeGraph(b, snd_ids)
    want = o.exchange([o.jagged_from_lists(s, dtype) for s in segs], snd_ids, graph.rcv)
    got = pa.exchange(segs, graph)
    for r in range(P):
        assert len(got[r]) == len(graph.rcv[r])
        for i in range(len(got[r])):
            assert got[r][i].dtype == np.dtype(dtype)
            assert np.array_equal(got[r][i], want[r].segment(i))
    b.close()


def test_exchange_random_graph_matches_oracle(pa):
    rng = np.random.default_rng(7)
    P = 6
    b = pa.CUDAArray(P, arena_bytes=32 << 20)
    snd_ids = [sorted(rng.choice([q for q in range(1, P + 1) if q != p], size=rng.integers(0, P - 1), replace=False).tolist())
               for p in range(1, P + 1)]
    snd_ids[2] = []  # a part that sends nothing (but may receive)
    segs = [[rng.standard_normal(int(rng.integers(0, 5000))) for _ in s] for s in snd_ids]
    graph = pa.ExchangeGraph(b, snd_ids)
    assert graph.rcv == o.find_rcv_ids(snd_ids)
    want = o.exchange([o.jagged_from_lists(s, np.float64) for s in segs], snd_ids, graph.rcv)
    for _ in range(2):  # twice: the arena slot is returned and handed out again
        got = pa.exchange(segs, graph)
        for r in range(P):
            assert len(got[r]) == len(graph.rcv[r])
            for i in range(len(got[r])):
                assert np.array_equal(got[r][i], want[r].segment(i))
    # vectors created around exchanges keep working (symmetric arena bookkeeping)
    rows = pa.uniform_partition(b, P, 60, True)
    v = pa.pones(rows)
    v.assemble_().wait()
    assert v.sum() == 60.0 + 2 * (P - 1)
    b.close()


def test_assemble_with_operations(pa):
    """assemble!(op, v) (src/p_vector.jl:699-708): op applied in neighbour order; insert(a,b) = b (:755)."""
    b = pa.CUDAArray(4, arena_bytes=8 << 20)
    parts = [([1, 2, 3, 5, 7, 8], [1, 1, 1, 2, 3, 3]), ([2, 4, 5, 10], [1, 2, 2, 4]), ([6, 7, 8, 5, 4, 10], [3, 3, 3, 2, 2, 4]),
             ([1, 3, 7, 9, 10], [1, 1, 3, 4, 4])]  # test/p_vector_tests.jl:95-107
    rows = pa.PRange(b, [pa.LocalIndices(10, p + 1, g, w) for p, (g, w) in enumerate(parts)])
    op_part = o_partition = [o.LocalIndices(10, p + 1, np.array(g), np.array(w, dtype=np.int32)) for p, (g, w) in enumerate(parts)]
    plan = o.assembly_plan(o_partition)
    rng = np.random.default_rng(11)
    vals = [rng.standard_normal(len(g)) for g, _ in parts]
    for op, f in (("+", lambda a, c: a + c), ("max", max), ("min", min), ("insert", lambda a, c: c)):
        v = pa.PVector(rows).set_local_values([x.copy() for x in vals])
        v.assemble_(op).wait()
        want = [x.copy() for x in vals]
        o.assemble(want, o_partition, plan, f)
        for got, w in zip(v.local_values(), want):
            assert np.array_equal(got, w), op
    b.close()


def test_reduce_maximum_minimum_norm_p(pa):
    """reduce(op,a), maximum, minimum (test/p_vector_tests.jl:172-175: == 16 and 0), norm(a,p) (src/p_vector.jl:1201-1206)."""
    b = pa.CUDAArray(4, arena_bytes=64 << 20)
    rows = pa.uniform_partition(b, (2, 2), (4, 4))  # test/p_vector_tests.jl:144-175: v[i] = global id, then v[1] .= 0
    v = pa.pvector(lambda ind: np.where(ind.local_to_global == 1, 0.0, ind.local_to_global.astype(float)), rows)
    assert v.maximum() == 16.0 and v.minimum() == 0.0
    assert v.reduce("+") == v.sum() == float(sum(range(2, 17)))
    # a larger permuted (halo) layout: ghosts never contribute
    n = 4 * 50_021
    rows = pa.uniform_partition(b, 4, n, True)
    xg = np.random.default_rng(2).standard_normal(n)
    x = pa.pvector_from_global(xg, rows)
    assert x.maximum() == xg.max() and x.minimum() == xg.min()
    assert abs(x.norm(1) - np.abs(xg).sum()) <= 1e-12 * np.abs(xg).sum()
    assert abs(x.norm(3) - np.linalg.norm(xg, 3)) <= 1e-12 * np.linalg.norm(xg, 3)
    assert abs(x.norm(2) - np.linalg.norm(xg)) <= 1e-12 * np.linalg.norm(xg)
    assert x.reduce("+") == x.reduce("+")  # deterministic
    with pytest.raises(ValueError):
        x.norm(0.5)
    b.close()


def _compress(I, J, V, m, n, fmt):
    """compresscoo (src/sparse_utils.jl:313-350) on the host: entries sorted, duplicates summed in input order; 0-based."""
    maj, mnr, nmaj = (I, J, m) if fmt == "csr" else (J, I, n)
    order = sorted(range(len(I)), key=lambda t: (maj[t], mnr[t]))  # stable
    ptr_, idx, val = [0] * (nmaj + 1), [], []
    last = None
    for t in order:
        if last == (maj[t], mnr[t]):
            val[-1] = val[-1] + V[t]
        else:
            idx.append(mnr[t] - 1)
            val.append(V[t])
            ptr_[maj[t]] += 1
            last = (maj[t], mnr[t])
    return np.cumsum([0] + ptr_[1:]).tolist(), idx, val


@pytest.mark.parametrize("Tv,Ti", [(np.float64, np.int64), (np.float32, np.int32), (np.float64, np.int32)])
@pytest.mark.parametrize("fmt,base", [("csc", 1), ("csr", 1), ("csr", 0)])
def test_local_spmv_spmtv_golden_7x6(pa, fmt, base, Tv, Ti):
    """test/sparse_utils_tests.jl:14-45,113-118: I=[1,2,5,4,1], J=[3,6,1,1,3], V=[4,5,3,2,5], 7x6, x=1:n; spmv! == mul!,
    spmtv! == transpose mul!, for SparseMatrixCSC / SparseMatrixCSR{1} / SparseMatrixCSR{0} x (Float64,Int) / (Float32,Int32)."""
    b = pa.CUDAArray(1, arena_bytes=1 << 20)
    I, J, V, m, n = [1, 2, 5, 4, 1], [3, 6, 1, 1, 3], [4, 5, 3, 2, 5], 7, 6
    p0, i0, v0 = _compress(I, J, [Tv(v) for v in V], m, n, fmt)
    ptr_, idx, val = np.array(p0, dtype=Ti) + base, np.array(i0, dtype=Ti) + base, np.array(v0, dtype=Tv)
    dense = np.zeros((m, n))
    for i, j, v in zip(I, J, V):
        dense[i - 1, j - 1] += v
    x = np.arange(1, n + 1, dtype=Tv)
    got = pa.spmv_(b, fmt, ptr_, idx, val, x, m, n, index_base=base)
    assert got.dtype == Tv and got.tolist() == (dense @ x).tolist()  # small integers: exact in both precisions
    xt = np.arange(1, m + 1, dtype=Tv)
    got = pa.spmtv_(b, fmt, ptr_, idx, val, xt, m, n, index_base=base)
    assert got.dtype == Tv and got.tolist() == (dense.T @ xt).tolist()
    with pytest.raises(pa.PAError):
        pa._capi.check(pa._capi.lib().pa_local_spmv(b.h, 0, base, ptr_.dtype.itemsize * 8, val.dtype.itemsize * 8, len(ptr_) - 1, 3,
                                                    pa._capi.ptr(ptr_), pa._capi.ptr(idx), pa._capi.ptr(val), pa._capi.ptr(x), len(x),
                                                    pa._capi.ptr(np.zeros(3, dtype=Tv))))  # length(b) != size(A,1)
    b.close()


@pytest.mark.parametrize("Tv", [np.float64, np.float32])
def test_local_spmv_random_bit_exact_vs_sequential_loops(pa, Tv):
    """Random rectangular matrices: same bits as the reference loops run in the same element type (spmv_csr!: bi += aij*xj
    in storage order; spmv_csc!: b[row] += aij*xj column by column)."""
    rng = np.random.default_rng(5)
    b = pa.CUDAArray(1, arena_bytes=1 << 20)
    m, n, nnz = 301, 257, 4000
    I, J = rng.integers(1, m + 1, nnz).tolist(), rng.integers(1, n + 1, nnz).tolist()
    V = rng.standard_normal(nnz).astype(Tv)
    for fmt in ("csr", "csc"):
        p0, i0, v0 = _compress(I, J, list(V), m, n, fmt)
        ptr_, idx, val = np.array(p0, dtype=np.int32) + 1, np.array(i0, dtype=np.int32) + 1, np.array(v0, dtype=Tv)
        for transpose in (False, True):
            x = rng.standard_normal(m if transpose else n).astype(Tv)
            gather = (fmt == "csr") != transpose  # spmv_csr! on the stored arrays, else spmv_csc!
            nb = n if transpose else m
            want = np.zeros(nb, dtype=Tv)
            if gather:
                for i in range(len(ptr_) - 1):
                    acc = Tv(0)
                    for p in range(ptr_[i] - 1, ptr_[i + 1] - 1):
                        acc = Tv(acc + Tv(val[p] * x[idx[p] - 1]))
                    want[i] = acc
            else:
                for j in range(len(ptr_) - 1):
                    for p in range(ptr_[j] - 1, ptr_[j + 1] - 1):
                        want[idx[p] - 1] = Tv(want[idx[p] - 1] + Tv(val[p] * x[j]))
            f = pa.spmtv_ if transpose else pa.spmv_
            got = f(b, fmt, ptr_, idx, val, x, m, n)
            assert np.array_equal(got, want), (fmt, transpose)
    b.close()
