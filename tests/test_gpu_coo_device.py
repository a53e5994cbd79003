"""On-device COO -> CSR compression and value refresh against the host path and the oracle:
sparse_matrix(I,J,V; reuse=true) / sparse_matrix!(A,V,K) (src/sparse_utils.jl:392-469), psparse / psparse!."""
import numpy as np
import pytest

from oracle import pa_oracle as o

pytestmark = pytest.mark.gpu


def _random_coo(rng, n, P, tab, dup=3):
    lens = rng.choice([0, 2, 5, 9], size=n)
    I = np.repeat(np.arange(1, n + 1), lens)
    J = rng.integers(1, n + 1, size=len(I))
    # duplicates (assembly-like) and a few out-of-range ids that must be skipped
    sel = rng.integers(0, len(I), size=len(I) // 2)
    I = np.concatenate([I] + [I[sel]] * dup)
    J = np.concatenate([J] + [J[sel]] * dup)
    V = rng.standard_normal(len(I))
    I[rng.integers(0, len(I), 7)] = 0
    J[rng.integers(0, len(J), 5)] = -1
    perm = rng.permutation(len(I))
    I, J, V = I[perm], J[perm], V[perm]
    Is, Js, Vs = [], [], []
    for p in range(P):
        m = (I < 1) | (tab[np.where(I >= 1, I, 1)] == p + 1)
        m &= (np.arange(len(I)) % P == p) | (I >= 1)  # spread the skipped ones over the parts
        m = np.where(I >= 1, tab[np.where(I >= 1, I, 1)] == p + 1, np.arange(len(I)) % P == p)
        Is.append(I[m]); Js.append(J[m]); Vs.append(V[m])
    return Is, Js, Vs


def test_device_compress_equals_host_and_oracle_and_refresh():
    import pa_b200 as pa

    rng = np.random.default_rng(7)
    n, P = 3000, 4
    orows = o.uniform_partition(P, n)
    tab = o.global_to_owner_table(orows)
    Is, Js, Vs = _random_coo(rng, n, P, tab)
    b = pa.CUDAArray(P, arena_bytes=16 << 20)
    rows = pa.uniform_partition(b, P, n)
    A_host = pa.psparse(Is, Js, Vs, rows, rows, assembled=True, compress="host", split_format=False)
    A_dev = pa.psparse(Is, Js, Vs, rows, rows, assembled=True, compress="device")
    Ao = o.psparse(Is, Js, Vs, orows, orows, assembled=True)
    for k in range(P):
        rh, ch, zh = A_host.download_csr(k)
        rd, cd, zd = A_dev.download_csr(k)
        assert np.array_equal(rh, rd) and np.array_equal(ch, cd)
        assert np.array_equal(zh, zd)  # duplicates added in input order on both paths: bit-identical
        L = Ao.local[k]
        assert np.array_equal(rd, L.rowptr.astype(np.int64) - 1) and np.array_equal(cd, L.colval - 1) and np.array_equal(zd, L.nzval)
    xg = rng.standard_normal(n)
    x = pa.pvector_from_global(xg, A_dev.cols)
    y1, y2 = pa.pzeros(A_dev.rows), pa.pzeros(A_host.rows)
    pa.mul_(y1, A_dev, x)
    x2 = pa.pvector_from_global(xg, A_host.cols)
    pa.mul_(y2, A_host, x2)
    assert np.array_equal(y1.collect(), y2.collect())
    # psparse!: new values, same pattern
    Vs2 = [rng.standard_normal(len(v)) for v in Vs]
    A_dev.update_coo_values_(Vs2)
    A_ref = pa.psparse(Is, Js, Vs2, rows, rows, assembled=True, compress="host", split_format=False)
    for k in range(P):
        assert np.array_equal(A_dev.download_csr(k)[2], A_ref.download_csr(k)[2])
    pa.mul_(y1, A_dev, x)
    x3 = pa.pvector_from_global(xg, A_ref.cols)
    pa.mul_(y2, A_ref, x3)
    assert np.array_equal(y1.collect(), y2.collect())
    b.close()


def test_device_stable_order_equals_lexsort():
    """pa_sort_perm_u64 behind parrays._stable_order (the host-side assembly stages sort their triplets on the device once they
    are long): the permutation is np.lexsort's — stable, so duplicates keep their input order and the sums their bits."""
    import pa_b200 as pa
    from pa_b200 import parrays

    rng = np.random.default_rng(2)
    n = (1 << 18) + 12345
    b = pa.CUDAArray(1, arena_bytes=8 << 20)
    rows = rng.integers(1, 5000, size=n)
    cols = np.clip(rows + rng.integers(-3, 4, size=n), 1, None)  # many duplicates
    got = parrays._stable_order(rows, cols, b)
    assert got.dtype == np.int32 and np.array_equal(got, np.lexsort((cols, rows)))
    v = rng.standard_normal(n)
    ei, ej, ev = parrays._stored_entries(rows.copy(), cols.copy(), v, 5000, 5010, "csr", b)
    hi, hj, hv = parrays._stored_entries(rows.copy(), cols.copy(), v, 5000, 5010, "csr", None)
    assert np.array_equal(ei, hi) and np.array_equal(ej, hj) and np.array_equal(ev, hv)
    # ids the packed key cannot hold fall back to the host sort
    big = rows.astype(np.int64) + (1 << 33)
    assert np.array_equal(parrays._stable_order(big, cols, b), np.lexsort((cols, big)))
    b.close()
