"""CPU-only: the host-side stage of psparse's per-part compression (parrays._stored_entries = sparse_matrix(I,J,V,m,n) of
SparseArrays / SparseMatricesCSR as called by src/p_sparse_matrix.jl:1186-1222): entries in storage order, duplicates added in
INPUT order (the order fixes the bits of the sums), ids < 1 -> one stored (1,1,0.0)."""
import numpy as np
import pytest

import pa_b200
from pa_b200 import parrays


def _reference(li, lj, v, fmt):
    """The definition, entry by entry: a dict filled in input order."""
    acc = {}
    for i, j, x in zip(li.tolist(), lj.tolist(), v.tolist()):
        if i < 1 or j < 1:
            i, j, x = 1, 1, 0.0
        acc[(i, j)] = acc.get((i, j), 0.0) + x  # sequential: ((v1 + v2) + v3) ...
    keys = sorted(acc, key=(lambda t: (t[0], t[1])) if fmt == "csr" else (lambda t: (t[1], t[0])))
    return np.array([k[0] for k in keys]), np.array([k[1] for k in keys]), np.array([acc[k] for k in keys])


@pytest.mark.parametrize("fmt", ["csr", "csc"])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_stored_entries_follow_the_definition(fmt, seed):
    rng = np.random.default_rng(seed)
    n = 4000
    li = rng.integers(0, 40, size=n)   # 0 = an id < 1: skipped like FilteredCooVector does
    lj = rng.integers(0, 35, size=n)
    v = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, size=n)  # sums whose bits depend on the order
    ei, ej, ev = parrays._stored_entries(li.copy(), lj.copy(), v.copy(), 40, 35, fmt)
    ri, rj, rv = _reference(li, lj, v, fmt)
    assert np.array_equal(ei, ri) and np.array_equal(ej, rj)
    assert np.array_equal(ev, rv)  # bit for bit: same association of the duplicate sums


def test_stable_order_host_path_is_lexsort():
    rng = np.random.default_rng(5)
    a, b = rng.integers(0, 50, 10000), rng.integers(0, 50, 10000)
    assert np.array_equal(parrays._stable_order(a, b, None), np.lexsort((b, a)))


def test_empty_and_degenerate_inputs():
    z = np.zeros(0, dtype=np.int64)
    ei, ej, ev = parrays._stored_entries(z, z, np.zeros(0), 5, 5, "csr")
    assert len(ei) == len(ej) == len(ev) == 0
    ei, ej, ev = parrays._stored_entries(np.array([0, -3]), np.array([2, 0]), np.array([1.5, 2.5]), 4, 4, "csr")
    assert ei.tolist() == [1] and ej.tolist() == [1] and ev.tolist() == [0.0]


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_global_to_local_dense_table_equals_the_sorted_table(seed):
    """LocalIndices.global_to_local switches to a dense gid -> lid table for long queries: same answers as the sorted-table
    path, including duplicate gids among the ghosts (periodic layers: the LAST ghost copy wins, an own id wins over any ghost,
    src/p_range.jl:928-935) and ids outside 1..n_global."""
    from pa_b200 import prange as pr

    rng = np.random.default_rng(seed)
    ng, part = 5000, 3
    own = rng.choice(np.arange(1, ng + 1), size=600, replace=False)
    ghosts = rng.choice(np.setdiff1d(np.arange(1, ng + 1), own), size=300, replace=False)
    ghosts = np.concatenate([ghosts, ghosts[:40], own[:10]])  # repeated ghost gids and ghost copies of own gids
    l2g = np.concatenate([own, ghosts])
    perm = rng.permutation(len(l2g)) if seed % 2 else np.arange(len(l2g))  # own-first and permuted local orders
    owner = np.concatenate([np.full(len(own), part), rng.integers(4, 9, size=len(ghosts))]).astype(np.int32)
    ind = pr.LocalIndices(ng, part, l2g[perm], owner[perm])
    q_short = rng.integers(-5, ng + 10, size=1000)
    q_long = rng.integers(-5, ng + 10, size=(1 << 16) + 17)
    want = ind.global_to_local(q_short)           # sorted-table path
    got_long = ind.global_to_local(q_long)        # dense path
    assert getattr(ind, "_g2l_dense", None) is not None
    ref = pr.LocalIndices(ng, part, l2g[perm], owner[perm])
    ref_long = np.concatenate([ref.global_to_local(q_long[i : i + 1000]) for i in range(0, len(q_long), 1000)])  # sorted path, in chunks
    assert np.array_equal(got_long, ref_long)
    assert np.array_equal(ind.global_to_local(q_short), want)
