"""Property tests (hypothesis) of the host-side partition and plan logic of the product (pa_b200.prange) against the oracle
and against size-independent invariants of the reference's definitions (src/p_range.jl:585-671, 806-818, 417-531):
every global id is owned exactly once, local_range tiles 1:n with the remainder on the LAST parts, ghost owners are
find_owner's, and the exchange plan is symmetric (what p sends to q is what q receives from p, in the same order)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import pa_oracle as o


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 9), st.integers(0, 60))
def test_local_range_tiles_the_range_with_remainder_on_the_last_parts(np_, n):
    from pa_b200 import prange as pr

    ranges = [pr.local_range(p, np_, n) for p in range(1, np_ + 1)]
    assert ranges == [o.local_range(p, np_, n) for p in range(1, np_ + 1)]
    nxt = 1
    for lo, hi in ranges:  # contiguous tiling of 1:n (src/p_range.jl:806-818)
        assert lo == nxt and hi >= lo - 1
        nxt = hi + 1
    assert nxt == n + 1
    lens = [hi - lo + 1 for lo, hi in ranges]
    assert max(lens) - min(lens) <= 1 and lens == sorted(lens)  # the longer blocks are the last ones


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(1, 3), st.integers(1, 7)), min_size=1, max_size=3), st.booleans())
def test_uniform_partition_owns_every_id_once_and_plans_are_symmetric(dims, ghost):
    from pa_b200 import prange as pr

    npd, n = tuple(d[0] for d in dims), tuple(max(d[1], d[0]) for d in dims)  # at least one node per part and direction
    P = int(np.prod(npd))
    g = tuple([ghost] * len(npd)) if ghost else None
    mine = [pr.uniform_partition_part(r, npd, n, g) for r in range(1, P + 1)]
    ref = o.uniform_partition(npd, n, g)
    owned = np.zeros(int(np.prod(n)), dtype=int)
    for a, b in zip(mine, ref):
        assert a.local_to_global.tolist() == b.local_to_global.tolist()
        assert a.local_to_owner.tolist() == b.local_to_owner.tolist()
        own = a.local_to_global[a.local_to_owner == a.part]
        owned[own - 1] += 1
    assert np.all(owned == 1)
    plans = pr.build_plans(mine, lambda objs: list(objs))
    oplan = o.assembly_plan(ref)
    for k, pl in enumerate(plans):
        assert pl.nbr_snd.tolist() == oplan.neighbors_snd[k].tolist() and pl.nbr_rcv.tolist() == oplan.neighbors_rcv[k].tolist()
        assert pl.snd_lids.tolist() == oplan.local_indices_snd[k].data.tolist()
        assert pl.rcv_lids.tolist() == oplan.local_indices_rcv[k].data.tolist()
        # symmetry by global id: my send list towards q == q's receive list from me
        pos = 0
        for i, q in enumerate(pl.nbr_snd):
            cnt = int(pl.snd_ptrs[i + 1] - pl.snd_ptrs[i])
            mine_g = mine[k].local_to_global[pl.snd_lids[pos : pos + cnt] - 1]
            pq = plans[q - 1]
            j = pq.nbr_rcv.tolist().index(k + 1)
            theirs = mine[q - 1].local_to_global[pq.rcv_lids[pq.rcv_ptrs[j] - 1 : pq.rcv_ptrs[j + 1] - 1] - 1]
            assert mine_g.tolist() == theirs.tolist()
            pos += cnt
