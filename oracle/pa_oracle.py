"""CPU ORACLE — test infrastructure, NOT product code.

A numpy / pure-Python restatement of the PartitionedArrays.jl algorithms that sit
on the PSparseMatrix x PVector hot path (partition, ghost discovery, exchange plan,
consistent!/assemble!, COO->CSR, split mul!, HPCG mul_no_lat!, dot/norm, CG) and
of the two input generators used by the benchmark configs.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module.  The product package
(``partitionedarrays.jl_b200``) never does.

Everything here is 1-based exactly like the Julia reference (global ids, part ids,
local ids, JaggedArray ptrs) so that the reference's golden vectors can be compared
verbatim.  Citations are ``path:line`` relative to /root/reference.

Parity status: PINNED against the reference's own golden vectors (see
tests/test_oracle_golden.py): local_range (test/p_range_tests.jl:7-15),
uniform_partition (test/p_range_tests.jl:210-263, src/p_range.jl:562-582),
consistent!/assemble! (src/p_vector.jl:666-693,719-745, test/p_vector_tests.jl:93-142),
exchange (src/primitives.jl:889-919), spmv (test/sparse_utils_tests.jl:14-45),
mul! (test/p_sparse_matrix_tests.jl:207-291), HPCG b (HPCG/test/hpcg_benchmark_tests.jl:20-28).
NOT pinned by the reference (no value-level test exists): laplacian_fdm values
(src/gallery.jl, pinned by source only), OpenBLAS dot/nrm2 summation order,
SparseMatricesCSR 0.6 5-arg mul! and IterativeSolvers 0.9 cg! (third-party, absent
from /root/reference).  The reference itself (Julia) cannot run in this environment.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# JaggedArray  (src/jagged_array.jl:11-32,107-122): data + 1-based ptrs, ptrs[i]:ptrs[i+1]-1
# --------------------------------------------------------------------------------------


@dataclass
class JaggedArray:
    data: np.ndarray
    ptrs: np.ndarray  # int32, 1-based, len = nsegments+1

    def __len__(self):
        return len(self.ptrs) - 1

    def segment(self, i: int) -> np.ndarray:
        """0-based segment index -> view of data."""
        return self.data[self.ptrs[i] - 1 : self.ptrs[i + 1] - 1]

    def tolist(self):
        return [self.segment(i).tolist() for i in range(len(self))]


def length_to_ptrs(lengths: Sequence[int]) -> np.ndarray:
    """src/jagged_array.jl:107-114 (length_to_ptrs!) — returns 1-based ptrs."""
    ptrs = np.ones(len(lengths) + 1, dtype=np.int32)
    if len(lengths):
        ptrs[1:] = 1 + np.cumsum(np.asarray(lengths, dtype=np.int64))
    return ptrs


def jagged_from_lists(lists, dtype) -> JaggedArray:
    ptrs = length_to_ptrs([len(l) for l in lists])
    if len(lists) and ptrs[-1] > 1:
        data = np.concatenate([np.asarray(l, dtype=dtype) for l in lists]).astype(dtype)
    else:
        data = np.zeros(0, dtype=dtype)
    return JaggedArray(data, ptrs)


# --------------------------------------------------------------------------------------
# local_range  (src/p_range.jl:806-818)
# --------------------------------------------------------------------------------------


def local_range(p: int, np_: int, n: int, ghost: bool = False, periodic: bool = False) -> Tuple[int, int]:
    """Inclusive 1-based (start, stop) of part p (1-based) of np_ over 1:n."""
    l, rem = divmod(n, np_)
    offset = l * (p - 1)
    if rem >= (np_ - p + 1):
        l += 1
        offset += p - (np_ - rem) - 1
    g = 1 if ghost else 0
    start = 1 + offset - g
    stop = l + offset + g
    if periodic:
        return start, stop
    return max(1, start), min(n, stop)


# --------------------------------------------------------------------------------------
# Local indices (AbstractLocalIndices API, src/p_range.jl:32-160)
# --------------------------------------------------------------------------------------


@dataclass
class LocalIndices:
    """Generic local index set of one part.

    Mirrors ``LocalIndices`` (src/p_range.jl:1100-1210): arbitrary local order given by
    local_to_global / local_to_owner.  Block partitions and permuted block partitions
    are expressed through the same arrays (own-first or Cartesian-with-halo order).
    """

    n_global: int
    part: int  # 1-based
    local_to_global: np.ndarray  # int64, 1-based gids
    local_to_owner: np.ndarray  # int32, 1-based part ids
    # optional block info (own box, 1-based inclusive ranges per dim; global dims; parts per dim)
    box: Optional[Tuple[Tuple[int, int], ...]] = None
    grid: Optional[Tuple[int, ...]] = None
    parts_per_dir: Optional[Tuple[int, ...]] = None
    _g2l: Optional[dict] = field(default=None, repr=False, compare=False)
    # own/ghost is decided by POSITION in the own ranges (src/p_range.jl:648-664), not by owner == part: with a
    # periodic ghost layer and one part in a direction the wrapped ghosts are owned by the part itself and stay ghosts
    own_mask: Optional[np.ndarray] = field(default=None, repr=False, compare=False)

    def __post_init__(self):
        self.local_to_global = np.asarray(self.local_to_global, dtype=np.int64)
        self.local_to_owner = np.asarray(self.local_to_owner, dtype=np.int32)
        if self.own_mask is None:
            self.own_mask = self.local_to_owner == self.part
        self.own_mask = np.asarray(self.own_mask, dtype=bool)

    # --- accessors (1-based values) ---
    @property
    def n_local(self):
        return len(self.local_to_global)

    @property
    def own_to_local(self) -> np.ndarray:
        return (np.nonzero(self.own_mask)[0] + 1).astype(np.int32)

    @property
    def ghost_to_local(self) -> np.ndarray:
        return (np.nonzero(~self.own_mask)[0] + 1).astype(np.int32)

    @property
    def n_own(self):
        return int(np.count_nonzero(self.own_mask))

    @property
    def n_ghost(self):
        return self.n_local - self.n_own

    @property
    def own_to_global(self):
        return self.local_to_global[self.own_to_local - 1]

    @property
    def ghost_to_global(self):
        return self.local_to_global[self.ghost_to_local - 1]

    @property
    def ghost_to_owner(self):
        return self.local_to_owner[self.ghost_to_local - 1]

    def global_to_local(self, gids) -> np.ndarray:
        """gid -> lid, 0 when absent (src/p_range.jl GlobalToLocal)."""
        gids = np.atleast_1d(np.asarray(gids, dtype=np.int64))
        if self.n_local > 200000:  # large parts (bench-size CPU baseline): sorted-table lookup instead of a Python dict
            if self._g2l is None:
                # duplicates: the own id wins over a (self-owned, periodic) ghost copy, else the LAST ghost (Dict order)
                rank = -np.arange(1, self.n_local + 1, dtype=np.int64)
                rank[self.own_mask] = -(self.n_local + 1)
                order = np.lexsort((rank, self.local_to_global))
                self._g2l = (self.local_to_global[order], order)
            sg, order = self._g2l
            out = np.zeros(len(gids), dtype=np.int32)
            pos = np.clip(np.searchsorted(sg, gids, side="left"), 0, len(sg) - 1)
            hit = sg[pos] == gids
            out[hit] = order[pos[hit]] + 1
            return out
        if self._g2l is None:  # own ids win over a (self-owned, periodic) ghost copy of the same gid
            self._g2l = {int(g): i + 1 for i, g in enumerate(self.local_to_global) if not self.own_mask[i]}
            self._g2l.update({int(g): i + 1 for i, g in enumerate(self.local_to_global) if self.own_mask[i]})
        return np.array([self._g2l.get(int(g), 0) for g in gids], dtype=np.int32)

    def own_is_prefix(self) -> bool:
        no = self.n_own
        return bool(np.all(self.own_mask[:no]))


def _cartesian_linear(idx: Sequence[np.ndarray], dims: Sequence[int]) -> np.ndarray:
    """Column-major LinearIndices (first dim fastest), 1-based in / 1-based out
    (src/p_range.jl:1477-1480; Appendix A.1)."""
    lin = np.zeros_like(idx[0], dtype=np.int64)
    stride = 1
    for d in range(len(dims)):
        lin = lin + (idx[d].astype(np.int64) - 1) * stride
        stride *= int(dims[d])
    return lin + 1


def _box_ids(ranges: Sequence[Tuple[int, int]], dims: Sequence[int], wrap: bool = False):
    """All grid points of a box in CartesianIndices order (first dim fastest).
    Returns (per-dim index arrays (1-based, possibly out of 1..n when wrap), linear gids)."""
    axes = [np.arange(a, b + 1, dtype=np.int64) for (a, b) in ranges]
    # first dim fastest => meshgrid with indexing 'ij' over reversed axes then reverse back
    grids = np.meshgrid(*axes[::-1], indexing="ij")[::-1]
    flat = [g.reshape(-1) for g in grids]
    if wrap:
        wrapped = [np.mod(f - 1, int(n)) + 1 for f, n in zip(flat, dims)]
    else:
        wrapped = flat
    return flat, _cartesian_linear(wrapped, dims)


def _part_linear(pcoord: Sequence[int], np_: Sequence[int]) -> int:
    """Part id = column-major linear index of the part's Cartesian coordinate
    (src/p_range.jl:617,1679-1682)."""
    lin, stride = 0, 1
    for c, n in zip(pcoord, np_):
        lin += (c - 1) * stride
        stride *= n
    return lin + 1


def _part_coord(rank: int, np_: Sequence[int]) -> Tuple[int, ...]:
    r = rank - 1
    out = []
    for n in np_:
        out.append(r % n + 1)
        r //= n
    return tuple(out)


def _owner_1d(np_: int, n: int) -> np.ndarray:
    """owner (1-based part coordinate) of each 1-based grid index along one dim."""
    o = np.zeros(n + 1, dtype=np.int32)
    for p in range(1, np_ + 1):
        a, b = local_range(p, np_, n)
        o[a : b + 1] = p
    return o


def uniform_partition(
    np_: Sequence[int],
    n: Sequence[int],
    ghost: Optional[Sequence[bool]] = None,
    periodic: Optional[Sequence[bool]] = None,
) -> List[LocalIndices]:
    """src/p_range.jl:585-671 (uniform_partition / block_with_constant_size).

    Without ``ghost``: own ids of each block in CartesianIndices order, no ghosts.
    With ``ghost``: one halo layer; local order = Cartesian order over the *local*
    ranges (PermutedLocalIndices, :621-671)."""
    if isinstance(np_, int):
        np_, n = (np_,), (n,)
    np_, n = tuple(int(x) for x in np_), tuple(int(x) for x in n)
    D = len(n)
    nparts = int(np.prod(np_))
    nglobal = int(np.prod(n))
    owners1d = [_owner_1d(np_[d], n[d]) for d in range(D)]
    out = []
    for rank in range(1, nparts + 1):
        p = _part_coord(rank, np_)
        own_ranges = tuple(local_range(p[d], np_[d], n[d]) for d in range(D))
        if ghost is None:
            _, gids = _box_ids(own_ranges, n)
            l2o = np.full(len(gids), rank, dtype=np.int32)
            out.append(LocalIndices(nglobal, rank, gids, l2o, box=own_ranges, grid=n, parts_per_dir=np_))
            continue
        per = tuple(periodic) if periodic is not None else tuple(False for _ in n)
        loc_ranges = tuple(local_range(p[d], np_[d], n[d], bool(ghost[d]), bool(per[d])) for d in range(D))
        flat, gids = _box_ids(loc_ranges, n, wrap=True)
        wrapped = [np.mod(f - 1, n[d]) + 1 for d, f in enumerate(flat)]
        is_own = np.ones(len(gids), dtype=bool)
        for d in range(D):
            is_own &= (flat[d] >= own_ranges[d][0]) & (flat[d] <= own_ranges[d][1])
        ocoord = [owners1d[d][wrapped[d]] for d in range(D)]
        owner = np.zeros(len(gids), dtype=np.int64)
        stride = 1
        for d in range(D):
            owner += (ocoord[d].astype(np.int64) - 1) * stride
            stride *= np_[d]
        owner = (owner + 1).astype(np.int32)
        owner[is_own] = rank
        out.append(LocalIndices(nglobal, rank, gids, owner, box=own_ranges, grid=n, parts_per_dir=np_, own_mask=is_own))
    return out


def variable_partition(n_own: Sequence[int], n_global: int) -> List[LocalIndices]:
    """1-D variable block partition (src/p_range.jl:705-729), no ghosts."""
    out, start = [], 1
    for rank, no in enumerate(n_own, start=1):
        gids = np.arange(start, start + no, dtype=np.int64)
        out.append(LocalIndices(n_global, rank, gids, np.full(no, rank, dtype=np.int32)))
        start += no
    return out


def global_to_owner_table(partition: List[LocalIndices]) -> np.ndarray:
    """Dense gid(1-based)->owner table built from the own ids of every part."""
    ng = partition[0].n_global
    tab = np.zeros(ng + 1, dtype=np.int32)
    for ind in partition:
        tab[ind.own_to_global] = ind.part
    return tab


def find_owner(partition: List[LocalIndices], gids_per_part: List[np.ndarray]) -> List[np.ndarray]:
    """src/p_range.jl:346-348,1609-1619: owner of each gid; ids < 1 map to owner 0."""
    tab = global_to_owner_table(partition)
    out = []
    for g in gids_per_part:
        g = np.asarray(g, dtype=np.int64)
        o = np.zeros(len(g), dtype=np.int32)
        ok = g >= 1
        o[ok] = tab[g[ok]]
        out.append(o)
    return out


def union_ghost(ind: LocalIndices, gids, owners) -> LocalIndices:
    """src/p_range.jl:205-259: append new ghosts in order of first appearance, skipping
    ids < 1, ids owned by this part, and ids that are already ghosts."""
    gids = np.asarray(gids, dtype=np.int64)
    owners = np.asarray(owners, dtype=np.int32)
    mask = (gids >= 1) & (owners != ind.part)
    cand, cand_o = gids[mask], owners[mask]
    # first appearance order
    _, first = np.unique(cand, return_index=True)
    first.sort()
    cand, cand_o = cand[first], cand_o[first]
    if ind.n_ghost:
        already = np.isin(cand, ind.ghost_to_global)
        cand, cand_o = cand[~already], cand_o[~already]
    if not ind.own_is_prefix() and len(cand):
        raise ValueError("replace_ghost only makes sense for un-permuted local indices (src/p_range.jl:1402)")
    return LocalIndices(
        ind.n_global,
        ind.part,
        np.concatenate([ind.local_to_global, cand]),
        np.concatenate([ind.local_to_owner, cand_o]),
        box=ind.box,
        grid=ind.grid,
        parts_per_dir=ind.parts_per_dir,
        own_mask=np.concatenate([ind.own_mask, np.zeros(len(cand), dtype=bool)]),
    )


# --------------------------------------------------------------------------------------
# ExchangeGraph / exchange  (src/primitives.jl:728-859, 1005-1042)
# --------------------------------------------------------------------------------------


def find_rcv_ids(snd: List[Sequence[int]]) -> List[List[int]]:
    """src/primitives.jl:826-859: transpose of the adjacency; rcv lists sorted ascending."""
    np_ = len(snd)
    rcv = [[] for _ in range(np_)]
    for p in range(np_):
        for q in sorted(set(int(x) for x in snd[p])):
            rcv[q - 1].append(p + 1)
    return rcv


def exchange(snd: List[JaggedArray], graph_snd, graph_rcv) -> List[JaggedArray]:
    """Vector-payload exchange (src/primitives.jl:1020-1042): rcv[r].segment(i) =
    snd[s].segment(j) with s = graph_rcv[r][i], graph_snd[s][j] == r."""
    out = []
    for r in range(len(snd)):
        segs = []
        for s in graph_rcv[r]:
            j = list(graph_snd[s - 1]).index(r + 1)
            segs.append(snd[s - 1].segment(j).copy())
        dtype = snd[0].data.dtype if len(snd) else np.float64
        out.append(jagged_from_lists(segs, dtype))
    return out


# --------------------------------------------------------------------------------------
# Exchange plan  (src/p_range.jl:417-531; src/p_vector.jl:418-468)
# --------------------------------------------------------------------------------------


@dataclass
class AssemblyPlan:
    """Per-part VectorAssemblyCache arrays (src/p_vector.jl:418-426), all 1-based."""

    neighbors_snd: List[np.ndarray]
    neighbors_rcv: List[np.ndarray]
    local_indices_snd: List[JaggedArray]
    local_indices_rcv: List[JaggedArray]


def assembly_neighbors(partition: List[LocalIndices], symmetric: bool = False):
    """src/p_range.jl:436-450: snd = sorted unique owners of ghosts."""
    snd = []
    for ind in partition:
        o = ind.local_to_owner
        snd.append(np.unique(o[o != ind.part]).astype(np.int32))
    rcv = [np.array(s, copy=True) for s in snd] if symmetric else [np.array(r, dtype=np.int32) for r in find_rcv_ids(snd)]
    return snd, rcv


def assembly_local_indices(partition, nbr_snd, nbr_rcv):
    """src/p_range.jl:489-531: send lids grouped by owner in local-id order; the gids are
    exchanged and the receive lids are global_to_local[gid] in the sender's order."""
    lids_snd, gids_snd = [], []
    for ind, ps in zip(partition, nbr_snd):
        segs_l, segs_g = [], []
        for owner in ps:
            l = np.nonzero(ind.local_to_owner == owner)[0]
            segs_l.append((l + 1).astype(np.int32))
            segs_g.append(ind.local_to_global[l])
        lids_snd.append(jagged_from_lists(segs_l, np.int32))
        gids_snd.append(jagged_from_lists(segs_g, np.int64))
    gids_rcv = exchange(gids_snd, nbr_snd, nbr_rcv)
    lids_rcv = []
    for ind, gr in zip(partition, gids_rcv):
        lids_rcv.append(JaggedArray(ind.global_to_local(gr.data).astype(np.int32), gr.ptrs))
    return lids_snd, lids_rcv


def assembly_plan(partition: List[LocalIndices]) -> AssemblyPlan:
    snd, rcv = assembly_neighbors(partition)
    ls, lr = assembly_local_indices(partition, snd, rcv)
    return AssemblyPlan(snd, rcv, ls, lr)


def reverse_plan(plan: AssemblyPlan) -> AssemblyPlan:
    """Base.reverse(::VectorAssemblyCache) src/p_vector.jl:427-437."""
    return AssemblyPlan(plan.neighbors_rcv, plan.neighbors_snd, plan.local_indices_rcv, plan.local_indices_snd)


def assemble_impl(f: Callable, values: List[np.ndarray], plan: AssemblyPlan) -> None:
    """src/p_vector.jl:587-612: pack, exchange, unpack values[lid]=f(values[lid],buf[p])
    in neighbour order."""
    bufs = []
    for v, ls in zip(values, plan.local_indices_snd):
        bufs.append(JaggedArray(v[ls.data - 1].copy(), ls.ptrs))
    rcv = exchange(bufs, plan.neighbors_snd, plan.neighbors_rcv)
    for v, lr, br in zip(values, plan.local_indices_rcv, rcv):
        for p, lid in enumerate(lr.data):
            v[lid - 1] = f(v[lid - 1], br.data[p])


def assemble(values: List[np.ndarray], partition: List[LocalIndices], plan: AssemblyPlan, op=lambda a, b: a + b):
    """assemble!(o,a::PVector) src/p_vector.jl:695-708: combine at owner, then zero ghosts."""
    assemble_impl(op, values, plan)
    for v, ind in zip(values, partition):
        v[ind.ghost_to_local - 1] = 0


def consistent(values: List[np.ndarray], plan: AssemblyPlan):
    """consistent!(a::PVector) src/p_vector.jl:747-755: reversed plan with insert(a,b)=b."""
    assemble_impl(lambda a, b: b, values, reverse_plan(plan))


# --------------------------------------------------------------------------------------
# Local sparse matrices  (src/sparse_utils.jl)
# --------------------------------------------------------------------------------------


@dataclass
class CSR:
    """SparseMatrixCSR{1} layout (1-based rowptr/colval)."""

    m: int
    n: int
    rowptr: np.ndarray
    colval: np.ndarray
    nzval: np.ndarray

    @property
    def nnz(self):
        return len(self.nzval)

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix((self.nzval, self.colval - 1, self.rowptr - 1), shape=(self.m, self.n))


def sparse_matrix_csr(I, J, V, m, n, skip=True, index_dtype=np.int32) -> CSR:
    """sparse_matrix/compresscoo for SparseMatrixCSR{1} (src/sparse_utils.jl:313-350,398-405).
    Entries with i<1 or j<1 are replaced by a stored (1,1,0) when skip (FilteredCooVector,
    :370-390); columns sorted within rows; duplicates combined with + in input order
    (sparsecsr of SparseMatricesCSR -> SparseArrays.sparse, stable)."""
    I = np.asarray(I, dtype=np.int64).copy()
    J = np.asarray(J, dtype=np.int64).copy()
    V = np.asarray(V, dtype=np.float64).copy()
    if skip:
        if m * n == 0:
            I, J, V = I[:0], J[:0], V[:0]
        else:
            bad = (I < 1) | (J < 1)
            I[bad], J[bad], V[bad] = 1, 1, 0.0
    order = np.lexsort((J, I))  # stable: primary I, secondary J, ties in input order
    I, J, V = I[order], J[order], V[order]
    if len(I):
        new = np.ones(len(I), dtype=bool)
        new[1:] = (I[1:] != I[:-1]) | (J[1:] != J[:-1])
        pos = np.cumsum(new) - 1
        nz = np.zeros(int(pos[-1]) + 1, dtype=np.float64)
        np.add.at(nz, pos, V)  # sequential, in order
        Iu, Ju = I[new], J[new]
    else:
        nz, Iu, Ju = V, I, J
    counts = np.bincount(Iu - 1, minlength=m) if len(Iu) else np.zeros(m, dtype=np.int64)
    rowptr = np.ones(m + 1, dtype=np.int64)
    rowptr[1:] = 1 + np.cumsum(counts)
    return CSR(m, n, rowptr.astype(index_dtype if rowptr[-1] < 2**31 else np.int64), Ju.astype(index_dtype), nz)


def spmv_csr_py(A: CSR, x: np.ndarray) -> np.ndarray:
    """spmv_csr! (src/sparse_utils.jl:649-669): per row sequential bi += aij*xj (no FMA)."""
    b = np.zeros(A.m, dtype=np.float64)
    rp, cv, nz = A.rowptr, A.colval, A.nzval
    for row in range(A.m):
        bi = 0.0
        for p in range(rp[row] - 1, rp[row + 1] - 1):
            bi += float(nz[p]) * float(x[cv[p] - 1])
        b[row] = bi
    return b


def spmv_csc_py(m, colptr, rowval, nzval, x) -> np.ndarray:
    """spmv_csc! (src/sparse_utils.jl:671-690): fill!(b,0) then column scatter."""
    b = np.zeros(m, dtype=np.float64)
    for col in range(len(x)):
        xj = float(x[col])
        for p in range(colptr[col] - 1, colptr[col + 1] - 1):
            b[rowval[p] - 1] += float(nzval[p]) * xj
    return b


def spmv_csr(A: CSR, x: np.ndarray, y0: Optional[np.ndarray] = None) -> np.ndarray:
    """Fast sequential-order spmv: C oracle if built, else pure Python.  When y0 is given
    computes y0 + A*x with the terms added to y0 one by one (the 5-arg mul!(b,A,x,1,1) of
    SparseMatricesCSR 0.6 used at src/p_sparse_matrix.jl:2088,2101)."""
    from . import c_oracle

    if c_oracle.available():
        return c_oracle.spmv_csr(A, x, y0)
    b = spmv_csr_py(A, x) if y0 is None else None
    if y0 is not None:
        b = y0.copy()
        for row in range(A.m):
            for p in range(A.rowptr[row] - 1, A.rowptr[row + 1] - 1):
                b[row] += float(A.nzval[p]) * float(x[A.colval[p] - 1])
    return b


# --------------------------------------------------------------------------------------
# PSparseMatrix restatement  (src/p_sparse_matrix.jl)
# --------------------------------------------------------------------------------------


@dataclass
class PSparse:
    row_partition: List[LocalIndices]
    col_partition: List[LocalIndices]
    # unsplit local matrices n_own_rows x n_local_cols?  No: reference local matrix is
    # n_local_rows x n_local_cols (src/p_sparse_matrix.jl:971-991).  For assembled matrices the
    # ghost rows are empty, so we keep rows in *local row id* order as well.
    local: List[CSR]
    # split blocks in own/ghost numbering (src/p_sparse_matrix.jl:588-593, 823-899)
    own_own: List[CSR]
    own_ghost: List[CSR]
    assembled: bool = True


def split_format_locally(A: CSR, rows: LocalIndices, cols: LocalIndices) -> Tuple[CSR, CSR]:
    """split_format_locally (src/p_sparse_matrix.jl:823-899), own-row blocks only (the
    ghost-row blocks of an assembled matrix are empty, :1704-1705).  Entry order inside
    each block follows the row-major traversal of the CSR input; column ids are converted
    to own / ghost ids."""
    n_own_r, n_own_c, n_gh_c = rows.n_own, cols.n_own, cols.n_ghost
    l2own_r = np.zeros(rows.n_local + 1, dtype=np.int64)
    l2own_r[rows.own_to_local] = np.arange(1, n_own_r + 1)
    l2own_c = np.zeros(cols.n_local + 1, dtype=np.int64)
    l2own_c[cols.own_to_local] = np.arange(1, n_own_c + 1)
    l2gh_c = np.zeros(cols.n_local + 1, dtype=np.int64)
    l2gh_c[cols.ghost_to_local] = np.arange(1, n_gh_c + 1)
    rowid = np.repeat(np.arange(1, A.m + 1), np.diff(A.rowptr.astype(np.int64)))
    ri = l2own_r[rowid]
    keep = ri > 0
    co, cg = l2own_c[A.colval], l2gh_c[A.colval]
    moo = keep & (co > 0)
    mog = keep & (cg > 0)
    oo = sparse_matrix_csr(ri[moo], co[moo], A.nzval[moo], n_own_r, n_own_c, skip=False)
    og = sparse_matrix_csr(ri[mog], cg[mog], A.nzval[mog], n_own_r, n_gh_c, skip=False)
    return oo, og


def psparse(I, J, V, row_partition, col_partition, assembled: bool = False, local_format: str = "csc") -> PSparse:
    """psparse (src/p_sparse_matrix.jl:1150-1286) restated at the semantic level.

    assembled=False: the reference's default, restated step by step in psparse_disassembled.
    assembled=True: every triplet already sits on its row owner (:1249-1270)."""
    nparts = len(row_partition)
    I = [np.asarray(i, dtype=np.int64) for i in I]
    J = [np.asarray(j, dtype=np.int64) for j in J]
    V = [np.asarray(v, dtype=np.float64) for v in V]
    if not assembled:
        return psparse_disassembled(I, J, V, row_partition, col_partition, local_format)
    Jown = find_owner(col_partition, J)
    cols = [union_ghost(c, j, o) for c, j, o in zip(col_partition, J, Jown)]
    rows = row_partition
    local, oo, og = [], [], []
    for p in range(nparts):
        li = rows[p].global_to_local(I[p]).astype(np.int64)
        lj = cols[p].global_to_local(J[p]).astype(np.int64)
        li[I[p] < 1] = 0
        lj[J[p] < 1] = 0
        A = sparse_matrix_csr(li, lj, V[p], rows[p].n_local, cols[p].n_local, skip=True)
        local.append(A)
        a, b = split_format_locally(A, rows[p], cols[p])
        oo.append(a); og.append(b)
    return PSparse(rows, cols, local, oo, og, True)


def _compress_coo(li, lj, V, m, n, fmt):
    """compresscoo (src/sparse_utils.jl:313-350) seen as a list of stored entries: unique (i,j) in STORAGE order of
    the local matrix type -- "csr": row-major (SparseMatrixCSR), "csc": column-major (SparseMatrixCSC, the default
    of psparse) -- values of duplicates added in input order; triplets with an id < 1 become a stored (1,1,0.0)
    (FilteredCooVector, :370-390)."""
    li = np.asarray(li, dtype=np.int64).copy()
    lj = np.asarray(lj, dtype=np.int64).copy()
    V = np.asarray(V, dtype=np.float64).copy()
    if m * n == 0:
        li, lj, V = li[:0], lj[:0], V[:0]
    bad = (li < 1) | (lj < 1)
    li[bad], lj[bad], V[bad] = 1, 1, 0.0
    order = np.lexsort((lj, li)) if fmt == "csr" else np.lexsort((li, lj))  # stable: ties stay in input order
    li, lj, V = li[order], lj[order], V[order]
    if len(li) == 0:
        return li, lj, V
    new = np.ones(len(li), dtype=bool)
    new[1:] = (li[1:] != li[:-1]) | (lj[1:] != lj[:-1])
    nz = np.zeros(int(new.sum()))
    np.add.at(nz, np.cumsum(new) - 1, V)  # sequential, in order
    return li[new], lj[new], nz


def psparse_disassembled(I, J, V, row_partition, col_partition, local_format: str = "csc") -> PSparse:
    """psparse(I,J,V,rows,cols) with its defaults (disassembled input, assemble=true, split_format=true), step by
    step as the reference does it, because both the association of the sums and the numbering of the ghost columns
    (hence the order of the terms of every row of the ghost block) follow from these steps:
      1. src/p_sparse_matrix.jl:1186-1201  rows_sa/cols_sa = union_ghost(rows/cols, I/J, owners), local ids,
         sub-assembled local matrix = compress(I,J,V) per part (duplicates of ONE part combined, in input order);
      2. :1207-1209 (split_format, :823-899)  four blocks, entries kept in storage order;
      3. :1600-1645 (setup_cache_snd)  entries in ghost rows, ghost_own block first then ghost_ghost, each in storage
         order, bucketed by the owner of the row (neighbours = sorted ghost-row owners);
      4. :1651-1684 (setup_own_triplets)  own_own list = findnz(own_own) ++ received entries with an own column,
         own_ghost list = findnz(own_ghost) ++ the other received entries, received in neighbour (ascending part) order;
      5. :1739 cols_fa = union_ghost(cols without ghosts, columns of the own_ghost list in that order);
      6. :1700-1703 compresscoo of both lists: value = own sum, then + each sender's sum, in neighbour order.
    local_format = storage of the local matrices: "csc" (SparseMatrixCSC, default) or "csr" (SparseMatrixCSR{1})."""
    nparts = len(row_partition)
    I = [np.asarray(i, dtype=np.int64) for i in I]
    J = [np.asarray(j, dtype=np.int64) for j in J]
    V = [np.asarray(v, dtype=np.float64) for v in V]
    rows_sa = [union_ghost(r, i, o) for r, i, o in zip(row_partition, I, find_owner(row_partition, I))]
    cols_sa = [union_ghost(c, j, o) for c, j, o in zip(col_partition, J, find_owner(col_partition, J))]
    for r, c in zip(rows_sa, cols_sa):
        assert r.own_is_prefix() and c.own_is_prefix(), "restated for own-first local orders"
    nbr_snd, nbr_rcv = assembly_neighbors(rows_sa)
    own_lists, outbox = [], []
    for p in range(nparts):
        r, c = rows_sa[p], cols_sa[p]
        li = r.global_to_local(I[p]).astype(np.int64)
        lj = c.global_to_local(J[p]).astype(np.int64)
        li[I[p] < 1] = 0
        lj[J[p] < 1] = 0
        ei, ej, ev = _compress_coo(li, lj, V[p], r.n_local, c.n_local, local_format)  # storage order
        gi, gj = r.local_to_global[ei - 1], c.local_to_global[ej - 1]
        row_own, col_own = ei <= r.n_own, ej <= c.n_own
        own_lists.append({"oo": (gi[row_own & col_own], gj[row_own & col_own], ev[row_own & col_own]),
                          "og": (gi[row_own & ~col_own], gj[row_own & ~col_own], ev[row_own & ~col_own])})
        # ghost rows: ghost_own block entries first, then ghost_ghost
        sel = np.concatenate([np.nonzero(~row_own & col_own)[0], np.nonzero(~row_own & ~col_own)[0]])
        owner = r.local_to_owner[ei[sel] - 1]
        outbox.append({int(q): (gi[sel][owner == q], gj[sel][owner == q], ev[sel][owner == q]) for q in nbr_snd[p]})
    cols_fa, oo, og, local = [], [], [], []
    for p in range(nparts):
        r, c = row_partition[p], col_partition[p]
        rcv = [outbox[int(s) - 1][p + 1] for s in nbr_rcv[p]]
        ri = np.concatenate([x[0] for x in rcv]) if rcv else np.zeros(0, np.int64)
        rj = np.concatenate([x[1] for x in rcv]) if rcv else np.zeros(0, np.int64)
        rv = np.concatenate([x[2] for x in rcv]) if rcv else np.zeros(0)
        jown = c.global_to_local(rj) > 0 if len(rj) else np.zeros(0, dtype=bool)
        a, b = own_lists[p]["oo"], own_lists[p]["og"]
        oo_i, oo_j, oo_v = np.concatenate([a[0], ri[jown]]), np.concatenate([a[1], rj[jown]]), np.concatenate([a[2], rv[jown]])
        og_i, og_j, og_v = np.concatenate([b[0], ri[~jown]]), np.concatenate([b[1], rj[~jown]]), np.concatenate([b[2], rv[~jown]])
        c_own = LocalIndices(c.n_global, c.part, c.own_to_global, c.local_to_owner[c.own_to_local - 1],
                             box=c.box, grid=c.grid, parts_per_dir=c.parts_per_dir)  # remove_ghost
        cfa = union_ghost(c_own, og_j, find_owner(col_partition, [og_j])[0])
        cols_fa.append(cfa)
        n_own_r = r.n_own
        row_own_id = lambda g: r.global_to_local(g).astype(np.int64)  # own-first: own id == local id
        ei, ej, ev = _compress_coo(row_own_id(oo_i), c_own.global_to_local(oo_j).astype(np.int64), oo_v, n_own_r, c_own.n_own, "csr")
        oo.append(_entries_to_csr(ei, ej, ev, n_own_r, c_own.n_own))
        gid = cfa.global_to_local(og_j).astype(np.int64) - cfa.n_own if len(og_j) else np.zeros(0, np.int64)
        ei, ej, ev = _compress_coo(row_own_id(og_i), gid, og_v, n_own_r, cfa.n_ghost, "csr")
        og.append(_entries_to_csr(ei, ej, ev, n_own_r, cfa.n_ghost))
        # unsplit view (own rows x local cols) for the callers that want one matrix
        rows_of = lambda M: np.repeat(np.arange(1, M.m + 1), np.diff(M.rowptr.astype(np.int64)))
        li = np.concatenate([rows_of(oo[-1]), rows_of(og[-1])])
        lj = np.concatenate([oo[-1].colval.astype(np.int64), og[-1].colval.astype(np.int64) + cfa.n_own])
        lv = np.concatenate([oo[-1].nzval, og[-1].nzval])
        local.append(sparse_matrix_csr(li, lj, lv, n_own_r, cfa.n_local, skip=False))
    return PSparse(list(row_partition), cols_fa, local, oo, og, True)  # rows_fa = rows (:1737)


def psparse_subassembled(I, J, V, row_partition, col_partition, local_format: str = "csc") -> PSparse:
    """psparse(I,J,V,rows,cols; assemble=false) (src/p_sparse_matrix.jl:1186-1222): steps 1-2 of psparse_disassembled only.
    `local` holds ALL local rows (own rows first, then the ghost rows) over rows_sa x cols_sa."""
    I = [np.asarray(i, dtype=np.int64) for i in I]
    J = [np.asarray(j, dtype=np.int64) for j in J]
    V = [np.asarray(v, dtype=np.float64) for v in V]
    rows_sa = [union_ghost(r, i, o) for r, i, o in zip(row_partition, I, find_owner(row_partition, I))]
    cols_sa = [union_ghost(c, j, o) for c, j, o in zip(col_partition, J, find_owner(col_partition, J))]
    local = []
    for p, (r, c) in enumerate(zip(rows_sa, cols_sa)):
        assert r.own_is_prefix() and c.own_is_prefix(), "restated for own-first local orders"
        li = r.global_to_local(I[p]).astype(np.int64)
        lj = c.global_to_local(J[p]).astype(np.int64)
        li[I[p] < 1] = 0
        lj[J[p] < 1] = 0
        ei, ej, ev = _compress_coo(li, lj, V[p], r.n_local, c.n_local, local_format)
        local.append(sparse_matrix_csr(ei, ej, ev, r.n_local, c.n_local, skip=False))
    return PSparse(rows_sa, cols_sa, local, [], [], False)


def pmul_subassembled(A: PSparse, b_vals: List[np.ndarray], c_vals: List[np.ndarray], alpha: float = 1.0, beta: float = 0.0):
    """mul!(c,A,b,alpha,beta) for !A.assembled (src/p_sparse_matrix.jl:2105-2142): consistent!(b); own and ghost rows of c get
    beta*c + alpha*(own block of the row, then its ghost block); assemble!(c).  With alpha = 1, beta = 0 the row sums are the
    sequential sums of spmv_csr! over the row sorted by local column (own columns first)."""
    consistent(b_vals, assembly_plan(A.col_partition))
    for p in range(len(b_vals)):
        y = spmv_csr(A.local[p], b_vals[p])
        c_vals[p][:] = y if (alpha == 1.0 and beta == 0.0) else alpha * y + (beta * c_vals[p] if beta != 0.0 else 0.0)
    assemble(c_vals, A.row_partition, assembly_plan(A.row_partition))


def _entries_to_csr(ei, ej, ev, m, n) -> CSR:
    """row-major unique entries -> SparseMatrixCSR{1} arrays."""
    counts = np.bincount(ei - 1, minlength=m) if len(ei) else np.zeros(m, dtype=np.int64)
    rowptr = np.ones(m + 1, dtype=np.int64)
    rowptr[1:] = 1 + np.cumsum(counts)
    return CSR(m, n, rowptr.astype(np.int32), np.asarray(ej, dtype=np.int32), np.asarray(ev, dtype=np.float64))


def pvector_disassembled(I, V, row_partition) -> List[np.ndarray]:
    """pvector(I,V,rows) with its defaults (src/p_vector.jl:887-926 + assemble :1331-1347): rows_sa =
    union_ghost(rows, I), dense_vector per part (a[i] += v in input order, ids < 1 skipped, :853-863), assemble!
    (owner += each neighbour's contribution, in neighbour order), result on `rows` (own values only)."""
    I = [np.asarray(i, dtype=np.int64) for i in I]
    rows_sa = [union_ghost(r, i, o) for r, i, o in zip(row_partition, I, find_owner(row_partition, I))]
    vals = []
    for r, i, v in zip(rows_sa, I, V):
        a = np.zeros(r.n_local)
        ok = i >= 1
        np.add.at(a, r.global_to_local(i[ok]).astype(np.int64) - 1, np.asarray(v, dtype=np.float64)[ok])
        vals.append(a)
    assemble(vals, rows_sa, assembly_plan(rows_sa))
    return [own_values(a, r).copy() for a, r in zip(vals, rows_sa)]


def own_values(v: np.ndarray, ind: LocalIndices) -> np.ndarray:
    return v[ind.own_to_local - 1]


def ghost_values(v: np.ndarray, ind: LocalIndices) -> np.ndarray:
    return v[ind.ghost_to_local - 1]


def pmul(A: PSparse, b_vals: List[np.ndarray], plan_cols: AssemblyPlan, c_vals: List[np.ndarray]):
    """mul!(c,A,b) split format (src/p_sparse_matrix.jl:2090-2103): consistent!(b);
    c_own = A_oo*b_own ; c_own += A_oh*b_ghost ; ghost entries of c untouched."""
    consistent(b_vals, plan_cols)
    for p in range(len(b_vals)):
        rows, cols = A.row_partition[p], A.col_partition[p]
        bo = own_values(b_vals[p], cols)
        bg = ghost_values(b_vals[p], cols)
        co = spmv_csr(A.own_own[p], bo)
        co = spmv_csr(A.own_ghost[p], bg, y0=co)
        c_vals[p][rows.own_to_local - 1] = co


def mul_no_lat(A: PSparse, b_vals, plan_cols, c_vals):
    """HPCG mul_no_lat! (HPCG/src/hpcg_utils.jl:6-17): consistent!(b)|>wait; one spmv! of the
    unsplit local CSR against the local x (own rows only; requires own rows first)."""
    consistent(b_vals, plan_cols)
    for p in range(len(b_vals)):
        rows = A.row_partition[p]
        y = spmv_csr(A.local[p], b_vals[p])
        c_vals[p][rows.own_to_local - 1] = y[rows.own_to_local - 1]


def pdot(a_vals, b_vals, partition) -> float:
    """dot(a::PVector,b::PVector) src/p_vector.jl:1189-1192: per-part dot(own,own), then sum
    over parts in part order (DebugArray reduce, src/primitives.jl:693-698)."""
    s = 0.0
    for a, b, ind in zip(a_vals, b_vals, partition):
        s += float(np.dot(own_values(a, ind), own_values(b, ind)))
    return s


def pnorm(a_vals, partition) -> float:
    """norm(a::PVector,2) src/p_vector.jl:1201-1206: (sum_parts norm(own)^2)^(1/2)."""
    s = 0.0
    for a, ind in zip(a_vals, partition):
        s += float(np.linalg.norm(own_values(a, ind))) ** 2
    return math.sqrt(s)


def pvector_from_global(xg: np.ndarray, partition: List[LocalIndices], ghosts: bool = True) -> List[np.ndarray]:
    """Local arrays whose own (and optionally ghost) entries are xg[gid-1]."""
    out = []
    for ind in partition:
        v = xg[ind.local_to_global - 1].astype(np.float64).copy()
        if not ghosts:
            v[ind.ghost_to_local - 1] = 0.0
        out.append(v)
    return out


def collect(vals: List[np.ndarray], partition: List[LocalIndices]) -> np.ndarray:
    out = np.zeros(partition[0].n_global, dtype=np.float64)
    for v, ind in zip(vals, partition):
        out[ind.own_to_global - 1] = own_values(v, ind)
    return out


# --------------------------------------------------------------------------------------
# Input generators
# --------------------------------------------------------------------------------------


def laplacian_fdm(nodes_per_dir: Sequence[int], parts_per_dir: Sequence[int]):
    """gallery laplacian_fdm (src/gallery.jl:12-86): COO per part in own-node order; per node
    the diagonal (alpha*2D) first, then for d=1..D, i in (-1,+1) the neighbour (-alpha) when
    inside the grid.  alpha = prod(n_i+1).  Returns I,J,V (lists per part), row/col partition."""
    n = tuple(int(x) for x in nodes_per_dir)
    D = len(n)
    alpha = float(np.prod([i + 1 for i in n]))
    part = uniform_partition(parts_per_dir, n)
    Is, Js, Vs = [], [], []
    for ind in part:
        flat, gids = _box_ids(ind.box, n)
        cols = [gids]
        vals = [np.full(len(gids), alpha * 2 * D)]
        valid = [np.ones(len(gids), dtype=bool)]
        for d in range(D):
            for i in (-1, 1):
                nb = [f.copy() for f in flat]
                nb[d] = nb[d] + i
                inside = (nb[d] >= 1) & (nb[d] <= n[d])
                nb[d] = np.clip(nb[d], 1, n[d])
                cols.append(_cartesian_linear(nb, n))
                vals.append(np.full(len(gids), -alpha))
                valid.append(inside)
        C = np.stack(cols, axis=1)
        Vv = np.stack(vals, axis=1)
        M = np.stack(valid, axis=1)
        R = np.repeat(gids[:, None], C.shape[1], axis=1)
        Is.append(R[M]); Js.append(C[M]); Vs.append(Vv[M])
    return Is, Js, Vs, part, part


def hpcg_build_matrix(nx, ny, nz, gnx, gny, gnz, gix0, giy0, giz0):
    """HPCG build_matrix (HPCG/src/sparse_matrix.jl:27-80): 27-pt, diag 26, off -1,
    b = 27 - nnz_row; COO in (iz,iy,ix) row order with (sz,sy,sx) neighbour order."""
    iz, iy, ix = np.meshgrid(np.arange(1, nz + 1), np.arange(1, ny + 1), np.arange(1, nx + 1), indexing="ij")
    ix, iy, iz = ix.reshape(-1).astype(np.int64), iy.reshape(-1).astype(np.int64), iz.reshape(-1).astype(np.int64)
    gix, giy, giz = gix0 + ix - 1, giy0 + iy - 1, giz0 + iz - 1
    grow = (giz - 1) * gnx * gny + (giy - 1) * gnx + (gix - 1) + 1
    cols, valid = [], []
    for sz in (-1, 0, 1):
        for sy in (-1, 0, 1):
            for sx in (-1, 0, 1):
                ok = (
                    (giz + sz > 0) & (giz + sz < gnz + 1) & (giy + sy > 0) & (giy + sy < gny + 1)
                    & (gix + sx > 0) & (gix + sx < gnx + 1)
                )
                cols.append(grow + sz * gnx * gny + sy * gnx + sx)
                valid.append(ok)
    C = np.stack(cols, axis=1)
    M = np.stack(valid, axis=1)
    R = np.repeat(grow[:, None], 27, axis=1)
    Vv = np.where(C == R, 26.0, -1.0)
    b = 27.0 - M.sum(axis=1).astype(np.float64)
    return R[M], C[M], Vv[M], b, grow


def hpcg_build_p_matrix(nx, ny, nz, npx, npy, npz):
    """build_p_matrix (HPCG/src/sparse_matrix.jl:105-122): unsplit CSR{1,Float64,Int32},
    assembled=true; b on the column partition (own values).  Returns (A:PSparse, b_vals)."""
    gnx, gny, gnz = nx * npx, ny * npy, nz * npz
    rows = uniform_partition((npx, npy, npz), (gnx, gny, gnz))
    Is, Js, Vs, bs = [], [], [], []
    for ind in rows:
        g0 = tuple(r[0] for r in ind.box)
        I, J, V, b, _ = hpcg_build_matrix(nx, ny, nz, gnx, gny, gnz, *g0)
        Is.append(I); Js.append(J); Vs.append(V); bs.append(b)
    A = psparse(Is, Js, Vs, rows, rows, assembled=True)
    b_vals = []
    for ind, b in zip(A.col_partition, bs):
        v = np.zeros(ind.n_local)
        v[ind.own_to_local - 1] = b
        b_vals.append(v)
    return A, b_vals


# --------------------------------------------------------------------------------------
# CG  (HPCG/src/ref_cg.jl:40-134), Pl = Identity
# --------------------------------------------------------------------------------------


def ref_cg(A: PSparse, b_vals, x_vals, maxiter: int, tolerance: float = 0.0, mul=None):
    """ref_cg! with Pl=Identity.  Broadcast updates write own AND ghost entries
    (src/p_vector.jl:1271-1276).  Returns (x_vals, residual0, residual, iters, history)."""
    part = A.col_partition
    plan = assembly_plan(part)
    mul = mul or mul_no_lat
    u = [np.zeros_like(x) for x in x_vals]
    r = [b.copy() for b in b_vals]
    c = [np.zeros_like(x) for x in x_vals]
    pmul(A, x_vals, plan, c)  # cg_iterator! uses the generic mul! (ref_cg.jl:87)
    for p in range(len(r)):
        r[p] -= c[p]
    residual0 = residual = pnorm(r, part)
    rho = 1.0
    hist = [residual]
    it = 0
    while not (it >= maxiter or (residual / residual0 <= tolerance if residual0 != 0 else True)):
        for p in range(len(r)):
            c[p][:] = r[p]  # ldiv!(c, Identity, r)
        rho_prev = rho
        rho = pdot(c, r, part)
        beta = rho / rho_prev
        for p in range(len(r)):
            u[p][:] = c[p] + beta * u[p]
        mul(A, u, plan, c)
        uc = pdot(u, c, part)
        alpha = rho / uc
        for p in range(len(r)):
            x_vals[p] += alpha * u[p]
            r[p] -= alpha * c[p]
        residual = pnorm(r, part)
        hist.append(residual)
        it += 1
    return x_vals, residual0, residual, it, hist


# --------------------------------------------------------------------------------------
# Deterministic pseudo-random vector by global id (shared with the CUDA fill kernel)
# --------------------------------------------------------------------------------------


def hash_uniform(gids_1based: np.ndarray, seed: int) -> np.ndarray:
    """splitmix64(gid0 + seed*0x9E3779B97F4A7C15) -> double in [-1,1): (top 53 bits)*2^-52 - 1.
    Mirrors pa_fill_hash in csrc/ (kept bit-identical; see tests)."""
    with np.errstate(over="ignore"):
        z = (np.asarray(gids_1based, dtype=np.uint64) - np.uint64(1)) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -52) - 1.0


def pmul_transpose(A: PSparse, b_vals: List[np.ndarray], c_vals: List[np.ndarray], alpha: float = 1.0, beta: float = 0.0):
    """mul!(c, transpose(A), b, alpha, beta) (src/p_sparse_matrix.jl:2144-2162): ghost entries of c receive
    alpha*A_oh' b_own, assemble!(c) adds them at the owners after c_own = beta*c_own + alpha*A_oo' b_own.
    Column sums run over ascending row index (transposed CSR with sorted rows)."""
    import scipy.sparse as sp

    plan = assembly_plan(A.col_partition)
    for p in range(len(b_vals)):
        rows, cols = A.row_partition[p], A.col_partition[p]
        L = A.local[p]
        At = sp.csr_matrix((L.nzval, L.colval - 1, L.rowptr.astype(np.int64) - 1), shape=(L.m, L.n)).T.tocsr()
        At.sort_indices()
        T = CSR(At.shape[0], At.shape[1], At.indptr.astype(np.int64) + 1, At.indices.astype(np.int32) + 1, At.data)
        bo = np.zeros(L.m)
        bo[rows.own_to_local - 1] = own_values(b_vals[p], rows)
        acc = spmv_csr(T, bo)  # length n_local cols, sequential ascending-row order
        c = c_vals[p]
        gl, ol = cols.ghost_to_local - 1, cols.own_to_local - 1
        c[gl] = alpha * acc[gl]
        c[ol] = alpha * acc[ol] + beta * c[ol]
    assemble(c_vals, A.col_partition, plan)
