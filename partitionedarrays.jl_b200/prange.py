"""Host-side index partitions and exchange plans (setup-time metadata; numpy).

Mirrors the reference's PRange layer (all ids 1-based like Julia):
  local_range            src/p_range.jl:806-818
  uniform_partition      src/p_range.jl:585-671   (block_with_constant_size, with/without ghost layer)
  variable_partition     src/p_range.jl:705-729
  find_owner             src/p_range.jl:346-348, 1502-1513, 1609-1619
  union_ghost            src/p_range.jl:205-259
  assembly_neighbors     src/p_range.jl:417-450
  assembly_local_indices src/p_range.jl:466-531
The data-moving work (consistent!/assemble!/mul!) happens on the GPU; this file only builds the
index arrays the C ABI ingests (pa_plan_set_part)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


def local_range(p: int, np_: int, n: int, ghost: bool = False, periodic: bool = False) -> Tuple[int, int]:
    """Inclusive 1-based range of part p; the remainder goes to the last parts."""
    base, rem = divmod(n, np_)
    length, offset = base, base * (p - 1)
    if rem >= np_ - p + 1:
        length += 1
        offset += p - (np_ - rem) - 1
    lo, hi = 1 + offset - int(ghost), length + offset + int(ghost)
    return (lo, hi) if periodic else (max(1, lo), min(n, hi))


def _lin(coords: Sequence[np.ndarray], dims: Sequence[int]) -> np.ndarray:
    """Column-major linear id (1-based in, 1-based out)."""
    out = np.zeros_like(coords[0], dtype=np.int64)
    stride = 1
    for c, n in zip(coords, dims):
        out += (c.astype(np.int64) - 1) * stride
        stride *= int(n)
    return out + 1


def _box_points(ranges, order_first_fastest=True):
    axes = [np.arange(a, b + 1, dtype=np.int64) for a, b in ranges]
    grids = np.meshgrid(*axes[::-1], indexing="ij")[::-1]
    return [g.reshape(-1) for g in grids]


@dataclass
class BlockInfo:
    """Cartesian block partition description (used for owner formulas and the stencil fast path)."""

    grid: Tuple[int, ...]
    parts_per_dir: Tuple[int, ...]
    box: Tuple[Tuple[int, int], ...]  # own box of this part, 1-based inclusive

    def starts(self, d: int) -> np.ndarray:
        return np.array([local_range(p, self.parts_per_dir[d], self.grid[d])[0] for p in range(1, self.parts_per_dir[d] + 1)])

    def owner_of(self, gids: np.ndarray) -> np.ndarray:
        """BlockPartitionGlobalToOwner (src/p_range.jl:1502-1513): owner part id of 1-based gids (0 for ids<1)."""
        g = np.asarray(gids, dtype=np.int64)
        ok = g >= 1
        r = np.where(ok, g, 1) - 1
        owner = np.zeros(len(g), dtype=np.int64)
        stride = 1
        for d, n in enumerate(self.grid):
            c = r % n + 1
            r = r // n
            pc = np.searchsorted(self.starts(d), c, side="right") - 1
            owner += pc * stride
            stride *= self.parts_per_dir[d]
        return np.where(ok, owner + 1, 0).astype(np.int32)

    def own_index_of(self, gids: np.ndarray, box=None) -> np.ndarray:
        """1-based own id (column-major position inside `box`, default this part's) of gids inside it."""
        box = box or self.box
        r = np.asarray(gids, dtype=np.int64) - 1
        out = np.zeros(len(r), dtype=np.int64)
        stride = 1
        for d, n in enumerate(self.grid):
            c = r % n + 1
            r = r // n
            out += (c - box[d][0]) * stride
            stride *= box[d][1] - box[d][0] + 1
        return out + 1


class LocalIndices:
    """AbstractLocalIndices of one part (src/p_range.jl:32-160) in array form.

    Own and ghost ids are stored separately; for Cartesian block partitions the own ids are implicit
    (column-major positions in the own box) so a 512^3 part never materialises a 1 GB id list."""

    def __init__(self, n_global, part, local_to_global=None, local_to_owner=None, block=None, *, n_own=None,
                 ghost_to_global=None, ghost_to_owner=None, is_own=None):
        self.n_global, self.part, self.block = int(n_global), int(part), block
        self._g2l = None
        self._own_to_global = None
        if local_to_global is not None:
            l2g = np.asarray(local_to_global, dtype=np.int64)
            l2o = np.asarray(local_to_owner, dtype=np.int32)
            # own vs ghost is positional when the caller knows it (block_with_constant_size, src/p_range.jl:648-664): the
            # wrapped ghosts of a periodic direction with ONE part are owned by the part itself and are still ghosts
            own = (l2o == self.part) if is_own is None else np.asarray(is_own, dtype=bool)
            self.n_local, self.n_own = len(l2g), int(np.count_nonzero(own))
            self.own_is_prefix = bool(np.all(own[: self.n_own]))
            self._own_to_global = l2g[own]
            self.ghost_to_global, self.ghost_to_owner = l2g[~own], l2o[~own]
            self._own_to_local = None if self.own_is_prefix else (np.nonzero(own)[0] + 1).astype(np.int32)
            self._ghost_to_local = None if self.own_is_prefix else (np.nonzero(~own)[0] + 1).astype(np.int32)
        else:  # implicit own block + explicit ghosts (own-first layout)
            assert block is not None and n_own is not None
            self.n_own = int(n_own)
            self.ghost_to_global = np.asarray(ghost_to_global if ghost_to_global is not None else [], dtype=np.int64)
            self.ghost_to_owner = np.asarray(ghost_to_owner if ghost_to_owner is not None else [], dtype=np.int32)
            self.n_local = self.n_own + len(self.ghost_to_global)
            self.own_is_prefix = True
            self._own_to_local = self._ghost_to_local = None
        self.n_ghost = self.n_local - self.n_own

    @property
    def own_to_local(self):
        return np.arange(1, self.n_own + 1, dtype=np.int32) if self._own_to_local is None else self._own_to_local

    @property
    def ghost_to_local(self):
        return np.arange(self.n_own + 1, self.n_local + 1, dtype=np.int32) if self._ghost_to_local is None else self._ghost_to_local

    @property
    def own_to_global(self):
        if self._own_to_global is None:
            self._own_to_global = _lin(_box_points(self.block.box), self.block.grid)
        return self._own_to_global

    @property
    def local_to_global(self):
        out = np.zeros(self.n_local, dtype=np.int64)
        out[self.own_to_local - 1] = self.own_to_global
        out[self.ghost_to_local - 1] = self.ghost_to_global
        return out

    @property
    def local_to_owner(self):
        out = np.full(self.n_local, self.part, dtype=np.int32)
        out[self.ghost_to_local - 1] = self.ghost_to_owner
        return out

    def global_to_local(self, gids) -> np.ndarray:
        """0 where the gid is not a local id."""
        gids = np.atleast_1d(np.asarray(gids, dtype=np.int64))
        if self.block is not None and self.own_is_prefix:
            # own ids by formula, ghosts through a sorted table
            out = np.zeros(len(gids), dtype=np.int64)
            inside = gids >= 1
            r = np.where(gids >= 1, gids, 1) - 1
            for d, n in enumerate(self.block.grid):
                c = r % n + 1
                r = r // n
                inside &= (c >= self.block.box[d][0]) & (c <= self.block.box[d][1])
            if inside.any():
                out[inside] = self.block.own_index_of(gids[inside])
            rest = ~inside & (gids >= 1)
            if rest.any() and self.n_ghost:
                gg = self.ghost_to_global
                order = np.argsort(gg, kind="stable")
                sg = gg[order]
                pos = np.clip(np.searchsorted(sg, gids[rest]), 0, len(gg) - 1)
                hit = sg[pos] == gids[rest]
                out[np.nonzero(rest)[0][hit]] = self.n_own + order[pos[hit]] + 1
            return out.astype(np.int32)
        # long id lists on a moderate global range (FEM assembly: tens of millions of triplet ids per part): a dense table
        # gid -> local id, written so that the same copy wins as below (ghosts in ascending local order: the LAST ghost copy
        # of a gid stays; then the own ids on top); one gather per query instead of a binary search
        if len(gids) >= (1 << 16) and self.n_global <= (1 << 27):
            if getattr(self, "_g2l_dense", None) is None:
                dense = np.zeros(self.n_global + 1, dtype=np.int32)
                l2g = self.local_to_global
                gl = np.sort(self.ghost_to_local)
                dense[l2g[gl - 1]] = gl
                dense[l2g[self.own_to_local - 1]] = self.own_to_local
                self._g2l_dense = dense
            bad = (gids < 1) | (gids > self.n_global)
            if bad.any():
                gids = np.where(bad, 0, gids)  # entry 0 of the table is 0
            return self._g2l_dense[gids]
        if self._g2l is None:  # sorted table of the local gids (vectorised lookup: FEM-size id lists)
            l2g = self.local_to_global
            # duplicate gids (periodic ghost layers): the own id wins, else the LAST ghost copy (the reference's
            # global_to_ghost is a Dict filled in ghost order, src/p_range.jl:928-935)
            rank = -np.arange(1, self.n_local + 1, dtype=np.int64)
            rank[self.own_to_local - 1] = -(self.n_local + 1)
            order = np.lexsort((rank, l2g))
            self._g2l = (l2g[order], order)
        sg, order = self._g2l
        out = np.zeros(len(gids), dtype=np.int32)
        if len(sg):
            pos = np.clip(np.searchsorted(sg, gids, side="left"), 0, len(sg) - 1)
            hit = sg[pos] == gids
            out[hit] = order[pos[hit]] + 1
        return out


def uniform_partition_part(rank: int, np_: Sequence[int], n: Sequence[int], ghost=None, periodic=None) -> LocalIndices:
    """block_with_constant_size for one part (rank is 1-based, column-major over the part grid)."""
    np_, n = tuple(int(x) for x in np_), tuple(int(x) for x in n)
    D = len(n)
    coord, r = [], rank - 1
    for m in np_:
        coord.append(r % m + 1)
        r //= m
    own = tuple(local_range(coord[d], np_[d], n[d]) for d in range(D))
    info = BlockInfo(n, np_, own)
    nglobal = int(np.prod(n))
    if ghost is None:
        n_own = int(np.prod([hi - lo + 1 for lo, hi in own]))
        return LocalIndices(nglobal, rank, block=info, n_own=n_own)
    per = tuple(periodic) if periodic is not None else (False,) * D
    loc = tuple(local_range(coord[d], np_[d], n[d], bool(ghost[d]), bool(per[d])) for d in range(D))
    pts = _box_points(loc)
    wrapped = [np.mod(p - 1, n[d]) + 1 for d, p in enumerate(pts)]
    gids = _lin(wrapped, n)
    is_own = np.ones(len(gids), dtype=bool)
    for d in range(D):
        is_own &= (pts[d] >= own[d][0]) & (pts[d] <= own[d][1])
    owner = info.owner_of(gids)
    owner[is_own] = rank
    return LocalIndices(nglobal, rank, gids, owner, block=info, is_own=is_own)


def variable_partition_part(rank: int, n_own_all: Sequence[int], n_global: int) -> LocalIndices:
    start = 1 + int(np.sum(n_own_all[: rank - 1]))
    no = int(n_own_all[rank - 1])
    return LocalIndices(n_global, rank, np.arange(start, start + no), np.full(no, rank, dtype=np.int32))


def union_ghost(ind: LocalIndices, gids, owners) -> LocalIndices:
    """Append unseen off-part ids as ghosts in order of first appearance (ids<1 skipped)."""
    gids = np.asarray(gids, dtype=np.int64)
    owners = np.asarray(owners, dtype=np.int32)
    m = (gids >= 1) & (owners != ind.part)
    cand, cown = gids[m], owners[m]
    _, first = np.unique(cand, return_index=True)
    first.sort()
    cand, cown = cand[first], cown[first]
    if ind.n_ghost and len(cand):
        keep = ~np.isin(cand, ind.ghost_to_global)
        cand, cown = cand[keep], cown[keep]
    if len(cand) and not ind.own_is_prefix:
        raise ValueError("replace_ghost only makes sense for un-permuted local indices")
    if ind.block is not None and ind.own_is_prefix:
        return LocalIndices(ind.n_global, ind.part, block=ind.block, n_own=ind.n_own,
                            ghost_to_global=np.concatenate([ind.ghost_to_global, cand]),
                            ghost_to_owner=np.concatenate([ind.ghost_to_owner, cown]))
    is_own = np.zeros(ind.n_local + len(cand), dtype=bool)
    is_own[ind.own_to_local - 1] = True
    return LocalIndices(ind.n_global, ind.part, np.concatenate([ind.local_to_global, cand]),
                        np.concatenate([ind.local_to_owner, cown]), block=ind.block, is_own=is_own)


@dataclass
class PartPlan:
    """VectorAssemblyCache arrays of one part (src/p_vector.jl:418-426), 1-based, plus the
    neighbour-side local ids needed for one-sided peer access."""

    nbr_snd: np.ndarray
    snd_ptrs: np.ndarray
    snd_lids: np.ndarray
    snd_remote_lids: np.ndarray
    nbr_rcv: np.ndarray
    rcv_ptrs: np.ndarray
    rcv_lids: np.ndarray
    rcv_remote_lids: np.ndarray


def _ptrs(lengths) -> np.ndarray:
    p = np.ones(len(lengths) + 1, dtype=np.int32)
    if len(lengths):
        p[1:] = 1 + np.cumsum(lengths)
    return p


def build_plans(local_inds: List[LocalIndices], gather_all) -> List[PartPlan]:
    """Exchange plan of every local part.  ``gather_all(list_of_local_objects)`` returns the objects of
    ALL parts of the job ordered by part id (identity when every part is local; an all-gather over
    torch.distributed otherwise) — the analogue of the reference's gather/scatter neighbour discovery
    (src/primitives.jl:826-859) and gid exchange (src/p_range.jl:517-518)."""
    # round 1: who sends what (ghost lids grouped by owner, in local-id order) + the gids
    mine = []
    for ind in local_inds:
        gown, glid, ggid = ind.ghost_to_owner, ind.ghost_to_local, ind.ghost_to_global
        # ghosts owned by the part itself (periodic, one part in that direction) are never exchanged:
        # compute_assembly_neighbors skips owner == rank (src/p_range.jl:436-450)
        nbr = np.unique(gown[gown != ind.part]).astype(np.int32)
        if not ind.own_is_prefix:  # send lists follow local-id order (src/p_range.jl:506-513)
            o = np.argsort(glid, kind="stable")
            gown, glid, ggid = gown[o], glid[o], ggid[o]
        segs_l = [glid[gown == q].astype(np.int32) for q in nbr]
        segs_g = [ggid[gown == q] for q in nbr]
        mine.append({"part": ind.part, "nbr": nbr, "lids": segs_l, "gids": segs_g})
    everyone = gather_all(mine)
    by_part = {e["part"]: e for e in everyone}
    # round 2: receive side = transpose; my lids of the gids the neighbour listed
    rcv_info = []
    for ind in local_inds:
        nbr_rcv = np.array(sorted(q for q, e in by_part.items() if ind.part in e["nbr"].tolist()), dtype=np.int32)
        segs = []
        for q in nbr_rcv:
            e = by_part[int(q)]
            j = e["nbr"].tolist().index(ind.part)
            segs.append(ind.global_to_local(e["gids"][j]).astype(np.int32))
        rcv_info.append({"part": ind.part, "nbr": nbr_rcv, "lids": segs})
    everyone_rcv = gather_all(rcv_info)
    rcv_by_part = {e["part"]: e for e in everyone_rcv}
    plans = []
    for ind, snd, rcv in zip(local_inds, mine, rcv_info):
        snd_remote = []
        for i, q in enumerate(snd["nbr"]):
            e = rcv_by_part[int(q)]
            snd_remote.append(e["lids"][e["nbr"].tolist().index(ind.part)])
        rcv_remote = []
        for i, q in enumerate(rcv["nbr"]):
            e = by_part[int(q)]
            rcv_remote.append(e["lids"][e["nbr"].tolist().index(ind.part)])
        cat = lambda segs: np.concatenate(segs).astype(np.int32) if len(segs) else np.zeros(0, dtype=np.int32)
        plans.append(PartPlan(snd["nbr"], _ptrs([len(s) for s in snd["lids"]]), cat(snd["lids"]), cat(snd_remote),
                              rcv["nbr"], _ptrs([len(s) for s in rcv["lids"]]), cat(rcv["lids"]), cat(rcv_remote)))
    return plans


# ---------------------------------------------------------------------------- stencil fast path
def stencil_offsets(kind: int):
    """Neighbour emission order of the two generators, as (dx,dy,dz) lists.
    7: gallery laplacian_fdm (src/gallery.jl:60-78): diagonal, then d=1..3, i in (-1,+1).
    27: HPCG build_matrix (HPCG/src/sparse_matrix.jl:50-57): sz, sy, sx ascending."""
    if kind == 7:
        return [(0, 0, 0), (-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]
    return [(sx, sy, sz) for sz in (-1, 0, 1) for sy in (-1, 0, 1) for sx in (-1, 0, 1)]


def stencil_ghosts(kind: int, info: BlockInfo) -> np.ndarray:
    """ghost_to_global (1-based) of the column partition union_ghost would produce from the generator's
    COO column list, computed from the boundary layer of the own box only (first-appearance order)."""
    n, box = info.grid, info.box
    b = [hi - lo + 1 for lo, hi in box]
    # own cells lying on a face of the box, in own (column-major) order
    idx = []
    for d in range(3):
        for side in (0, b[d] - 1):
            rng = [np.arange(b[0]), np.arange(b[1]), np.arange(b[2])]
            rng[d] = np.array([side])
            g = np.meshgrid(*rng[::-1], indexing="ij")[::-1]
            idx.append(g[0].reshape(-1) + b[0] * (g[1].reshape(-1) + b[1] * g[2].reshape(-1)))
    cells = np.unique(np.concatenate(idx))
    cx, cy, cz = cells % b[0], (cells // b[0]) % b[1], cells // (b[0] * b[1])
    gx, gy, gz = cx + box[0][0], cy + box[1][0], cz + box[2][0]  # 1-based global coords
    offs = stencil_offsets(kind)
    cand = np.zeros((len(cells), len(offs)), dtype=np.int64)
    for k, (dx, dy, dz) in enumerate(offs):
        x, y, z = gx + dx, gy + dy, gz + dz
        in_grid = (x >= 1) & (x <= n[0]) & (y >= 1) & (y <= n[1]) & (z >= 1) & (z <= n[2])
        in_box = (x >= box[0][0]) & (x <= box[0][1]) & (y >= box[1][0]) & (y <= box[1][1]) & (z >= box[2][0]) & (z <= box[2][1])
        gid = (x - 1) + n[0] * ((y - 1) + n[1] * (z - 1)) + 1
        cand[:, k] = np.where(in_grid & ~in_box, gid, 0)
    flat = cand.reshape(-1)
    flat = flat[flat > 0]
    _, first = np.unique(flat, return_index=True)
    first.sort()
    return flat[first]


def stencil_col_indices(kind: int, rank: int, np_: Sequence[int], n: Sequence[int]) -> LocalIndices:
    """Column partition of the stencil operator for one part: own block + ghosts in reference order."""
    rows = uniform_partition_part(rank, np_, n)
    gh = stencil_ghosts(kind, rows.block)
    return LocalIndices(rows.n_global, rank, block=rows.block, n_own=rows.n_own, ghost_to_global=gh,
                        ghost_to_owner=rows.block.owner_of(gh))
