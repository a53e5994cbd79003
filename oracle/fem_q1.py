"""CPU ORACLE — test infrastructure, NOT product code.

Restatement of the reference's FEM example DRIVER (test/fem_example.jl:12-262): the user-side code that turns a
2-D Q1 grid into disassembled triplets (I,J,V), right-hand-side contributions (II,VV) and the exact solution.  It is
the input generator of BASELINE config C5 ("fem_example.jl unstructured assembly -> PSparseMatrix"); the matrix
assembly itself (psparse / pvector / cg) is what the oracle (pa_oracle.psparse_disassembled, pvector_disassembled) and
the product (pa_b200.psparse(assembled=False), pvector_from_triplets) are compared on.

What the driver computes, in closed form (derived from the loops it follows):
  * cells are block-partitioned with one layer of ghost cells (fem_example.jl:270-271); a node is a free dof iff it is
    not on the boundary (:72-79); a dof is owned by the largest part id among the cells around its node (:84-97) = the
    owner of the cell to its upper right (part ids grow with the cell coordinates);
  * a part numbers its own dofs in the order of its local nodes (column-major, :81-83,121-123), after the offsets of
    variable_partition(n_own_dofs) (:274);
  * triplets: own cells in local (column-major) order; per cell the 4x4 element matrix row by row, element nodes in
    column-major order of the 2x2 reference cell, boundary rows/columns skipped (:169-199);
  * rhs: per own cell ge = Ae*ue with ue = u at the cell's boundary nodes (0 elsewhere), contribution -ge[row] to
    every free row of the cell (:201-235).  `mul!(ge,Ae,ue)` is a dense 4x4 product inside LinearAlgebra/BLAS: its
    summation order is NOT pinned by the reference (column-by-column accumulation is used here).
Parity status: pinned by the reference's own known answer, norm(x - x_exact) < 1e-5 after cg (fem_example.jl:289),
see tests/test_oracle_fem.py.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from .pa_oracle import local_range


def element_matrix(h: float) -> np.ndarray:
    """fem_example.jl:22-27"""
    return (h * h / 6.0) * np.array([[4.0, -1.0, -1.0, -2.0], [-1.0, 4.0, -2.0, -1.0], [-1.0, -2.0, 4.0, -1.0], [-2.0, -1.0, -1.0, 4.0]])


def u_exact(x, y):
    """fem_example.jl:11"""
    return x + y


class Q1Problem:
    """Everything the driver hands to psparse/pvector, for all parts (1-based global dof ids)."""

    def __init__(self, parts_per_dir: Sequence[int] = (2, 2), cells_per_dir: Sequence[int] = (10, 10), length_per_dir=(2.0, 2.0)):
        px, py = (int(p) for p in parts_per_dir)
        cx, cy = (int(c) for c in cells_per_dir)
        self.parts_per_dir, self.cells_per_dir = (px, py), (cx, cy)
        self.h = max(length_per_dir[0] / cx, length_per_dir[1] / cy)
        self.Ae = element_matrix(self.h)
        nparts = px * py
        # own cell ranges (inclusive, 1-based) and own dof boxes: node (i,j) is free iff 2<=i<=cx, 2<=j<=cy and is owned by the
        # owner of cell (i,j)
        self.cell_box, self.dof_box = [], []
        for rank in range(1, nparts + 1):
            a, b = (rank - 1) % px + 1, (rank - 1) // px + 1
            rx, ry = local_range(a, px, cx), local_range(b, py, cy)
            self.cell_box.append((rx, ry))
            self.dof_box.append(((max(rx[0], 2), min(rx[1], cx)), (max(ry[0], 2), min(ry[1], cy))))
        self.n_own_dofs = [max(0, bx[1] - bx[0] + 1) * max(0, by[1] - by[0] + 1) for bx, by in self.dof_box]
        self.n_global_dofs = int(sum(self.n_own_dofs))
        self.offset = np.concatenate([[0], np.cumsum(self.n_own_dofs)])[:-1]
        # per-dimension owner coordinate of a cell index
        self._own_x = self._owner_1d(px, cx)
        self._own_y = self._owner_1d(py, cy)
        self.I, self.J, self.V, self.II, self.VV = [], [], [], [], []
        for rank in range(1, nparts + 1):
            i, j, v, ii, vv = self._part_contributions(rank)
            self.I.append(i); self.J.append(j); self.V.append(v); self.II.append(ii); self.VV.append(vv)

    @staticmethod
    def _owner_1d(np_, n):
        out = np.zeros(n + 2, dtype=np.int64)  # 1-based cell index -> 1-based part coordinate
        for p in range(1, np_ + 1):
            lo, hi = local_range(p, np_, n)
            out[lo : hi + 1] = p
        return out

    def global_dof(self, i, j):
        """global dof id of node (i,j) (1-based node coordinates), 0 for boundary nodes."""
        i, j = np.asarray(i, dtype=np.int64), np.asarray(j, dtype=np.int64)
        cx, cy = self.cells_per_dir
        free = (i >= 2) & (i <= cx) & (j >= 2) & (j <= cy)
        ic, jc = np.clip(i, 1, cx), np.clip(j, 1, cy)
        rank = self._own_x[ic] + (self._own_y[jc] - 1) * self.parts_per_dir[0]  # 1-based
        out = np.zeros(i.shape, dtype=np.int64)
        for r in np.unique(rank[free]):
            (x0, x1), (y0, _) = self.dof_box[r - 1]
            m = free & (rank == r)
            out[m] = self.offset[r - 1] + (i[m] - x0) + (j[m] - y0) * (x1 - x0 + 1) + 1
        return out

    def _part_contributions(self, rank):
        (x0, x1), (y0, y1) = self.cell_box[rank - 1]
        cx, cy = self.cells_per_dir
        # own cells in column-major order; element nodes (0,0),(1,0),(0,1),(1,1)
        ci, cj = np.meshgrid(np.arange(x0, x1 + 1), np.arange(y0, y1 + 1), indexing="xy")
        ci, cj = ci.reshape(-1), cj.reshape(-1)  # x fastest
        di, dj = np.array([0, 1, 0, 1]), np.array([0, 0, 1, 1])
        ni, nj = ci[:, None] + di[None, :], cj[:, None] + dj[None, :]  # (ncell, 4) node coordinates
        dofs = self.global_dof(ni, nj)  # 0 on the boundary
        rows = np.broadcast_to(dofs[:, :, None], dofs.shape + (4,))
        cols = np.broadcast_to(dofs[:, None, :], dofs.shape[:1] + (4, 4))
        vals = np.broadcast_to(self.Ae[None, :, :], rows.shape)
        ok = (rows > 0) & (cols > 0)
        I, J, V = rows[ok], cols[ok], vals[ok]  # C order = cell, element row, element col
        # rhs
        ue = np.where(dofs <= 0, u_exact((ni - 1) * self.h, (nj - 1) * self.h), 0.0)
        ge = np.zeros(ue.shape)
        for c in range(4):  # column-by-column accumulation (summation order unpinned, see header)
            ge = ge + self.Ae[None, :, c] * ue[:, c : c + 1]
        okr = dofs > 0
        return I.astype(np.int64), J.astype(np.int64), V.astype(np.float64), dofs[okr].astype(np.int64), (-ge)[okr]

    def exact_solution(self) -> np.ndarray:
        """u at every free node, indexed by global dof id - 1 (fem_example.jl:237-262)."""
        cx, cy = self.cells_per_dir
        i, j = np.meshgrid(np.arange(2, cx + 1), np.arange(2, cy + 1), indexing="xy")
        g = self.global_dof(i.reshape(-1), j.reshape(-1))
        out = np.zeros(self.n_global_dofs)
        out[g - 1] = u_exact((i.reshape(-1) - 1) * self.h, (j.reshape(-1) - 1) * self.h)
        return out
