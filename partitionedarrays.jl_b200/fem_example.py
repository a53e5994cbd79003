"""The reference's FEM example driver (test/fem_example.jl) on the CUDA backend: BASELINE config C5.

  q1_part(rank, parts_per_dir, cells_per_dir, length_per_dir)  triplets (I,J,V), rhs contributions (II,VV) of ONE part
  fem_example(backend, parts_per_dir, cells_per_dir)           -> A, rhs, info   (psparse + pvector on the device)

The driver is user-side code in the reference (a 2-D Q1 Poisson problem, u = x + y): cells are block partitioned, every
part loops over its own cells and emits the 4x4 element matrix `Ae = (h^2/6)[4 -1 -1 -2; ...]` (fem_example.jl:22-27) for
the free (interior) nodes, boundary values go to the right-hand side (:169-235).  What the library then does with the
disassembled triplets — psparse(I,J,V,rows,cols) / pvector(II,VV,rows) / psparse! / mul! / cg — is the path this package
accelerates.  Each process generates only the triplets of the parts it holds (closed form of the driver's loops):
  * node (i,j), 1-based, is a free dof iff 2 <= i <= cx and 2 <= j <= cy (:72-79);
  * a dof is owned by the largest part id among the cells around its node (:84-97) = the owner of cell (i,j);
  * a part numbers its own dofs column-major inside its dof box, after the offsets of variable_partition (:121-123,274);
  * triplets: own cells column-major, per cell the element matrix row by row, element nodes (0,0),(1,0),(0,1),(1,1)."""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

from .prange import local_range


def _element_matrix(h: float) -> np.ndarray:
    return (h * h / 6.0) * np.array([[4.0, -1.0, -1.0, -2.0], [-1.0, 4.0, -2.0, -1.0], [-1.0, -2.0, 4.0, -1.0], [-2.0, -1.0, -1.0, 4.0]])


class Q1Layout:
    """Partition of the cells / dofs of the example (closed form; no per-node arrays of the whole grid)."""

    def __init__(self, parts_per_dir: Sequence[int], cells_per_dir: Sequence[int], length_per_dir: Sequence[float] = (2.0, 2.0)):
        self.px, self.py = (int(p) for p in parts_per_dir)
        self.cx, self.cy = (int(c) for c in cells_per_dir)
        self.h = max(length_per_dir[0] / self.cx, length_per_dir[1] / self.cy)
        self.Ae = _element_matrix(self.h)
        self.nparts = self.px * self.py
        self.cell_box, self.dof_box = [], []
        for rank in range(1, self.nparts + 1):
            a, b = (rank - 1) % self.px + 1, (rank - 1) // self.px + 1
            rx, ry = local_range(a, self.px, self.cx), local_range(b, self.py, self.cy)
            self.cell_box.append((rx, ry))
            self.dof_box.append(((max(rx[0], 2), min(rx[1], self.cx)), (max(ry[0], 2), min(ry[1], self.cy))))
        self.n_own_dofs = [max(0, bx[1] - bx[0] + 1) * max(0, by[1] - by[0] + 1) for bx, by in self.dof_box]
        self.n_global_dofs = int(sum(self.n_own_dofs))
        self.offset = np.concatenate([[0], np.cumsum(self.n_own_dofs)])[:-1]
        self._sx = np.array([local_range(p, self.px, self.cx)[0] for p in range(1, self.px + 1)])
        self._sy = np.array([local_range(p, self.py, self.cy)[0] for p in range(1, self.py + 1)])

    def global_dof(self, i, j) -> np.ndarray:
        """1-based global dof id of node (i,j); 0 on the boundary."""
        i, j = np.asarray(i, dtype=np.int64), np.asarray(j, dtype=np.int64)
        free = (i >= 2) & (i <= self.cx) & (j >= 2) & (j <= self.cy)
        ic, jc = np.clip(i, 1, self.cx), np.clip(j, 1, self.cy)
        pxc = np.searchsorted(self._sx, ic, side="right")  # 1-based part coordinate of the owning cell
        pyc = np.searchsorted(self._sy, jc, side="right")
        rank = pxc + (pyc - 1) * self.px
        x0 = np.array([b[0][0] for b in self.dof_box])[rank - 1]
        x1 = np.array([b[0][1] for b in self.dof_box])[rank - 1]
        y0 = np.array([b[1][0] for b in self.dof_box])[rank - 1]
        out = self.offset[rank - 1] + (i - x0) + (j - y0) * (x1 - x0 + 1) + 1
        return np.where(free, out, 0).astype(np.int64)

    def exact_own(self, rank: int) -> np.ndarray:
        """u = x + y at the own dofs of `rank`, in own order."""
        (x0, x1), (y0, y1) = self.dof_box[rank - 1]
        i, j = np.meshgrid(np.arange(x0, x1 + 1), np.arange(y0, y1 + 1), indexing="xy")
        return ((i.reshape(-1) - 1) * self.h + (j.reshape(-1) - 1) * self.h).astype(np.float64)


def q1_part(lay: Q1Layout, rank: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """(I, J, V, II, VV) emitted by part `rank` (global 1-based dof ids)."""
    (x0, x1), (y0, y1) = lay.cell_box[rank - 1]
    ci, cj = np.meshgrid(np.arange(x0, x1 + 1), np.arange(y0, y1 + 1), indexing="xy")
    ci, cj = ci.reshape(-1), cj.reshape(-1)  # x fastest: the driver's column-major cell loop
    di, dj = np.array([0, 1, 0, 1]), np.array([0, 0, 1, 1])
    ni, nj = ci[:, None] + di[None, :], cj[:, None] + dj[None, :]
    dofs = lay.global_dof(ni, nj)
    rows = np.broadcast_to(dofs[:, :, None], dofs.shape + (4,))
    cols = np.broadcast_to(dofs[:, None, :], dofs.shape[:1] + (4, 4))
    vals = np.broadcast_to(lay.Ae[None, :, :], rows.shape)
    ok = (rows > 0) & (cols > 0)
    ue = np.where(dofs <= 0, (ni - 1) * lay.h + (nj - 1) * lay.h, 0.0)
    ge = np.zeros(ue.shape)
    for c in range(4):
        ge = ge + lay.Ae[None, :, c] * ue[:, c : c + 1]
    okr = dofs > 0
    return rows[ok].astype(np.int64), cols[ok].astype(np.int64), vals[ok].astype(np.float64), dofs[okr].astype(np.int64), (-ge)[okr]


def fem_example(backend, parts_per_dir: Sequence[int], cells_per_dir: Sequence[int], length_per_dir=(2.0, 2.0), compress: str = "device",
                ship: str = "device", local_format: str = "csr"):
    """The example's assembly on the CUDA backend: A = psparse(I,J,V,rows,rows) (disassembled input, the reference's
    default), rhs = pvector(II,VV,rows).  Returns (A, rhs, layout, triplets) — `triplets` per local part for psparse!."""
    from .parrays import psparse, pvector_from_triplets, variable_partition

    lay = Q1Layout(parts_per_dir, cells_per_dir, length_per_dir)
    assert lay.nparts == backend.nparts
    trip = [q1_part(lay, p) for p in backend.parts]
    rows = variable_partition(backend, lay.n_own_dofs, lay.n_global_dofs)
    A = psparse([t[0] for t in trip], [t[1] for t in trip], [t[2] for t in trip], rows, rows, assembled=False, local_format=local_format,
                compress=compress, ship=ship)
    rhs = pvector_from_triplets([t[3] for t in trip], [t[4] for t in trip], rows)
    return A, rhs, lay, trip
