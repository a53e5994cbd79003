"""Input generators of the benchmark configs on the CUDA backend.

  laplacian_fdm        gallery 7-pt FDM Laplacian as COO (src/gallery.jl:12-98) — host arrays, for psparse()
  stencil_matrix       the same operator (kind=7) or the HPCG 27-pt operator (kind=27, HPCG/src/sparse_matrix.jl:27-122)
                       generated directly as per-part CSR on the GPU: 512^3 per part never exists as COO on the host
  build_p_matrix       HPCG build_p_matrix (HPCG/src/sparse_matrix.jl:105-122): (A, b) with b = 27 - nnz_row
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import check, i32, i64, ptr
from . import prange as pr
from .parrays import CUDAArray, PRange, PSparseMatrix, PVector, uniform_partition


def laplacian_fdm(nodes_per_dir: Sequence[int], parts_per_dir: Sequence[int], backend: CUDAArray):
    """I,J,V (per local part, global 1-based ids), row partition, col partition — src/gallery.jl:12-86.
    Per own node: diagonal alpha*2D, then d=1..D, i in (-1,+1): -alpha when the neighbour is inside the grid;
    alpha = prod(n_i+1)."""
    n = tuple(int(x) for x in nodes_per_dir)
    D = len(n)
    alpha = float(np.prod([m + 1 for m in n]))
    rows = uniform_partition(backend, parts_per_dir, n)
    Is, Js, Vs = [], [], []
    for ind in rows.indices:
        pts = pr._box_points(ind.block.box)
        gids = pr._lin(pts, n)
        cols, vals, ok = [gids], [np.full(len(gids), alpha * 2 * D)], [np.ones(len(gids), dtype=bool)]
        for d in range(D):
            for s in (-1, 1):
                q = [p.copy() for p in pts]
                q[d] = q[d] + s
                inside = (q[d] >= 1) & (q[d] <= n[d])
                q[d] = np.clip(q[d], 1, n[d])
                cols.append(pr._lin(q, n)); vals.append(np.full(len(gids), -alpha)); ok.append(inside)
        Cm, Vm, M = np.stack(cols, 1), np.stack(vals, 1), np.stack(ok, 1)
        R = np.repeat(gids[:, None], Cm.shape[1], axis=1)
        Is.append(R[M]); Js.append(Cm[M]); Vs.append(Vm[M])
    return Is, Js, Vs, rows, rows


def stencil_matrix(kind: int, nodes_per_dir: Sequence[int], parts_per_dir: Sequence[int], backend: CUDAArray,
                   with_rhs: bool = True) -> Tuple[PSparseMatrix, PVector]:
    """Device-generated PSparseMatrix of the 7-pt (gallery) or 27-pt (HPCG) operator on a uniform partition and,
    optionally, the vector b (kind 27: 27 - nnz_row, i.e. A*ones; kind 7: A*ones) on the column partition."""
    n = tuple(int(x) for x in nodes_per_dir)
    assert len(n) == 3 and kind in (7, 27)
    rows = uniform_partition(backend, parts_per_dir, n)
    cols = PRange(backend, [pr.stencil_col_indices(kind, p, parts_per_dir, n) for p in backend.parts])
    A = PSparseMatrix(rows, cols)
    b = PVector(cols) if with_rhs else None
    L = _capi.lib()
    gn = i64(n)
    for k, ind in enumerate(cols.indices):
        lo = i64([r[0] - 1 for r in ind.block.box])
        hi = i64([r[1] for r in ind.block.box])
        g0 = ind.ghost_to_global - 1
        order = np.argsort(g0, kind="stable")
        sg, sl = i64(g0[order]), i32(order)
        check(L.pa_mat_set_stencil(A.h, k, kind, ptr(gn), ptr(lo), ptr(hi), len(sg), ptr(sg), ptr(sl), b.h if b is not None else None))
    A.commit()
    return A, b


def build_p_matrix(backend: CUDAArray, nx: int, ny: int, nz: int, npx: int, npy: int, npz: int):
    """HPCG build_p_matrix(ranks,nx,ny,nz,gnx,gny,gnz,npx,npy,npz) -> A, b (HPCG/src/sparse_matrix.jl:105-122)."""
    return stencil_matrix(27, (nx * npx, ny * npy, nz * npz), (npx, npy, npz), backend)


def fill_hash(v: PVector, seed: int) -> PVector:
    """v[own] = hash(gid, seed) in [-1,1), ghosts 0 — deterministic by global id (block partitions only)."""
    L = _capi.lib()
    for k, ind in enumerate(v.rows.indices):
        blk = ind.block
        assert blk is not None and len(blk.grid) == 3
        check(L.pa_vec_fill_hash_box(v.h, k, ptr(i64(blk.grid)), ptr(i64([r[0] - 1 for r in blk.box])), ptr(i64([r[1] for r in blk.box])), seed))
    return v


def compute_optimal_shape_xyz(np_: int) -> Tuple[int, int, int]:
    """Part grid for np parts.  The configs use 1->(1,1,1), 2->(2,1,1), 4->(2,2,1), 8->(2,2,2)
    (HPCG/src/compute_optimal_xyz.jl:8-64 for these counts); other counts: most cubic factorisation, x >= y >= z."""
    best = None
    for x in range(1, np_ + 1):
        if np_ % x:
            continue
        for y in range(1, np_ // x + 1):
            if (np_ // x) % y:
                continue
            z = np_ // x // y
            if x >= y >= z:
                cand = (x - z, (x, y, z))
                best = cand if best is None or cand < best else best
    return best[1]
