"""HPCG multigrid preconditioner on the CUDA backend — host-side mirror of HPCG/src/mg_preconditioner.jl.

  pc_setup(backend, l, nx, ny, nz, npx, npy, npz) -> MgPreconditioner   (mg_preconditioner.jl:137-185)
  ldiv_(x, P, b)                                                          (:202-206)
  ref_cg_(x, A, b; Pl=P)                                                  (HPCG/src/ref_cg.jl:119-134)
  GaussSeidel(A).smooth_(x, b, zero_guess)                                (PartitionedSolvers/src/smoothers.jl:82-125)
All arithmetic happens in libpa_b200 (csrc/pa_mg.cu); this file only builds the per-level operators."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from . import _capi
from ._capi import check, i64, ptr
from .gallery import stencil_matrix
from .parrays import CGResult, CUDAArray, PSparseMatrix, PVector


class GaussSeidel:
    """gauss_seidel(p; iterations=1, sweep=:symmetric) for a PSparseMatrix (bit-exact wavefront sweeps)."""

    def __init__(self, A: PSparseMatrix, kind: Optional[int] = None):
        self.A = A
        h = C.c_void_p()
        L = _capi.lib()
        check(L.pa_gs_create(A.h, C.byref(h)))
        self.h = h
        if kind is not None:
            for k, ind in enumerate(A.rows.indices):
                dims = i64([hi - lo + 1 for lo, hi in ind.block.box])
                check(L.pa_gs_set_box(h, k, kind, ptr(dims)))
        check(L.pa_gs_commit(h))
        self.order = "lexicographic"

    def set_order(self, order: str = "lexicographic"):
        """"lexicographic": the reference's sequential sweep order (bit-identical iterates; default).  "multicolor": colour by
        colour (needs the geometry hint): same per-row arithmetic and fixed point, different iterates."""
        code = {"lexicographic": _capi.PA_GS_LEXICOGRAPHIC, "multicolor": _capi.PA_GS_MULTICOLOR}.get(order)
        if code is None:
            raise ValueError(f"unknown Gauss-Seidel order {order!r}")
        check(_capi.lib().pa_gs_set_order(self.h, code))
        self.order = order
        return self

    def smooth_(self, x: PVector, b: PVector, zero_guess: bool = False) -> PVector:
        check(_capi.lib().pa_gs_smooth(self.h, x.h, b.h, int(zero_guess)))
        return x

    def free(self):
        if getattr(self, "h", None) and self.A.backend.h:
            _capi.lib().pa_gs_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MgPreconditioner:
    """Mg_preconditioner: A_vec / gs_states / r / x / Axf per level (index 0 = coarsest, l-1 = finest)."""

    def __init__(self, backend: CUDAArray, levels: int, nx: int, ny: int, nz: int, npx: int, npy: int, npz: int, order: str = "lexicographic"):
        assert nx % (1 << (levels - 1)) == 0 and ny % (1 << (levels - 1)) == 0 and nz % (1 << (levels - 1)) == 0
        self.backend, self.l = backend, levels
        self.A_vec: List[PSparseMatrix] = [None] * levels
        self.b_vec: List[PVector] = [None] * levels
        self.gs: List[GaussSeidel] = [None] * levels
        dims = np.zeros((levels, len(backend.parts), 3), dtype=np.int64)
        for lev in reversed(range(levels)):  # finest first, like pc_setup
            f = 1 << (levels - 1 - lev)
            lx, ly, lz = nx // f, ny // f, nz // f
            A, b = stencil_matrix(27, (lx * npx, ly * npy, lz * npz), (npx, npy, npz), backend)
            self.A_vec[lev], self.b_vec[lev] = A, b
            self.gs[lev] = GaussSeidel(A, kind=27)
            for k, ind in enumerate(A.rows.indices):
                dims[lev, k] = [hi - lo + 1 for lo, hi in ind.block.box]
        Ah = (C.c_void_p * levels)(*[A.h for A in self.A_vec])
        Gh = (C.c_void_p * levels)(*[g.h for g in self.gs])
        h = C.c_void_p()
        self._dims = np.ascontiguousarray(dims)
        check(_capi.lib().pa_mg_create(levels, Ah, Gh, ptr(self._dims), C.byref(h)))
        self.h = h
        self.order = "lexicographic"
        if order != "lexicographic":
            self.set_order(order)

    def set_order(self, order: str):
        """Sweep order of the smoother of every level (see GaussSeidel.set_order)."""
        for g in self.gs:
            g.set_order(order)
        self.order = order
        return self

    @property
    def A(self) -> PSparseMatrix:
        return self.A_vec[self.l - 1]

    @property
    def b(self) -> PVector:
        return self.b_vec[self.l - 1]

    def ldiv_(self, x: PVector, b: PVector) -> PVector:
        check(_capi.lib().pa_mg_apply(self.h, x.h, b.h))
        return x

    def free(self):
        if getattr(self, "h", None) and self.backend.h:
            _capi.lib().pa_mg_destroy(self.h)
            for g in self.gs:
                g.free()
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pc_setup(backend: CUDAArray, l: int, nx: int, ny: int, nz: int, npx: int, npy: int, npz: int, order: str = "lexicographic") -> MgPreconditioner:
    return MgPreconditioner(backend, l, nx, ny, nz, npx, npy, npz, order)


def ref_cg_pc_(x: PVector, A: PSparseMatrix, b: PVector, Pl: Optional[MgPreconditioner], tolerance: float = 0.0, maxiter: int = 50,
               flags: int = 0) -> CGResult:
    """ref_cg!(x, A, b; tolerance, maxiter, Pl) (HPCG/src/ref_cg.jl:119-134)."""
    res = _capi.CGResult()
    hist = np.zeros(maxiter + 1, dtype=np.float64)
    check(_capi.lib().pa_cg_precond(A.h, x.h, b.h, Pl.h if Pl is not None else None, maxiter, float(tolerance), flags, C.byref(res), ptr(hist)))
    return CGResult(res.iters, bool(res.converged), res.residual0, res.residual, hist[: res.iters + 1])


def smoother_name(P: "MgPreconditioner") -> str:
    """Which Gauss-Seidel schedule a preconditioner runs (reported by bench.py)."""
    if P.order == "multicolor":
        return "multi-colour Gauss-Seidel (8 colours, one launch per colour; convergence-level parity: different iterates than the reference)"
    return "bit-exact wavefront Gauss-Seidel (same iterates as the reference's sequential sweeps)"
