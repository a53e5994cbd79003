"""CPU ORACLE (test infrastructure) — HPCG multigrid preconditioner + preconditioned CG.

Restates HPCG/src/mg_preconditioner.jl (pc_setup :137-185, restrict!/prolongate! :224-251, pc_solve! :314-328,
ldiv! :202-206), the symmetric Gauss-Seidel smoother of PartitionedSolvers/src/smoothers.jl:82-125,162-176,248-269
and the preconditioned iteration of HPCG/src/ref_cg.jl:40-97.  Pinned by the reference's own known answer:
np=4, 32^3 per part, 4 levels, 50 iterations -> ||r||/||r0|| = 2.877476184683206e-13
(HPCG/test/hpcg_benchmark_tests.jl:31-41), see tests/test_oracle_golden.py."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import c_oracle
from . import pa_oracle as o


def restrict_operator(nx, ny, nz):
    """f2c (1-based fine local row of every coarse local row) — mg_preconditioner.jl:81-101."""
    nxc, nyc, nzc = nx // 2, ny // 2, nz // 2
    izc, iyc, ixc = np.meshgrid(np.arange(nzc), np.arange(nyc), np.arange(nxc), indexing="ij")
    fine = (2 * izc) * nx * ny + (2 * iyc) * nx + 2 * ixc
    return (fine.reshape(-1) + 1).astype(np.int32)


class Level:
    def __init__(self, nx, ny, nz, npd):
        self.n = (nx, ny, nz)
        self.A, self.r = o.hpcg_build_p_matrix(nx, ny, nz, *npd)
        self._setup_local()

    @classmethod
    def from_psparse(cls, A):
        """A smoother level for any assembled oracle PSparseMatrix (7-pt gallery operator, FEM stencil, ...)."""
        self = cls.__new__(cls)
        self.n, self.A, self.r = None, A, None
        self._setup_local()
        return self

    def _setup_local(self):
        self.part = self.A.col_partition
        self.plan = o.assembly_plan(self.part)
        self.x = [np.zeros(i.n_local) for i in self.part]
        self.Axf = [np.zeros(i.n_local) for i in self.part]
        self.mats = []
        self.diag = []
        for p, ind in enumerate(self.part):
            L = self.A.local[p]
            rp = np.ascontiguousarray(L.rowptr.astype(np.int64) - 1)
            cv = np.ascontiguousarray(L.colval.astype(np.int32) - 1)
            nz = np.ascontiguousarray(L.nzval)
            self.mats.append((ind.n_own, rp, cv, nz))
            rowid = np.repeat(np.arange(ind.n_own), np.diff(rp))
            d = np.zeros(ind.n_own)
            m = cv == rowid
            d[rowid[m]] = nz[m]  # dense_diag! (src/p_sparse_matrix.jl:2166-2188)
            self.diag.append(d)


def color_perm(box_dims, kind=27):
    """Rows of a box sorted by colour (27-pt: (ix&1) + 2(iy&1) + 4(iz&1), 7-pt: (ix+iy+iz)&1), ascending row id within a
    colour: the sweep order of the multi-colour smoother (csrc/pa_mg.cu k_levels_color)."""
    bx, by, bz = (int(d) for d in box_dims)
    i = np.arange(bx * by * bz)
    ix, iy, iz = i % bx, (i // bx) % by, i // (bx * by)
    col = (ix & 1) + 2 * (iy & 1) + 4 * (iz & 1) if kind == 27 else (ix + iy + iz) & 1
    return np.argsort(col, kind="stable").astype(np.int32)


def gs_sweep(level: Level, x, b, backward: bool, zero_guess: bool):
    assert c_oracle.available()
    lib = c_oracle._lib
    if getattr(level, "order", "lexicographic") == "multicolor":
        lib.pa_oracle_gs_sweep_perm.argtypes = [C.c_int64] + [C.c_void_p] * 8 + [C.c_int, C.c_int]
        lib.pa_oracle_gs_sweep_perm.restype = None
        for p, ind in enumerate(level.part):
            n, rp, cv, nz = level.mats[p]
            dims = [hi - lo + 1 for lo, hi in ind.box]
            perm = color_perm(dims, level.kind)
            done = np.zeros(n, dtype=np.uint8)
            bo = np.ascontiguousarray(b[p][:n])
            lib.pa_oracle_gs_sweep_perm(n, c_oracle._p(rp), c_oracle._p(cv), c_oracle._p(nz), c_oracle._p(level.diag[p]), c_oracle._p(bo),
                                        c_oracle._p(x[p]), c_oracle._p(perm), c_oracle._p(done), int(backward), int(zero_guess))
        return
    lib.pa_oracle_gs_sweep.argtypes = [C.c_int64] + [C.c_void_p] * 6 + [C.c_int, C.c_int]
    lib.pa_oracle_gs_sweep.restype = None
    for p in range(len(level.part)):
        n, rp, cv, nz = level.mats[p]
        bo = np.ascontiguousarray(b[p][:n])
        lib.pa_oracle_gs_sweep(n, c_oracle._p(rp), c_oracle._p(cv), c_oracle._p(nz), c_oracle._p(level.diag[p]), c_oracle._p(bo),
                               c_oracle._p(x[p]), int(backward), int(zero_guess))


def smooth(level: Level, x, b, zero_guess=False):
    """gauss_seidel(iterations=1, sweep=:symmetric) step — smoothers.jl:98-125."""
    if not zero_guess:
        o.consistent(x, level.plan)
    gs_sweep(level, x, b, False, zero_guess)
    gs_sweep(level, x, b, True, False)


class MG:
    def __init__(self, npd, levels, nx, ny, nz, order="lexicographic"):
        self.order = order
        self.l = levels
        self.levels = [None] * levels  # index l-1 = finest
        self.f2c = [None] * (levels - 1)
        self.levels[levels - 1] = Level(nx, ny, nz, npd)
        for i in reversed(range(levels - 1)):
            self.f2c[i] = restrict_operator(nx, ny, nz)
            nx, ny, nz = nx // 2, ny // 2, nz // 2
            self.levels[i] = Level(nx, ny, nz, npd)
        for L in self.levels:
            L.order, L.kind = order, 27

    def solve(self, x, b, l, zero_guess=False):
        """pc_solve! — mg_preconditioner.jl:314-328 (l is 1-based like the reference)."""
        L = self.levels[l - 1]
        if l == 1:
            smooth(L, x, b, zero_guess)
            return x
        smooth(L, x, b, zero_guess)
        o.mul_no_lat(L.A, x, L.plan, L.Axf)
        Lc = self.levels[l - 2]
        f2c = self.f2c[l - 2]
        rc = Lc.r
        for p in range(len(L.part)):
            rc[p][: len(f2c)] = b[p][f2c - 1] - L.Axf[p][f2c - 1]  # restrict! :224-229
            Lc.x[p][:] = 0.0
        self.solve(Lc.x, rc, l - 1, zero_guess=True)
        for p in range(len(L.part)):
            x[p][f2c - 1] += Lc.x[p][: len(f2c)]  # prolongate! :246-251
        smooth(L, x, b, False)
        return x

    def ldiv(self, x, b):
        """ldiv!(x, P, b) :202-206."""
        for v in x:
            v[:] = 0.0
        return self.solve(x, b, self.l, zero_guess=True)


def pcg(mg: MG, b_vals, x_vals, maxiter, tolerance=0.0):
    """ref_cg! with Pl = Mg_preconditioner (HPCG/src/ref_cg.jl:40-134)."""
    L = mg.levels[mg.l - 1]
    A, part, plan = L.A, L.part, L.plan
    u = [np.zeros_like(x) for x in x_vals]
    r = [b.copy() for b in b_vals]
    c = [np.zeros_like(x) for x in x_vals]
    o.pmul(A, x_vals, plan, c)
    for p in range(len(r)):
        r[p] -= c[p]
    residual0 = residual = o.pnorm(r, part)
    rho, it, hist = 1.0, 0, [residual]
    while not (it >= maxiter or residual / residual0 <= tolerance):
        mg.ldiv(c, r)
        rho_prev, rho = rho, o.pdot(c, r, part)
        beta = rho / rho_prev
        for p in range(len(r)):
            u[p][:] = c[p] + beta * u[p]
        o.mul_no_lat(A, u, plan, c)
        alpha = rho / o.pdot(u, c, part)
        for p in range(len(r)):
            x_vals[p] += alpha * u[p]
            r[p] -= alpha * c[p]
        residual = o.pnorm(r, part)
        hist.append(residual)
        it += 1
    return x_vals, residual0, residual, it, hist
