"""ctypes loader for the C twin of the oracle (oracle/pa_oracle.c).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpa_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pa_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        cc = "/usr/bin/gcc" if os.access("/usr/bin/gcc", os.X_OK) else "gcc"
        base = [cc, "-O2", "-fno-fast-math", "-ffp-contract=off", "-fPIC", "-std=c11", "-shared", "-o", _SO, src, "-lm"]
        try:
            subprocess.run(base[:1] + ["-fopenmp"] + base[1:], check=True, capture_output=True)
        except subprocess.CalledProcessError:
            subprocess.run(base, check=True, capture_output=True)
    return _SO


def available() -> bool:
    global _lib
    if _lib is not None:
        return True
    try:
        build()
        _lib = C.CDLL(_SO)
    except Exception:
        return False
    L = _lib
    L.pa_oracle_spmv_csr.argtypes = [C.c_int64] + [C.c_void_p] * 6
    L.pa_oracle_spmv_csr.restype = None
    L.pa_oracle_dot.argtypes = [C.c_int64, C.c_void_p, C.c_void_p]
    L.pa_oracle_dot.restype = C.c_double
    L.pa_oracle_stencil_csr.argtypes = [C.c_int] + [C.c_void_p] * 3 + [C.c_int64] + [C.c_void_p] * 6
    L.pa_oracle_stencil_csr.restype = C.c_int64
    L.pa_oracle_cg.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    L.pa_oracle_cg.restype = C.c_int
    L.pa_oracle_time_spmv.argtypes = [C.c_int, C.c_void_p, C.c_int]
    L.pa_oracle_time_spmv.restype = C.c_double
    L.pa_oracle_max_threads.restype = C.c_int
    return True


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def max_threads() -> int:
    assert available()
    return int(_lib.pa_oracle_max_threads())


def spmv_csr(A, x: np.ndarray, y0: Optional[np.ndarray] = None) -> np.ndarray:
    """A: oracle CSR (1-based).  Sequential-order y = [y0 +] A x."""
    assert available()
    rp = np.ascontiguousarray(A.rowptr.astype(np.int64) - 1)
    cv = np.ascontiguousarray(A.colval.astype(np.int32) - 1)
    nz = np.ascontiguousarray(A.nzval, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert len(x) == A.n
    y = np.zeros(A.m, dtype=np.float64)
    y0c = None if y0 is None else np.ascontiguousarray(y0, dtype=np.float64)
    _lib.pa_oracle_spmv_csr(A.m, _p(rp), _p(cv), _p(nz), _p(x), _p(y0c), _p(y))
    return y


def spmv_csr0(m, rowptr0, colval0, nzval, x) -> np.ndarray:
    """0-based arrays (int64 rowptr, int32 colval)."""
    assert available()
    y = np.zeros(m, dtype=np.float64)
    _lib.pa_oracle_spmv_csr(m, _p(rowptr0), _p(colval0), _p(nzval), _p(x), None, _p(y))
    return y


def stencil_csr(kind: int, gn, lo, hi, ghost_gids0: np.ndarray, want_b: bool = True):
    """Unsplit local CSR (0-based; cols: own box id, then n_own + ghost id) of the 7-/27-pt stencil
    on the own box [lo,hi) of a gn grid; ghost_gids0 = 0-based ghost gids in ghost-id order."""
    assert available()
    gn = np.asarray(gn, dtype=np.int64); lo = np.asarray(lo, dtype=np.int64); hi = np.asarray(hi, dtype=np.int64)
    order = np.argsort(ghost_gids0, kind="stable")
    sg = np.ascontiguousarray(ghost_gids0[order].astype(np.int64))
    sl = np.ascontiguousarray(order.astype(np.int32))
    n_own = int(np.prod(hi - lo))
    rowptr = np.zeros(n_own + 1, dtype=np.int64)
    nnz = _lib.pa_oracle_stencil_csr(kind, _p(gn), _p(lo), _p(hi), len(sg), _p(sg), _p(sl), _p(rowptr), None, None, None)
    if nnz < 0:
        raise RuntimeError("stencil_csr: ghost lookup failed")
    colval = np.zeros(nnz, dtype=np.int32)
    nzval = np.zeros(nnz, dtype=np.float64)
    b = np.zeros(n_own, dtype=np.float64) if want_b else None
    _lib.pa_oracle_stencil_csr(kind, _p(gn), _p(lo), _p(hi), len(sg), _p(sg), _p(sl), _p(rowptr), _p(colval), _p(nzval), _p(b))
    return rowptr, colval, nzval, b


class _Part(C.Structure):
    _fields_ = [
        ("n_own", C.c_int64), ("n_local", C.c_int64),
        ("rowptr", C.c_void_p), ("colval", C.c_void_p), ("nzval", C.c_void_p),
        ("n_snd", C.c_int32), ("nbr_snd", C.c_void_p), ("snd_ptrs", C.c_void_p), ("snd_lids", C.c_void_p),
        ("n_rcv", C.c_int32), ("nbr_rcv", C.c_void_p), ("rcv_ptrs", C.c_void_p), ("rcv_lids", C.c_void_p),
        ("buf_snd", C.c_void_p), ("buf_rcv", C.c_void_p),
        ("b", C.c_void_p), ("x", C.c_void_p), ("r", C.c_void_p), ("c", C.c_void_p), ("u", C.c_void_p),
    ]


class _Times(C.Structure):
    _fields_ = [("t_total", C.c_double), ("t_spmv", C.c_double), ("t_dot", C.c_double), ("t_waxpby", C.c_double), ("t_exch", C.c_double)]


class CGProblem:
    """Multi-part problem for pa_oracle_cg.  parts: list of dicts with 0-based arrays:
    n_own, n_local, rowptr(int64), colval(int32), nzval, b(n_local), x(n_local),
    and the *assembly* plan (1-based, as produced by pa_oracle.assembly_plan) per part."""

    def __init__(self, mats, plan, b_vals, x_vals):
        assert available()
        self.keep = []
        n = len(mats)
        self.arr = (_Part * n)()
        self.x = [np.ascontiguousarray(x, dtype=np.float64).copy() for x in x_vals]
        self.c = [np.zeros_like(x) for x in self.x]
        self.u = [np.zeros_like(x) for x in self.x]
        self.r = [np.zeros_like(x) for x in self.x]
        for p in range(n):
            n_own, n_local, rp, cv, nz = mats[p]
            k = lambda a, dt: self._k(np.ascontiguousarray(a, dtype=dt))
            P = self.arr[p]
            P.n_own, P.n_local = n_own, n_local
            P.rowptr, P.colval, P.nzval = _p(k(rp, np.int64)), _p(k(cv, np.int32)), _p(k(nz, np.float64))
            # consistent! = reversed assembly plan: send own lids (assembly rcv), receive ghost lids (assembly snd)
            snd_n, rcv_n = plan.neighbors_rcv[p], plan.neighbors_snd[p]
            ls, lr = plan.local_indices_rcv[p], plan.local_indices_snd[p]
            P.n_snd, P.n_rcv = len(snd_n), len(rcv_n)
            P.nbr_snd = _p(k(np.asarray(snd_n) - 1, np.int32)); P.nbr_rcv = _p(k(np.asarray(rcv_n) - 1, np.int32))
            P.snd_ptrs = _p(k(ls.ptrs - 1, np.int32)); P.snd_lids = _p(k(ls.data - 1, np.int32))
            P.rcv_ptrs = _p(k(lr.ptrs - 1, np.int32)); P.rcv_lids = _p(k(lr.data - 1, np.int32))
            P.buf_snd = _p(self._k(np.zeros(max(1, len(ls.data))))); P.buf_rcv = _p(self._k(np.zeros(max(1, len(lr.data)))))
            P.b = _p(k(b_vals[p], np.float64))
            P.x, P.r, P.c, P.u = _p(self.x[p]), _p(self.r[p]), _p(self.c[p]), _p(self.u[p])
        self.n = n

    def _k(self, a):
        self.keep.append(a)
        return a

    def cg(self, maxiter: int, tol: float = 0.0):
        hist = np.zeros(maxiter + 1, dtype=np.float64)
        tm = _Times()
        it = _lib.pa_oracle_cg(self.n, C.byref(self.arr), maxiter, tol, _p(hist), C.byref(tm))
        times = {f[0]: getattr(tm, f[0]) for f in _Times._fields_}
        return it, hist[: it + 1], times

    def time_spmv(self, reps: int) -> float:
        return float(_lib.pa_oracle_time_spmv(self.n, C.byref(self.arr), reps))
