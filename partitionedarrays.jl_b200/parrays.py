"""Host-side mirror of the reference's operator API for the hot path, on the CUDAArray backend.

Same names, argument meaning and error behaviour as the reference (file:line in /root/reference):
  with_cuda / CUDAArray      <-> with_debug / DebugArray (src/debug_array.jl:7-31), with_mpi / MPIArray (src/mpi_array.jl:42-117)
  PRange                     <-> src/p_range.jl:1776-1842
  PVector, pvector/pfill/pzeros/pones, own_values/ghost_values/local_values, consistent!, assemble!,
  dot, norm, sum, copy!, fill!, rmul!, broadcast updates            <-> src/p_vector.jl
  PSparseMatrix, psparse(…; assembled=true), mul!, fillstored!       <-> src/p_sparse_matrix.jl
Every data-touching call goes through the C ABI (include/pa_b200.h) into the CUDA library; this module
holds no arithmetic.  Julia's `f!` is spelled `f_` here (consistent_ = consistent!)."""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _capi
from ._capi import check, f64, i32, i64, ptr
from . import prange as pr


class CUDAArray:
    """The array-of-parts backend.  ``mode='sequential'``: every part lives in this process on one GPU
    (the DebugArray execution model); ``mode='distributed'``: one part per process / GPU, rank r owns
    part r+1 (the MPIArray model), peers mapped with CUDA IPC, scalars all-reduced with NCCL."""

    def __init__(self, nparts: int, mode: str = "sequential", device: Optional[int] = None, arena_bytes: int = 1 << 30,
                 stream: Optional[int] = None, group=None):
        L = _capi.lib()
        self.nparts, self.mode = int(nparts), mode
        self.group = group
        if mode == "sequential":
            self.parts = list(range(1, nparts + 1))
            self.rank, self.world = 0, 1
            device = 0 if device is None else device
        elif mode == "distributed":
            import torch.distributed as dist

            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
            if self.world != nparts:
                raise ValueError(f"distributed CUDAArray needs one process per part: {self.world} processes, {nparts} parts")
            self.parts = [self.rank + 1]
            device = int(os.environ.get("LOCAL_RANK", self.rank)) if device is None else device
        else:
            raise ValueError("mode must be 'sequential' or 'distributed'")
        self.device = device
        h = C.c_void_p()
        ids = i32(self.parts)
        check(L.pa_ctx_create(nparts, len(self.parts), ptr(ids), device, arena_bytes, stream, C.byref(h)))
        self.h = h
        if mode == "distributed" and self.world > 1:
            self._link_peers()

    @classmethod
    def _adopt(cls, handle, nparts: int, part: int, device: int, comm: "_ThreadComm"):
        """Backend object around one context of pa_ctx_create_multi (one process, one host thread per GPU)."""
        self = cls.__new__(cls)
        self.nparts, self.mode, self.group = int(nparts), "multi", None
        self.parts, self.rank, self.world, self.device = [part], part - 1, nparts, device
        self.h, self._comm = handle, comm
        return self

    # -- distributed plumbing (torch.distributed is only used to move handles around) --
    def _link_peers(self):
        import torch.distributed as dist

        L = _capi.lib()
        handle = np.zeros(64, dtype=np.uint8)
        check(L.pa_ctx_arena_export(self.h, 0, ptr(handle)))
        uid = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            check(L.pa_nccl_unique_id(ptr(uid)))
        objs = [None] * self.world
        dist.all_gather_object(objs, (handle.tobytes(), uid.tobytes()), group=self.group)
        for r, (hb, _) in enumerate(objs):
            if r != self.rank:
                hh = np.frombuffer(hb, dtype=np.uint8).copy()
                check(L.pa_ctx_arena_import(self.h, r + 1, ptr(hh)))
        uid0 = np.frombuffer(objs[0][1], dtype=np.uint8).copy()
        check(L.pa_ctx_nccl_init(self.h, ptr(uid0), self.rank, self.world))
        dist.barrier(group=self.group)

    def gather_all(self, local_objs: list) -> list:
        """Objects of all parts ordered by part id (setup-time metadata only)."""
        if self.mode == "sequential" or self.world == 1:
            return list(local_objs)
        if self.mode == "multi":
            return [o for per_rank in self._comm.all_gather(self.rank, local_objs) for o in per_rank]
        import torch.distributed as dist

        out = [None] * self.world
        dist.all_gather_object(out, local_objs, group=self.group)
        return [o for per_rank in out for o in per_rank]

    def allreduce_max(self, v: int) -> int:
        if self.mode == "sequential" or self.world == 1:
            return v
        return max(self.gather_all([v]))

    # -- reference-like helpers --
    def linear_indices(self):
        return list(self.parts)

    def i_am_main(self) -> bool:
        return 1 in self.parts

    def map(self, f: Callable, *arrays):
        """map over the local parts (host metadata only, like the reference's map on index arrays)."""
        return [f(*(a[k] for a in arrays)) for k in range(len(self.parts))]

    def sync(self):
        check(_capi.lib().pa_ctx_sync(self.h))

    def launch_count(self) -> int:
        out = C.c_int64()
        check(_capi.lib().pa_ctx_launch_count(self.h, C.byref(out)))
        return out.value

    def set_knob(self, key: str, value: int):
        check(_capi.lib().pa_ctx_set_knob(self.h, key.encode(), int(value)))

    def close(self):
        if getattr(self, "h", None):
            _capi.lib().pa_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def with_cuda(f: Callable, nparts: int = 1, **kw):
    """with_cuda(f) = f(distribute) — mirrors with_debug (src/debug_array.jl:7-9).  ``distribute(ranks)``
    returns the backend whose ``parts`` are the part ids held by this process."""
    backend = CUDAArray(nparts, **kw)
    try:
        return f(lambda ranks=None: backend)
    finally:
        backend.close()


class _ThreadComm:
    """All-gather of setup-time metadata between the host threads of one process (with_cuda_multi)."""

    def __init__(self, n: int):
        import threading

        self.n, self.slots, self.barrier = n, [None] * n, threading.Barrier(n)

    def all_gather(self, rank: int, obj):
        self.slots[rank] = obj
        self.barrier.wait()
        out = list(self.slots)
        self.barrier.wait()  # nobody overwrites a slot before everybody has read
        return out


def with_cuda_multi(f: Callable, devices: Sequence[int], arena_bytes: int = 1 << 30):
    """One PROCESS driving several GPUs: the DebugArray execution model (src/debug_array.jl) spanning the box.
    `f(backend)` runs once per device in its own host thread (SPMD, like one MPI rank per GPU, but threads of one process:
    ctypes releases the GIL inside every library call); the contexts are created and peer-linked by pa_ctx_create_multi.
    Returns the list of results, one per part."""
    import threading

    L = _capi.lib()
    n = len(devices)
    devs = i32(list(devices))
    handles = (C.c_void_p * n)()
    check(L.pa_ctx_create_multi(n, ptr(devs), arena_bytes, handles))
    comm = _ThreadComm(n)
    backends = [CUDAArray._adopt(C.c_void_p(handles[k]), n, k + 1, int(devices[k]), comm) for k in range(n)]
    results, errors = [None] * n, [None] * n

    def run(k):
        try:
            results[k] = f(backends[k])
        except BaseException as e:  # noqa: BLE001 - reported after the join
            errors[k] = e
            comm.barrier.abort()

    threads = [threading.Thread(target=run, args=(k,)) for k in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for b in backends:
        b.close()
    for e in errors:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in errors:
        if e is not None:
            raise e
    return results


class PRange:
    """PRange(index_partition) (src/p_range.jl:1776-1787) on the CUDA backend: the per-part
    AbstractLocalIndices plus the device-resident exchange plan (AssemblyCache, :354-387)."""

    def __init__(self, backend: CUDAArray, indices: List[pr.LocalIndices]):
        assert len(indices) == len(backend.parts)
        self.backend, self.indices = backend, indices
        self._plan = None
        self.plans: Optional[List[pr.PartPlan]] = None

    def partition(self):
        return self.indices

    def __len__(self):
        return self.indices[0].n_global

    @property
    def plan(self):
        if self._plan is None:
            L = _capi.lib()
            b = self.backend
            self.plans = pr.build_plans(self.indices, b.gather_all)
            sym = b.allreduce_max(max(ind.n_local for ind in self.indices))
            h = C.c_void_p()
            check(L.pa_plan_create(b.h, C.byref(h)))
            for k, (ind, p) in enumerate(zip(self.indices, self.plans)):
                o2l = None if ind.own_is_prefix else i32(ind.own_to_local)
                g2l = None if ind.own_is_prefix else i32(ind.ghost_to_local)
                arrs = [i32(a) for a in (p.nbr_snd, p.snd_ptrs, p.snd_lids, p.snd_remote_lids, p.nbr_rcv, p.rcv_ptrs, p.rcv_lids, p.rcv_remote_lids)]
                check(L.pa_plan_set_part(h, k, ind.n_local, ind.n_own, ptr(o2l), ptr(g2l), len(p.nbr_snd), ptr(arrs[0]), ptr(arrs[1]),
                                         ptr(arrs[2]), ptr(arrs[3]), len(p.nbr_rcv), ptr(arrs[4]), ptr(arrs[5]), ptr(arrs[6]), ptr(arrs[7])))
            check(L.pa_plan_commit(h, sym))
            self._plan = h
        return self._plan

    def __del__(self):
        try:
            if self._plan is not None and self.backend.h:
                _capi.lib().pa_plan_destroy(self._plan)
        except Exception:
            pass


def uniform_partition(backend: CUDAArray, np_, n, ghost=None, periodic=None) -> PRange:
    """uniform_partition(ranks,np,n[,ghost,periodic]) (src/p_range.jl:585-599)."""
    if isinstance(np_, int):
        np_, n = (np_,), (n,)
        ghost = None if ghost is None else (ghost,) if isinstance(ghost, bool) else ghost
        periodic = None if periodic is None else (periodic,) if isinstance(periodic, bool) else periodic
    assert int(np.prod(np_)) == backend.nparts
    return PRange(backend, [pr.uniform_partition_part(p, np_, n, ghost, periodic) for p in backend.parts])


def variable_partition(backend: CUDAArray, n_own_all: Sequence[int], n_global: int) -> PRange:
    return PRange(backend, [pr.variable_partition_part(p, n_own_all, n_global) for p in backend.parts])


class PVector:
    """PVector (src/p_vector.jl:324-345): one device-resident local array per part (own + ghost entries in the
    reference's local order)."""

    def __init__(self, rows: PRange):
        self.rows = rows
        h = C.c_void_p()
        check(_capi.lib().pa_vec_create(rows.plan, C.byref(h)))
        self.h = h

    # --- construction / transfer ---
    @property
    def backend(self):
        return self.rows.backend

    def set_local_values(self, values: List[np.ndarray]):
        L = _capi.lib()
        for k, v in enumerate(values):
            v = f64(v)
            check(L.pa_vec_upload(self.h, k, ptr(v), len(v)))
        self.backend.sync()
        return self

    def local_values(self) -> List[np.ndarray]:
        """local_values(v) (src/p_vector.jl:350-373) copied to the host."""
        L = _capi.lib()
        out = []
        for k, ind in enumerate(self.rows.indices):
            a = np.zeros(ind.n_local, dtype=np.float64)
            check(L.pa_vec_download(self.h, k, ptr(a), len(a)))
            out.append(a)
        return out

    def own_values(self) -> List[np.ndarray]:
        return [v[ind.own_to_local - 1] for v, ind in zip(self.local_values(), self.rows.indices)]

    def ghost_values(self) -> List[np.ndarray]:
        return [v[ind.ghost_to_local - 1] for v, ind in zip(self.local_values(), self.rows.indices)]

    def similar(self) -> "PVector":
        return PVector(self.rows)

    def copy(self) -> "PVector":
        return PVector(self.rows).copy_(self)

    # --- in-place operations (Julia's f! -> f_) ---
    def fill_(self, a: float):
        check(_capi.lib().pa_vec_fill(self.h, float(a)))
        return self

    def copy_(self, src: "PVector"):
        check(_capi.lib().pa_vec_copy(self.h, src.h))
        return self

    def rmul_(self, a: float):
        check(_capi.lib().pa_vec_scale(self.h, float(a)))
        return self

    def axpby_(self, a: float, x: "PVector", b: float):
        """self .= a.*x .+ b.*self"""
        check(_capi.lib().pa_vec_axpby(self.h, float(a), x.h, float(b)))
        return self

    def waxpby_(self, a: float, x: "PVector", b: float, y: "PVector"):
        """self .= a.*x .+ b.*y"""
        check(_capi.lib().pa_vec_waxpby(self.h, float(a), x.h, float(b), y.h))
        return self

    def consistent_(self):
        """consistent!(v) (src/p_vector.jl:747-755); returns a waitable like the reference's task."""
        check(_capi.lib().pa_vec_consistent(self.h))
        return _Task(self)

    def assemble_(self, op: str = "+"):
        """assemble!([op,] v) (src/p_vector.jl:695-708): op in '+', 'insert', 'max', 'min'."""
        code = {"+": _capi.PA_OP_SUM, "insert": _capi.PA_OP_INSERT, "max": _capi.PA_OP_MAX, "min": _capi.PA_OP_MIN}.get(op)
        if code is None:
            raise ValueError(f"assemble!: unsupported operation {op!r}")
        check(_capi.lib().pa_vec_assemble_op(self.h, code))
        return _Task(self)

    # --- reductions ---
    def dot(self, other: "PVector") -> float:
        out = C.c_double()
        check(_capi.lib().pa_vec_dot(self.h, other.h, C.byref(out)))
        return out.value

    def norm(self, p: float = 2) -> float:
        """norm(a, p) = (sum over parts of norm(own, p)^p)^(1/p) (src/p_vector.jl:1201-1206)."""
        if p == 2:
            out = C.c_double()
            check(_capi.lib().pa_vec_norm2(self.h, C.byref(out)))
            return math.sqrt(out.value)
        if not (p >= 1 and math.isfinite(p)):
            raise ValueError("norm(a, p): p must be a finite number >= 1")
        return self.reduce("+", _op=_capi.PA_OP_ABSSUM if p == 1 else _capi.PA_OP_ABSPOW, _p=float(p)) ** (1.0 / p)

    def reduce(self, op: str = "+", _op: Optional[int] = None, _p: float = 0.0) -> float:
        """reduce(op, a) (src/p_vector.jl:1178-1183): per-part reduction of the own values on the device (neutral element
        as init, :1170-1175), then the reduction over parts (over processes: a host all-gather of one number per part)."""
        code = _op if _op is not None else {"+": _capi.PA_OP_SUM, "max": _capi.PA_OP_MAX, "min": _capi.PA_OP_MIN}.get(op)
        if code is None:
            raise ValueError(f"reduce: unsupported operation {op!r}")
        nl = len(self.backend.parts)
        part = np.zeros(nl, dtype=np.float64)
        check(_capi.lib().pa_vec_reduce_parts(self.h, code, float(_p), ptr(part)))
        allp = self.backend.gather_all([float(v) for v in part])
        if op == "max":
            return max(allp)
        if op == "min":
            return min(allp)
        acc = 0.0
        for v in allp:  # part order, like sum over the array of parts
            acc += v
        return acc

    def maximum(self) -> float:
        return self.reduce("max")

    def minimum(self) -> float:
        return self.reduce("min")

    def sum(self) -> float:
        out = C.c_double()
        check(_capi.lib().pa_vec_sum(self.h, C.byref(out)))
        return out.value

    def collect(self) -> np.ndarray:
        """collect(v): the global vector (gathered through the host; debugging / tests)."""
        own = self.own_values()
        pieces = self.backend.gather_all([(ind.own_to_global, o) for ind, o in zip(self.rows.indices, own)])
        out = np.zeros(len(self.rows), dtype=np.float64)
        for g, o in pieces:
            out[g - 1] = o
        return out

    def free(self):
        if getattr(self, "h", None) and self.backend.h:
            _capi.lib().pa_vec_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _Task:
    def __init__(self, obj):
        self.obj = obj

    def wait(self):
        self.obj.backend.sync()
        return self.obj

    fetch = wait


def pvector(f: Callable, rows: PRange) -> PVector:
    """pvector(f, index_partition): f(indices) -> local values (src/p_vector.jl:832-846)."""
    return PVector(rows).set_local_values([np.asarray(f(ind), dtype=np.float64) for ind in rows.indices])


def pfill(a: float, rows: PRange) -> PVector:
    return PVector(rows).fill_(a)


def pzeros(rows: PRange) -> PVector:
    return pfill(0.0, rows)


def pones(rows: PRange) -> PVector:
    return pfill(1.0, rows)


def pvector_from_global(xg: np.ndarray, rows: PRange, ghosts: bool = False) -> PVector:
    vals = []
    for ind in rows.indices:
        v = np.zeros(ind.n_local)
        v[ind.own_to_local - 1] = xg[ind.own_to_global - 1]
        if ghosts:
            v[ind.ghost_to_local - 1] = xg[ind.ghost_to_global - 1]
        vals.append(v)
    return PVector(rows).set_local_values(vals)


def dot(a: PVector, b: PVector) -> float:
    return a.dot(b)


def norm(a: PVector, p: float = 2) -> float:
    return a.norm(p)


def consistent_(v: PVector):
    return v.consistent_()


def assemble_(v: PVector, op: str = "+"):
    return v.assemble_(op)


class ExchangeGraph:
    """ExchangeGraph(snd[, rcv]) (src/primitives.jl:728-741, 776-789): ``snd[k]`` = 1-based ids of the parts local part k
    sends to.  Without ``rcv`` the receive sides are found like default_find_rcv_ids (:826-859): the transpose of the
    adjacency, sources sorted ascending (one host all-gather of the send lists)."""

    def __init__(self, backend: CUDAArray, snd: Sequence[Sequence[int]], rcv: Optional[Sequence[Sequence[int]]] = None):
        self.backend = backend
        self.snd = [[int(q) for q in s] for s in snd]
        if len(self.snd) != len(backend.parts):
            raise ValueError("ExchangeGraph: one send list per local part")
        if rcv is None:
            all_snd = backend.gather_all(self.snd)
            rcv = [sorted(p + 1 for p, s in enumerate(all_snd) if me in set(s)) for me in backend.parts]
        self.rcv = [[int(q) for q in r] for r in rcv]


def _ptrs1(lengths) -> np.ndarray:
    out = np.ones(len(lengths) + 1, dtype=np.int64)
    np.cumsum(lengths, out=out[1:])
    out[1:] += 1
    return out


def exchange_layout(graph: ExchangeGraph, snd_len: Sequence[Sequence[int]]):
    """allocate_exchange (src/primitives.jl:921-947) on the host: from the send lengths of every local part, the receive
    lengths (= what each source sends here) and, for every source, where the segment addressed to this part starts inside
    the SOURCE's send buffer (the address a receiver pulls from).  One all-gather of the send ids and lengths.
    Returns (rcv_len, rcv_src_off, sym_snd_len) with one list per local part."""
    b = graph.backend
    all_snd_ids, all_snd_len = b.gather_all(graph.snd), b.gather_all([list(map(int, l)) for l in snd_len])
    rcv_len, rcv_off = [], []
    for k, me in enumerate(b.parts):
        rl, off = [], []
        for src in graph.rcv[k]:
            ids = list(all_snd_ids[src - 1])
            if me not in ids:
                raise ValueError(f"exchange: graph not consistent (part {me} receives from {src}, which does not send to it)")
            j = ids.index(me)
            rl.append(int(all_snd_len[src - 1][j]))
            off.append(int(sum(all_snd_len[src - 1][:j])))
        rcv_len.append(rl)
        rcv_off.append(off)
    return rcv_len, rcv_off, max([sum(l) for l in all_snd_len] + [0])


def exchange(snd: Sequence[Sequence], graph: ExchangeGraph) -> List[List[np.ndarray]]:
    """rcv = fetch(exchange(snd, graph)) (src/primitives.jl:876-925, 992-1042): ``snd[k][j]`` (a scalar or a vector of
    Float64 / Int64 / Int32 / Float32 / ...) goes to part ``graph.snd[k][j]``; ``rcv[k][i]`` is what part ``graph.rcv[k][i]`` sent here.  The
    payload moves on the device: each receiver reads its segments from the senders' HBM (pa_xchg_*)."""
    b = graph.backend
    L = _capi.lib()
    nl = len(b.parts)
    segs = [[np.atleast_1d(np.asarray(v)) for v in s] for s in snd]
    # element type on the device = the common type of everything any part sends (all processes must agree): Int32 / Float32
    # payloads (the reference's index lists are JaggedArray{Int32,Int32}) travel as 4-byte elements, Int16 / Int8 / Bool as
    # 2 / 1 bytes; everything else as Float64 when any part sends floats, Int64 otherwise
    names = sorted({a.dtype.str for s in segs for a in s})
    every = sorted({n for part in b.gather_all([names]) for n in part})
    dtype = np.result_type(*[np.dtype(n) for n in every]) if every else np.dtype(np.int64)
    if dtype.kind == "b":
        dtype = np.dtype(np.uint8)
    if dtype.kind not in "iuf" or dtype.itemsize not in (1, 2, 4, 8) or dtype == np.float16:
        dtype = np.dtype(np.float64) if dtype.kind == "f" else np.dtype(np.int64)
    for k in range(nl):
        if len(segs[k]) != len(graph.snd[k]):
            raise ValueError("exchange: one send item per destination")
    snd_len = [[len(a) for a in s] for s in segs]
    rcv_len, rcv_off, sym = exchange_layout(graph, snd_len)
    h = C.c_void_p()
    check(L.pa_xchg_create(b.h, C.byref(h)))
    try:
        check(L.pa_xchg_set_elem_size(h, dtype.itemsize))
        for k in range(nl):
            si, ri = i32(graph.snd[k]), i32(graph.rcv[k])
            sp, rp, so = _ptrs1(snd_len[k]), _ptrs1(rcv_len[k]), i64(rcv_off[k])
            check(L.pa_xchg_set_part(h, k, len(si), ptr(si), ptr(sp), len(ri), ptr(ri), ptr(rp), ptr(so)))
        check(L.pa_xchg_commit(h, sym))
        for k in range(nl):
            flat = np.ascontiguousarray(np.concatenate([a.astype(dtype) for a in segs[k]]) if segs[k] else np.zeros(0, dtype=dtype))
            check(L.pa_xchg_upload_snd(h, k, ptr(flat), len(flat)))
        check(L.pa_xchg_exchange(h))
        out = []
        for k in range(nl):
            flat = np.zeros(sum(rcv_len[k]), dtype=dtype)
            check(L.pa_xchg_download_rcv(h, k, ptr(flat), len(flat)))
            cuts = np.cumsum([0] + rcv_len[k])
            out.append([flat[cuts[i] : cuts[i + 1]].copy() for i in range(len(rcv_len[k]))])
    finally:
        L.pa_xchg_destroy(h)
    return out


class PSparseMatrix:
    """PSparseMatrix (src/p_sparse_matrix.jl:971-991), assembled, per-part CSR on the device."""

    def __init__(self, rows: PRange, cols: PRange):
        self.rows, self.cols = rows, cols
        h = C.c_void_p()
        check(_capi.lib().pa_mat_create(rows.plan, cols.plan, C.byref(h)))
        self.h = h
        self.assembled = True

    @property
    def backend(self):
        return self.rows.backend

    def axes(self, d: int) -> PRange:
        return self.rows if d == 1 else self.cols

    def set_csr(self, k: int, rowptr, colval, nzval, index_base: int = 1):
        """Unsplit local matrix: own rows x local columns (SparseMatrixCSR{Bi,Float64,Ti})."""
        rowptr, colval = np.ascontiguousarray(rowptr), np.ascontiguousarray(colval)
        nz = f64(nzval)
        pb, cb = rowptr.dtype.itemsize * 8, (colval.dtype.itemsize * 8 if len(colval) else 32)
        check(_capi.lib().pa_mat_set_csr(self.h, k, len(rowptr) - 1, self.cols.indices[k].n_local, index_base, pb, cb, ptr(rowptr),
                                         ptr(colval), ptr(nz)))

    def set_csr_split(self, k: int, oo, oh, index_base: int = 1):
        """Split format: oo/oh = (rowptr, colval, nzval) of the own_own and own_ghost blocks."""
        rp1, cv1, nz1 = (np.ascontiguousarray(a) for a in oo)
        rp2, cv2, nz2 = (np.ascontiguousarray(a) for a in oh)
        assert rp1.dtype == rp2.dtype
        cv2 = cv2.astype(cv1.dtype) if len(cv1) else cv2
        cv1 = cv1.astype(cv2.dtype)
        check(_capi.lib().pa_mat_set_csr_split(self.h, k, len(rp1) - 1, index_base, rp1.dtype.itemsize * 8, cv1.dtype.itemsize * 8,
                                               ptr(rp1), ptr(cv1), ptr(f64(nz1)), ptr(rp2), ptr(cv2), ptr(f64(nz2))))

    def set_csc(self, k: int, nrows: int, colptr, rowval, nzval, index_base: int = 1):
        """Unsplit SparseMatrixCSC local matrix (n_local_rows x n_local_cols; the reference's default storage)."""
        colptr, rowval = np.ascontiguousarray(colptr), np.ascontiguousarray(rowval)
        check(_capi.lib().pa_mat_set_csc(self.h, k, nrows, len(colptr) - 1, index_base, colptr.dtype.itemsize * 8,
                                         rowval.dtype.itemsize * 8 if len(rowval) else 32, ptr(colptr), ptr(rowval), ptr(f64(nzval))))

    def set_csc_split(self, k: int, nrows: int, oo, oh, index_base: int = 1):
        """Split format with SparseMatrixCSC blocks: oo/oh = (colptr, rowval, nzval)."""
        cp1, rv1, nz1 = (np.ascontiguousarray(a) for a in oo)
        cp2, rv2, nz2 = (np.ascontiguousarray(a) for a in oh)
        rv2 = rv2.astype(rv1.dtype)
        cp2 = cp2.astype(cp1.dtype)
        check(_capi.lib().pa_mat_set_csc_split(self.h, k, nrows, index_base, cp1.dtype.itemsize * 8, rv1.dtype.itemsize * 8,
                                               ptr(cp1), ptr(rv1), ptr(f64(nz1)), ptr(cp2), ptr(rv2), ptr(f64(nz2))))

    def set_coo(self, k: int, I_own, J_local, V):
        """Device-side sparse_matrix(I,J,V; reuse=true): 1-based own row ids / local column ids, ids < 1 skipped."""
        Ia, Ja, Va = i64(I_own), i64(J_local), f64(V)
        check(_capi.lib().pa_mat_set_coo(self.h, k, len(Ia), 64, ptr(Ia), ptr(Ja), ptr(Va)))
        self._coo_values = getattr(self, "_coo_values", {})
        self._coo_values[k] = Va

    def coo_values(self, k: int) -> np.ndarray:
        """The value list V of the COO pattern part k was compressed from (what update_coo_values_ expects a new version of)."""
        return self._coo_values[k]

    def update_coo_values_(self, V: List):
        """psparse!(A, V, cache): refresh the values of the same COO pattern (src/p_sparse_matrix.jl:1291-1305)."""
        for k, v in enumerate(V):
            v = f64(v)
            check(_capi.lib().pa_mat_update_coo_values(self.h, k, ptr(v), len(v)))
        return self

    def commit(self):
        check(_capi.lib().pa_mat_commit(self.h))
        return self

    def nnz(self, k: int = 0) -> int:
        out = C.c_int64()
        check(_capi.lib().pa_mat_nnz(self.h, k, C.byref(out)))
        return out.value

    def download_csr(self, k: int):
        nnz, nrows = self.nnz(k), self.rows.indices[k].n_own
        rp, cv, nz = np.zeros(nrows + 1, np.int64), np.zeros(nnz, np.int32), np.zeros(nnz, np.float64)
        check(_capi.lib().pa_mat_download_csr(self.h, k, ptr(rp), ptr(cv), ptr(nz)))
        return rp, cv, nz

    def download_csr_all(self, k: int):
        """All stored rows of part k (own rows, and the ghost rows of a sub-assembled matrix): 0-based rowptr / colval."""
        nnz = self.nnz(k)
        nrows = C.c_int64()
        check(_capi.lib().pa_mat_nrows(self.h, k, C.byref(nrows)))
        rp, cv, nz = np.zeros(nrows.value + 1, np.int64), np.zeros(nnz, np.int32), np.zeros(nnz, np.float64)
        check(_capi.lib().pa_mat_download_csr(self.h, k, ptr(rp), ptr(cv), ptr(nz)))
        return rp, cv, nz

    def fillstored_(self, a: float):
        check(_capi.lib().pa_mat_fill_stored(self.h, float(a)))
        return self

    def free(self):
        if getattr(self, "h", None) and self.backend.h:
            _capi.lib().pa_mat_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _local_spmv(backend: CUDAArray, kind: int, ptr_, idx, val, x, nb: int, index_base: int) -> np.ndarray:
    ptr_, idx, val, x = (np.ascontiguousarray(a) for a in (ptr_, idx, val, x))
    if ptr_.dtype != idx.dtype or ptr_.dtype not in (np.int32, np.int64):
        raise TypeError("spmv!: ptr and idx must both be Int32 or Int64")
    if val.dtype != x.dtype or val.dtype not in (np.float32, np.float64):
        raise TypeError("spmv!: values and x must both be Float32 or Float64")
    out = np.empty(nb, dtype=val.dtype)
    check(_capi.lib().pa_local_spmv(backend.h, kind, index_base, ptr_.dtype.itemsize * 8, val.dtype.itemsize * 8, len(ptr_) - 1, nb,
                                    ptr(ptr_), ptr(idx), ptr(val), ptr(x), len(x), ptr(out)))
    return out


def spmv_(backend: CUDAArray, fmt: str, ptr_, idx, val, x, m: int, n: int, index_base: int = 1) -> np.ndarray:
    """b = spmv!(b, A, x) for one local m x n matrix (src/sparse_utils.jl:617-640): fmt 'csr' (ptr_=rowptr, idx=colval) or
    'csc' (ptr_=colptr, idx=rowval); Float64/Float32 values, Int32/Int64 indices, index_base 1 (Julia) or 0 (CSR{0})."""
    assert len(x) == n and len(ptr_) - 1 == (m if fmt == "csr" else n)
    return _local_spmv(backend, 0 if fmt == "csr" else 1, ptr_, idx, val, x, m, index_base)


def spmtv_(backend: CUDAArray, fmt: str, ptr_, idx, val, x, m: int, n: int, index_base: int = 1) -> np.ndarray:
    """b = spmtv!(b, A, x) = transpose(A)*x for one local m x n matrix (src/sparse_utils.jl:626-647)."""
    assert len(x) == m and len(ptr_) - 1 == (m if fmt == "csr" else n)
    return _local_spmv(backend, 1 if fmt == "csr" else 0, ptr_, idx, val, x, n, index_base)


def mul_(c: PVector, A: PSparseMatrix, b: PVector, alpha: float = 1.0, beta: float = 0.0, flags: int = 0) -> PVector:
    """mul!(c,A,b[,alpha,beta]) (src/p_sparse_matrix.jl:2090-2142)."""
    check(_capi.lib().pa_spmv(A.h, b.h, c.h, float(alpha), float(beta), flags))
    return c


def mul_transpose_(c: PVector, A: PSparseMatrix, b: PVector, alpha: float = 1.0, beta: float = 0.0) -> PVector:
    """mul!(c, transpose(A), b[, alpha, beta]) (src/p_sparse_matrix.jl:2144-2162); c lives on axes(A,2), b on axes(A,1)."""
    check(_capi.lib().pa_spmv_transpose(A.h, b.h, c.h, float(alpha), float(beta)))
    return c


def mul_no_lat_(c: PVector, A: PSparseMatrix, b: PVector) -> PVector:
    """HPCG mul_no_lat! (HPCG/src/hpcg_utils.jl:6-17)."""
    return mul_(c, A, b)


def _coo_to_csr(I, J, V, m, n):
    """Host COO -> CSR{1} (sorted columns, duplicates summed in input order, ids<1 -> stored (1,1,0));
    setup-time helper standing in for sparse_matrix/compresscoo (src/sparse_utils.jl:313-405)."""
    I = np.asarray(I, dtype=np.int64).copy(); J = np.asarray(J, dtype=np.int64).copy(); V = np.asarray(V, dtype=np.float64).copy()
    bad = (I < 1) | (J < 1)
    if m * n == 0:
        I, J, V = I[:0], J[:0], V[:0]
    else:
        I[bad], J[bad], V[bad] = 1, 1, 0.0
    order = np.lexsort((J, I))
    I, J, V = I[order], J[order], V[order]
    if len(I):
        new = np.ones(len(I), dtype=bool)
        new[1:] = (I[1:] != I[:-1]) | (J[1:] != J[:-1])
        pos = np.cumsum(new) - 1
        nz = np.zeros(int(pos[-1]) + 1)
        np.add.at(nz, pos, V)
        I, J = I[new], J[new]
    else:
        nz = V
    rowptr = np.ones(m + 1, dtype=np.int64)
    rowptr[1:] = 1 + np.cumsum(np.bincount(I - 1, minlength=m)) if len(I) else 1
    return rowptr.astype(np.int32), J.astype(np.int32), nz


def _csr_to_csc(rp, cv, nz, ncols):
    """1-based CSR (rows sorted by column) -> 1-based CSC (what SparseArrays.sparse builds; setup-time host helper)."""
    m = len(rp) - 1
    rowid = np.repeat(np.arange(1, m + 1), np.diff(rp.astype(np.int64)))
    order = np.lexsort((rowid, cv))
    colptr = np.ones(ncols + 1, dtype=np.int64)
    colptr[1:] = 1 + np.cumsum(np.bincount(cv.astype(np.int64) - 1, minlength=ncols)) if len(cv) else 1
    return colptr, rowid[order].astype(np.int64), nz[order]


def psparse(I: List, J: List, V: List, rows: PRange, cols: PRange, assembled: bool = True, split_format: bool = True,
            local_format: str = "csr", compress: str = "host", ship: str = "host", assemble: bool = True) -> PSparseMatrix:
    """psparse([T,] I,J,V,row_partition,col_partition; assembled=true) (src/p_sparse_matrix.jl:1150-1286).
    local_format: "csr" = SparseMatrixCSR{1,Float64,Int32} local matrices; "csc" = the reference's default
    SparseMatrixCSC{Float64,Int} (converted to CSR at upload, summation order of spmv_csc! preserved).
    Per local part: COO in global ids.  assembled=True: every row is owned by the part that lists it (what
    build_p_matrix / fdm_example use, :1249-1270).  assembled=False (the reference's default): rows owned elsewhere are
    shipped to their owner first.  Ghost columns are discovered with find_owner + union_ghost (:1226-1236).
    The COO->CSR compression runs on the host at setup time (on-device compression is 'next')."""
    b = rows.backend
    if not assembled and not assemble:
        return _psparse_subassembled(I, J, V, rows, cols, local_format)
    if not assembled:
        return _psparse_disassembled(I, J, V, rows, cols, split_format, local_format, compress, ship)
    new_cols = []
    for ind_c, j in zip(cols.indices, J):
        j = np.asarray(j, dtype=np.int64)
        owners = _find_owner(cols, ind_c, j)
        new_cols.append(pr.union_ghost(ind_c, j, owners))
    colr = PRange(b, new_cols)
    A = PSparseMatrix(rows, colr)
    for k, (ind_r, ind_c) in enumerate(zip(rows.indices, colr.indices)):
        Ik, Jk = np.asarray(I[k], dtype=np.int64), np.asarray(J[k], dtype=np.int64)
        li = ind_r.global_to_local(Ik).astype(np.int64)
        lj = ind_c.global_to_local(Jk).astype(np.int64)
        li[Ik < 1] = 0
        lj[Jk < 1] = 0
        if np.any((li == 0) & (Ik >= 1)):
            raise ValueError("psparse(assembled=true): a row id is not local to its part")
        if compress == "device":
            # COO -> CSR on the GPU (sort + in-order duplicate sums), pattern cache kept for update_coo_values_ (psparse!)
            if not (ind_r.own_is_prefix and ind_c.own_is_prefix):
                raise ValueError("device compression needs own-first layouts")
            A.set_coo(k, li, lj, V[k])
            continue
        rp, cv, nz = _coo_to_csr(li, lj, V[k], ind_r.n_local, ind_c.n_local)
        # keep own rows only, in own order (ghost rows of an assembled matrix are empty)
        o2l = ind_r.own_to_local.astype(np.int64)
        starts, ends = rp[o2l - 1].astype(np.int64), rp[o2l].astype(np.int64)
        lens = ends - starts
        rp2 = np.ones(len(o2l) + 1, dtype=np.int64)
        rp2[1:] = 1 + np.cumsum(lens)
        take = np.concatenate([np.arange(s - 1, e - 1) for s, e in zip(starts, ends)]) if len(o2l) and lens.sum() else np.zeros(0, dtype=np.int64)
        cv2, nz2 = cv[take], nz[take]
        if split_format:
            # own_own / own_ghost blocks in own / ghost numbering (src/p_sparse_matrix.jl:823-899)
            l2o = np.zeros(ind_c.n_local + 1, dtype=np.int64); l2o[ind_c.own_to_local] = np.arange(1, ind_c.n_own + 1)
            l2g = np.zeros(ind_c.n_local + 1, dtype=np.int64); l2g[ind_c.ghost_to_local] = np.arange(1, ind_c.n_ghost + 1)
            rowid = np.repeat(np.arange(len(o2l)), lens)
            isown = l2o[cv2] > 0
            def block(mask, ids):
                cnt = np.bincount(rowid[mask], minlength=len(o2l))
                p = np.ones(len(o2l) + 1, dtype=np.int32); p[1:] = 1 + np.cumsum(cnt)
                return p, ids[mask].astype(np.int32), nz2[mask]
            oo, oh = block(isown, l2o[cv2]), block(~isown, l2g[cv2])
            if local_format == "csc":
                A.set_csc_split(k, len(o2l), _csr_to_csc(*oo, ind_c.n_own), _csr_to_csc(*oh, ind_c.n_ghost))
            else:
                A.set_csr_split(k, oo, oh)
        elif local_format == "csc":
            if not ind_r.own_is_prefix:
                raise ValueError("unsplit CSC upload needs own rows first")
            A.set_csc(k, ind_r.n_local, *_csr_to_csc(rp, cv, nz, ind_c.n_local))
        else:
            A.set_csr(k, rp2.astype(np.int32), cv2, nz2)
    return A.commit()


_DEVICE_SORT_MIN = 1 << 18


def _stable_order(primary: np.ndarray, secondary: np.ndarray, backend: Optional[CUDAArray] = None) -> np.ndarray:
    """np.lexsort((secondary, primary)): the stable order by (primary, secondary).  Long arrays of non-negative 32-bit ids are
    sorted on the device (pa_sort_perm_u64: one radix sort of packed 64-bit keys, same permutation); short ones on the host."""
    n = len(primary)
    if backend is None or n < _DEVICE_SORT_MIN or n >= (1 << 31):
        return np.lexsort((secondary, primary))
    p64, s64 = np.asarray(primary, dtype=np.int64), np.asarray(secondary, dtype=np.int64)
    if p64.min() < 0 or s64.min() < 0 or p64.max() >= (1 << 31) or s64.max() >= (1 << 32):
        return np.lexsort((secondary, primary))
    keys = np.ascontiguousarray((p64.astype(np.uint64) << np.uint64(32)) | s64.astype(np.uint64))
    perm = np.empty(n, dtype=np.int32)
    check(_capi.lib().pa_sort_perm_u64(backend.h, ptr(keys), n, ptr(perm)))
    return perm


def _stored_entries(li, lj, v, m, n, fmt, backend: Optional[CUDAArray] = None):
    """The stored entries of sparse_matrix(I,J,V,m,n) in the storage order of the local matrix type ("csc":
    column-major, "csr": row-major): duplicates added in input order, ids < 1 -> a stored (1,1,0.0)."""
    li, lj, v = li.copy(), lj.copy(), v.copy()
    if m * n == 0:
        li, lj, v = li[:0], lj[:0], v[:0]
    bad = (li < 1) | (lj < 1)
    li[bad], lj[bad], v[bad] = 1, 1, 0.0
    order = _stable_order(li, lj, backend) if fmt == "csr" else _stable_order(lj, li, backend)
    li, lj, v = li[order], lj[order], v[order]
    if len(li) == 0:
        return li, lj, v
    new = np.ones(len(li), dtype=bool)
    new[1:] = (li[1:] != li[:-1]) | (lj[1:] != lj[:-1])
    # (bincount adds the weights of a bin in input order, like the sequential np.add.at it replaces: same bits, 1.5x faster)
    nz = np.bincount(np.cumsum(new) - 1, weights=v, minlength=int(np.count_nonzero(new))).astype(np.float64)
    return li[new], lj[new], nz


def _ship_ghost_rows(b: CUDAArray, outgoing, ship: str):
    """The entries of ghost rows travel to the row owners (assemble, src/p_sparse_matrix.jl:1590-1756): ``outgoing[k]`` =
    (my part id, {destination part: (gi, gj, v)}).  Returns per local part the received (gi, gj, v), sender by sender in
    ascending part order.  ship='host': the metadata channel (an all-gather of everything, setup-time only);
    ship='device': three exchange! calls (I, J as Int64, V as Float64) -- every owner pulls exactly its segments from the
    senders' HBM (pa_xchg_*), nothing is broadcast."""
    if ship == "host":
        everyone = dict(b.gather_all(outgoing))
        out = []
        for me, _ in outgoing:
            rcv = [everyone[q][me] for q in sorted(everyone) if q != me and me in everyone[q]]
            out.append(tuple(np.concatenate([x[t] for x in rcv]) if rcv else np.zeros(0, np.float64 if t == 2 else np.int64) for t in range(3)))
        return out
    if ship != "device":
        raise ValueError("ship must be 'host' or 'device'")
    dests = [sorted(dst.keys()) for _, dst in outgoing]
    graph = ExchangeGraph(b, dests)  # receive sides: sources in ascending order (default_find_rcv_ids)
    parts = []
    for t, dt in ((0, np.int64), (1, np.int64), (2, np.float64)):
        rcv = exchange([[np.asarray(dst[q][t], dtype=dt) for q in d] for (_, dst), d in zip(outgoing, dests)], graph)
        parts.append([np.concatenate(r).astype(dt) if r else np.zeros(0, dt) for r in rcv])
    return [tuple(parts[t][k] for t in range(3)) for k in range(len(outgoing))]


def _psparse_subassembled(I, J, V, rows: PRange, cols: PRange, local_format: str) -> PSparseMatrix:
    """psparse(I,J,V,rows,cols; assemble=false) (src/p_sparse_matrix.jl:1186-1222): the SUB-ASSEMBLED matrix — every part
    compresses its own triplets over rows_sa x cols_sa = union_ghost(rows/cols, I/J) and keeps the rows it does not own as
    ghost rows (blocks ghost_own / ghost_ghost).  mul!(c,A,b) then multiplies own and ghost rows and finishes with
    assemble!(c) (:2109-2142); c lives on axes(A,1) (ghost rows included)."""
    b = rows.backend
    rsa_all, csa_all, mats = [], [], []
    for ind_r, ind_c, i, j, v in zip(rows.indices, cols.indices, I, J, V):
        i, j, v = np.asarray(i, dtype=np.int64), np.asarray(j, dtype=np.int64), np.asarray(v, dtype=np.float64)
        rsa = pr.union_ghost(ind_r, i, _find_owner(rows, ind_r, i))
        csa = pr.union_ghost(ind_c, j, _find_owner(cols, ind_c, j))
        if not (rsa.own_is_prefix and csa.own_is_prefix):
            raise ValueError("psparse(assemble=false): needs own-first local orders")
        li, lj = rsa.global_to_local(i).astype(np.int64), csa.global_to_local(j).astype(np.int64)
        li[i < 1] = 0
        lj[j < 1] = 0
        ei, ej, ev = _stored_entries(li, lj, v, rsa.n_local, csa.n_local, local_format, b)
        # rows sorted by local column id = own-block entries first, then the ghost block: the order of mul! (:2119-2139)
        mats.append(_coo_to_csr(ei, ej, ev, rsa.n_local, csa.n_local))
        rsa_all.append(rsa)
        csa_all.append(csa)
    A = PSparseMatrix(PRange(b, rsa_all), PRange(b, csa_all))
    for k, m in enumerate(mats):
        A.set_csr(k, *m)
    A.commit()
    A.assembled = False
    return A


def _psparse_disassembled(I, J, V, rows: PRange, cols: PRange, split_format: bool, local_format: str, compress: str, ship: str = "host") -> PSparseMatrix:
    """Disassembled input, the reference's default (src/p_sparse_matrix.jl:1186-1222, then split_format :823-899 and
    assemble :1590-1756), reproduced stage by stage because the stages fix both how the sums associate and how the ghost
    columns are numbered (= the order of the terms in every row of the ghost block):
      each part compresses ITS triplets first (duplicates in input order) into a sub-assembled local matrix over
      rows_sa/cols_sa = union_ghost(rows/cols, I/J); the stored entries of its ghost rows -- ghost_own block, then
      ghost_ghost block, each in storage order -- go to the row owners; an owner appends what it receives, sender by
      sender in ascending part order, to the stored entries of its own rows (own_own / own_ghost lists), numbers the
      ghost columns by first appearance in the own_ghost list and compresses again: own sum + sender sums in order.
    This is setup-time index work on the host; the ghost-row entries travel through the metadata channel (ship="host") or
    through the device exchange! (ship="device"); the final compression runs on the device with compress="device"."""
    b = rows.backend
    if local_format not in ("csr", "csc"):
        raise ValueError("local_format must be 'csr' or 'csc'")
    own_lists, outgoing = [], []
    for ind_r, ind_c, i, j, v in zip(rows.indices, cols.indices, I, J, V):
        i, j, v = np.asarray(i, dtype=np.int64), np.asarray(j, dtype=np.int64), np.asarray(v, dtype=np.float64)
        rsa = pr.union_ghost(ind_r, i, _find_owner(rows, ind_r, i))
        csa = pr.union_ghost(ind_c, j, _find_owner(cols, ind_c, j))
        if not (rsa.own_is_prefix and csa.own_is_prefix):
            raise ValueError("psparse(disassembled): needs own-first local orders")
        li, lj = rsa.global_to_local(i).astype(np.int64), csa.global_to_local(j).astype(np.int64)
        li[i < 1] = 0
        lj[j < 1] = 0
        ei, ej, ev = _stored_entries(li, lj, v, rsa.n_local, csa.n_local, local_format, b)
        row_own, col_own = ei <= rsa.n_own, ej <= csa.n_own
        gj_ghost = lambda m: csa.ghost_to_global[ej[m] - csa.n_own - 1]
        m_oo, m_og = row_own & col_own, row_own & ~col_own
        own_lists.append(((ei[m_oo], ej[m_oo], ev[m_oo]), (ei[m_og], gj_ghost(m_og), ev[m_og])))
        sel = np.concatenate([np.nonzero(~row_own & col_own)[0], np.nonzero(~row_own & ~col_own)[0]])
        gi = rsa.ghost_to_global[ei[sel] - rsa.n_own - 1]
        gj = csa.local_to_global[ej[sel] - 1] if len(sel) else np.zeros(0, np.int64)
        owner = rsa.ghost_to_owner[ei[sel] - rsa.n_own - 1]
        outgoing.append((ind_r.part, {int(q): (gi[owner == q], gj[owner == q], ev[sel][owner == q]) for q in np.unique(owner)}))
    received = _ship_ghost_rows(b, outgoing, ship)
    new_cols, lists = [], []
    for (oo, og), ind_r, ind_c, (ri, rj, rv) in zip(own_lists, rows.indices, cols.indices, received):
        me = ind_r.part
        rli = ind_r.global_to_local(ri).astype(np.int64)   # own-first: own id == local id
        rlj = ind_c.global_to_local(rj).astype(np.int64)
        if np.any(rli < 1) or np.any(rli > ind_r.n_own):
            raise ValueError("psparse(disassembled): received a row this part does not own")
        jown = (rlj >= 1) & (rlj <= ind_c.n_own)
        oo_all = (np.concatenate([oo[0], rli[jown]]), np.concatenate([oo[1], rlj[jown]]), np.concatenate([oo[2], rv[jown]]))
        og_i, og_gj, og_v = np.concatenate([og[0], rli[~jown]]), np.concatenate([og[1], rj[~jown]]), np.concatenate([og[2], rv[~jown]])
        if ind_c.block is not None and ind_c.own_is_prefix:  # remove_ghost
            c_own = pr.LocalIndices(ind_c.n_global, me, block=ind_c.block, n_own=ind_c.n_own)
        else:
            if not ind_c.own_is_prefix:
                raise ValueError("psparse(disassembled): needs own-first local orders")
            c_own = pr.LocalIndices(ind_c.n_global, me, ind_c.own_to_global, np.full(ind_c.n_own, me, dtype=np.int32))
        cfa = pr.union_ghost(c_own, og_gj, _find_owner(cols, ind_c, og_gj))
        new_cols.append(cfa)
        og_all = (og_i, cfa.global_to_local(og_gj).astype(np.int64) - cfa.n_own, og_v)
        lists.append((oo_all, og_all))
    colr = PRange(b, new_cols)
    A = PSparseMatrix(rows, colr)
    for k, ((oo, og), ind_r, ind_c) in enumerate(zip(lists, rows.indices, colr.indices)):
        if not ind_r.own_is_prefix:
            raise ValueError("psparse(disassembled): needs own-first local orders")
        no_r, no_c, ng_c = ind_r.n_own, ind_c.n_own, ind_c.n_ghost
        if compress == "device":
            A.set_coo(k, np.concatenate([oo[0], og[0]]), np.concatenate([oo[1], og[1] + no_c]), np.concatenate([oo[2], og[2]]))
            continue
        boo, bog = _coo_to_csr(*oo, no_r, no_c), _coo_to_csr(*og, no_r, ng_c)
        if split_format:
            if local_format == "csc":
                A.set_csc_split(k, no_r, _csr_to_csc(*boo, no_c), _csr_to_csc(*bog, ng_c))
            else:
                A.set_csr_split(k, boo, bog)
        else:
            rid = lambda blk: np.repeat(np.arange(1, no_r + 1), np.diff(blk[0].astype(np.int64)))
            rp, cv, nz = _coo_to_csr(np.concatenate([rid(boo), rid(bog)]),
                                     np.concatenate([boo[1].astype(np.int64), bog[1].astype(np.int64) + no_c]),
                                     np.concatenate([boo[2], bog[2]]), no_r, no_c + ng_c)
            if local_format == "csc":
                if ind_r.n_local != no_r:
                    raise ValueError("unsplit CSC upload needs a ghost-free row partition")
                A.set_csc(k, no_r, *_csr_to_csc(rp, cv, nz, no_c + ng_c))
            else:
                A.set_csr(k, rp, cv, nz)
    return A.commit()


def pvector_from_triplets(I: List, V: List, rows: PRange) -> PVector:
    """pvector(I,V,rows) with its defaults (src/p_vector.jl:887-926): disassembled contributions in global ids;
    rows_sa = union_ghost(rows, I), dense_vector per part (a[i] += v in input order, ids < 1 skipped, :853-863), then
    assemble(a, rows) (:1331-1347): assemble! on the device (owners add their neighbours' contributions in neighbour
    order), result on `rows`."""
    b = rows.backend
    sa, vals = [], []
    for ind, i, v in zip(rows.indices, I, V):
        i, v = np.asarray(i, dtype=np.int64), np.asarray(v, dtype=np.float64)
        rsa = pr.union_ghost(ind, i, _find_owner(rows, ind, i))
        a = np.zeros(rsa.n_local)
        ok = i >= 1
        np.add.at(a, rsa.global_to_local(i[ok]).astype(np.int64) - 1, v[ok])
        sa.append(rsa)
        vals.append(a)
    w = PVector(PRange(b, sa)).set_local_values(vals)
    w.assemble_()
    out = PVector(rows)
    out.copy_(w)  # w .= v2: own values (the layouts differ only in their ghosts)
    w.free()
    return out


def _find_owner(cols: PRange, ind: pr.LocalIndices, gids: np.ndarray) -> np.ndarray:
    """find_owner (src/p_range.jl:346-348): block formula when available, else a gathered table."""
    if ind.block is not None:
        return ind.block.owner_of(gids)
    b = cols.backend
    if not hasattr(cols, "_owner_table"):
        tab = np.zeros(len(cols) + 1, dtype=np.int32)
        for part, own in b.gather_all([(i.part, i.own_to_global) for i in cols.indices]):
            tab[own] = part
        cols._owner_table = tab
    out = np.zeros(len(gids), dtype=np.int32)
    ok = gids >= 1
    out[ok] = cols._owner_table[gids[ok]]
    return out


class CGResult:
    def __init__(self, iters, converged, residual0, residual, history):
        self.iters, self.converged, self.residual0, self.residual, self.history = iters, converged, residual0, residual, history


def ref_cg_(x: PVector, A: PSparseMatrix, b: PVector, tolerance: float = 0.0, maxiter: Optional[int] = None, flags: int = 0) -> CGResult:
    """ref_cg!(x,A,b; tolerance, maxiter, Pl=Identity) (HPCG/src/ref_cg.jl:119-134) -> x updated in place."""
    maxiter = len(A.cols) if maxiter is None else int(maxiter)
    res = _capi.CGResult()
    hist = np.zeros(maxiter + 1, dtype=np.float64)
    check(_capi.lib().pa_cg(A.h, x.h, b.h, maxiter, float(tolerance), flags, C.byref(res), ptr(hist)))
    return CGResult(res.iters, bool(res.converged), res.residual0, res.residual, hist[: res.iters + 1])


def opt_cg_(x, A, b, **kw):
    """opt_cg! forwards to ref_cg! in the reference (HPCG/src/opt_cg.jl:25-32); here it is the fused schedule."""
    return ref_cg_(x, A, b, **kw)


# ---------------------------------------------------------------------------------------------------- spmm / spmtm / rap
def _remove_ghost(r: PRange) -> PRange:
    """remove_ghost on every part (src/p_range.jl:1404-1410): the own ids only."""
    out = []
    for ind in r.indices:
        if ind.block is not None and ind.own_is_prefix:
            out.append(pr.LocalIndices(ind.n_global, ind.part, block=ind.block, n_own=ind.n_own))
        else:
            out.append(pr.LocalIndices(ind.n_global, ind.part, ind.own_to_global, np.full(ind.n_own, ind.part, dtype=np.int32)))
    return PRange(r.backend, out)


def consistent_matrix(B: PSparseMatrix, rows_co: PRange) -> PSparseMatrix:
    """C = consistent(B, rows_co) (src/p_sparse_matrix.jl:2243 and the consistent(::PSparseMatrix, rows_co) it calls): a
    matrix with one row per LOCAL id of rows_co — the own rows of B plus, for every ghost id of rows_co, the row of B its
    owner holds.  The ghost rows travel with exchange! on the device (row lengths, global column ids, values: each receiver
    pulls its segments from the owners' HBM); the index bookkeeping (which rows, local numbering of the new ghost columns)
    is setup-time host work, as in the reference."""
    b = B.backend
    rows_co.plan  # builds rows_co.plans
    snd_to, payload = [], [[], [], []]
    parts_csr = []
    for k, (ind_b, ind_c, pl) in enumerate(zip(B.cols.indices, rows_co.indices, rows_co.plans)):
        if not (ind_c.own_is_prefix and ind_b.own_is_prefix):
            raise ValueError("consistent(B, rows_co): needs own-first local orders")
        if B.rows.indices[k].n_own != ind_c.n_own:
            raise ValueError("consistent(B, rows_co): the own rows of B and of rows_co differ")
        rp, cv, nz = B.download_csr(k)
        l2g = ind_b.local_to_global
        parts_csr.append((rp, cv, nz, l2g))
        snd_to.append([int(q) for q in pl.nbr_rcv])
        lens, gids, vals = [], [], []
        for i in range(len(pl.nbr_rcv)):
            rows = pl.rcv_lids[pl.rcv_ptrs[i] - 1 : pl.rcv_ptrs[i + 1] - 1].astype(np.int64) - 1  # my own rows this neighbour needs
            cnt = rp[rows + 1] - rp[rows]
            take = np.concatenate([np.arange(rp[r], rp[r + 1]) for r in rows]) if len(rows) and cnt.sum() else np.zeros(0, np.int64)
            lens.append(cnt.astype(np.int64)); gids.append(l2g[cv[take]].astype(np.int64)); vals.append(nz[take])
        payload[0].append(lens); payload[1].append(gids); payload[2].append(vals)
    graph = ExchangeGraph(b, snd_to)
    got = [exchange(p, graph) for p in payload]
    new_cols, locals_ = [], []
    for k, (ind_b, ind_c, pl) in enumerate(zip(B.cols.indices, rows_co.indices, rows_co.plans)):
        rp, cv, nz, l2g = parts_csr[k]
        if [int(q) for q in graph.rcv[k]] != [int(q) for q in pl.nbr_snd]:
            raise RuntimeError("consistent(B, rows_co): exchange graph and ghost owners disagree")
        n_own, n_loc = ind_c.n_own, ind_c.n_local
        row_len = np.zeros(n_loc, dtype=np.int64)
        row_len[:n_own] = np.diff(rp)
        ghost_rows = {}
        all_g = []
        for i in range(len(pl.nbr_snd)):
            lids = pl.snd_lids[pl.snd_ptrs[i] - 1 : pl.snd_ptrs[i + 1] - 1].astype(np.int64) - 1  # my ghost rows owned by this neighbour
            lens_i, g_i, v_i = got[0][k][i].astype(np.int64), got[1][k][i].astype(np.int64), got[2][k][i].astype(np.float64)
            cuts = np.concatenate([[0], np.cumsum(lens_i)])
            for t, L in enumerate(lids):
                ghost_rows[int(L)] = (g_i[cuts[t] : cuts[t + 1]], v_i[cuts[t] : cuts[t + 1]])
            row_len[lids] = lens_i
            all_g.append(g_i)
        all_g = np.concatenate(all_g) if all_g else np.zeros(0, np.int64)
        cC = pr.union_ghost(ind_b, all_g, _find_owner(B.cols, ind_b, all_g))
        new_cols.append(cC)
        rp2 = np.zeros(n_loc + 1, dtype=np.int64)
        np.cumsum(row_len, out=rp2[1:])
        cv2 = np.zeros(int(rp2[-1]), dtype=np.int32)
        nz2 = np.zeros(int(rp2[-1]), dtype=np.float64)
        cv2[: len(cv)] = cv  # the own rows keep B's local column ids (cC appends its new ghosts behind B's)
        nz2[: len(nz)] = nz
        for L, (g, v) in ghost_rows.items():
            lc = cC.global_to_local(g).astype(np.int64) - 1
            o = np.argsort(lc, kind="stable")
            cv2[rp2[L] : rp2[L + 1]] = lc[o]
            nz2[rp2[L] : rp2[L + 1]] = v[o]
        locals_.append((rp2, cv2, nz2))
    C = PSparseMatrix(rows_co, PRange(b, new_cols))
    for k, (rp2, cv2, nz2) in enumerate(locals_):
        C.set_csr(k, rp2, cv2, nz2, index_base=0)
    return C.commit()


def spmm(A: PSparseMatrix, B: PSparseMatrix) -> PSparseMatrix:
    """spmm(A, B) = A*B (src/p_sparse_matrix.jl:2237-2262): C = consistent(B, axes(A,2)), then the local products
    D_k = A_k * C_k on the device (pa_mat_spmm_local); D is assembled on (axes(A,1), axes(C,2))."""
    if not (A.assembled and B.assembled):
        raise ValueError("spmm: both matrices must be assembled")
    Cm = consistent_matrix(B, A.cols)
    D = PSparseMatrix(A.rows, Cm.cols)
    check(_capi.lib().pa_mat_spmm_local(A.h, Cm.h, D.h))
    D.commit()
    Cm.free()
    return D


def spmtm(A: PSparseMatrix, B: PSparseMatrix) -> PSparseMatrix:
    """spmtm(A, B) = transpose(A)*B (src/p_sparse_matrix.jl:2276-2290): local products (A_k)^T * B_k on the device give a
    sub-assembled matrix on (axes(A,2), axes(B,2)) whose ghost rows then travel to their owners (assemble)."""
    if not (A.assembled and B.assembled):
        raise ValueError("spmtm: both matrices must be assembled")
    b = A.backend
    T = PSparseMatrix(A.cols, A.rows)
    check(_capi.lib().pa_mat_transpose_local(A.h, T.h))
    T.commit()
    Ds = PSparseMatrix(A.cols, B.cols)
    check(_capi.lib().pa_mat_spmm_local(T.h, B.h, Ds.h))
    Ds.commit()
    I, J, V = [], [], []
    for k, (ir, ic) in enumerate(zip(A.cols.indices, B.cols.indices)):
        rp, cv, nz = Ds.download_csr_all(k)
        rowid = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
        I.append(ir.local_to_global[rowid]); J.append(ic.local_to_global[cv]); V.append(nz)
    T.free(); Ds.free()
    return psparse(I, J, V, _remove_ghost(A.cols), _remove_ghost(B.cols), assembled=False, compress="device")


def rap(R: PSparseMatrix, A: PSparseMatrix, P: PSparseMatrix) -> PSparseMatrix:
    """rap(R, A, P) = R*A*P (src/p_sparse_matrix.jl:2212-2218)."""
    RA = spmm(R, A)
    out = spmm(RA, P)
    RA.free()
    return out
