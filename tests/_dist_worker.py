"""Worker for the multi-process tests (launched by torchrun).

mode=gloo : CPU only — host-side distributed plan logic (build_plans over an all-gather) vs the oracle.
mode=gpu  : one part per GPU — CUDA IPC peer mapping, in-kernel NVLink ghost loads, signalling flags,
            NCCL scalar all-reduce, CG — all against the oracle computed redundantly on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import c_oracle, pa_oracle as o  # noqa: E402


def check_plan(pa, npd, n, ghost, rank, world, gather_all):
    from pa_b200 import prange as pr

    op = o.uniform_partition(npd, n, ghost)
    mine = pr.uniform_partition_part(rank + 1, npd, n, ghost)
    plan = pr.build_plans([mine], gather_all)[0]
    oplan = o.assembly_plan(op)
    k = rank
    assert plan.nbr_snd.tolist() == oplan.neighbors_snd[k].tolist()
    assert plan.nbr_rcv.tolist() == oplan.neighbors_rcv[k].tolist()
    assert plan.snd_lids.tolist() == oplan.local_indices_snd[k].data.tolist()
    assert plan.rcv_lids.tolist() == oplan.local_indices_rcv[k].data.tolist()
    # remote lids: my snd entry j towards q is q's rcv entry at the same position
    pos = 0
    for i, q in enumerate(plan.nbr_snd):
        seg = oplan.local_indices_rcv[q - 1]
        j = oplan.neighbors_rcv[q - 1].tolist().index(rank + 1)
        want = seg.segment(j)
        assert plan.snd_remote_lids[pos : pos + len(want)].tolist() == want.tolist()
        pos += len(want)


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if mode == "gloo":
        dist.init_process_group("gloo")
        import pa_b200 as pa

        def gather_all(objs):
            out = [None] * world
            dist.all_gather_object(out, objs)
            return [x for per in out for x in per]

        check_plan(pa, (world,), (11,), (True,), rank, world, gather_all)
        check_plan(pa, (world, 1), (6, 5), (True, True), rank, world, gather_all)
        check_plan(pa, (1, world), (4, 7), None, rank, world, gather_all)
        # stencil column partitions through the same machinery
        from pa_b200 import prange as pr

        for kind, npd, nloc in ((7, (world, 1, 1), (3, 4, 2)), (27, (1, world, 1), (3, 3, 3))):
            gn = tuple(a * b for a, b in zip(npd, nloc))
            if kind == 7:
                I, J, V, rows, cols = o.laplacian_fdm(gn, npd)
                Ao = o.psparse(I, J, V, rows, cols, assembled=True)
            else:
                Ao, _ = o.hpcg_build_p_matrix(*nloc, *npd)
            mine = pr.stencil_col_indices(kind, rank + 1, npd, gn)
            plan = pr.build_plans([mine], gather_all)[0]
            oplan = o.assembly_plan(Ao.col_partition)
            assert plan.snd_lids.tolist() == oplan.local_indices_snd[rank].data.tolist()
            assert plan.rcv_lids.tolist() == oplan.local_indices_rcv[rank].data.tolist()
        # exchange primitive, host side: receive ids and buffer layout over a real all-gather (one part per process)
        class _Meta:
            parts = [rank + 1]

            @staticmethod
            def gather_all(objs):
                return gather_all(objs)

        snd_all = [[q for q in range(1, world + 1) if q != p] for p in range(1, world + 1)]  # everybody sends to everybody else
        len_all = [[3 * p + q for q in s] for p, s in enumerate(snd_all, 1)]
        g = pa.ExchangeGraph(_Meta, [snd_all[rank]])
        assert g.rcv == [o.find_rcv_ids(snd_all)[rank]]
        rcv_len, rcv_off, sym = pa.exchange_layout(g, [len_all[rank]])
        assert sym == max(sum(l) for l in len_all)
        for i, src in enumerate(g.rcv[0]):
            j = snd_all[src - 1].index(rank + 1)
            assert rcv_len[0][i] == len_all[src - 1][j] and rcv_off[0][i] == sum(len_all[src - 1][:j])
        dist.barrier()
        if rank == 0:
            print("GLOO_WORKER_OK")
        dist.destroy_process_group()
        return

    # ---------------- GPU mode
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    meta = dist.new_group(backend="gloo")
    import pa_b200 as pa

    shape = pa.compute_optimal_shape_xyz(world)
    backend = pa.CUDAArray(world, mode="distributed", device=local_rank, arena_bytes=64 << 20, group=meta)
    for kind, nloc in ((7, (8, 6, 5)), (27, (6, 6, 6))):
        gn = tuple(a * b for a, b in zip(shape, nloc))
        if kind == 7:
            I, J, V, rows, cols = o.laplacian_fdm(gn, shape)
            Ao = o.psparse(I, J, V, rows, cols, assembled=True)
            bo = None
        else:
            Ao, bo = o.hpcg_build_p_matrix(*nloc, *shape)
        part = Ao.col_partition
        plan = o.assembly_plan(part)
        A, rhs = pa.stencil_matrix(kind, gn, shape, backend)
        ind = part[rank]
        rp, cv, nz = A.download_csr(0)
        assert np.array_equal(rp, Ao.local[rank].rowptr.astype(np.int64) - 1) and np.array_equal(cv, Ao.local[rank].colval - 1)
        xg = o.hash_uniform(np.arange(1, int(np.prod(gn)) + 1), 9)
        xo = o.pvector_from_global(xg, part, ghosts=False)
        co = [np.zeros(i.n_local) for i in Ao.row_partition]
        o.mul_no_lat(Ao, xo, plan, co)
        x = pa.fill_hash(pa.PVector(A.cols), 9)
        y = pa.pzeros(A.rows)
        for flags in (pa.PA_SPMV_DEFAULT, pa.PA_SPMV_FUSED_EXCHANGE, pa.PA_SPMV_OVERLAP, pa.PA_SPMV_INLINE_PEER_LOADS,
                      pa.PA_SPMV_INLINE_PEER_LOADS | pa.PA_SPMV_SKIP_GHOST_REFRESH, pa.PA_SPMV_FUSED_EXCHANGE):
            y.fill_(-7.0)
            x2 = pa.fill_hash(pa.PVector(A.cols), 9)
            pa.mul_(y, A, x2, flags=flags)
            got = y.local_values()[0][: ind.n_own]
            assert np.array_equal(got, co[rank][: ind.n_own]), f"rank {rank} kind {kind} flags {flags}: SpMV differs from the oracle"
            x2.free()
        # repeated mul! on changing data: exercises the ready/done epochs
        for rep in range(5):
            x.rmul_(2.0)
            pa.mul_(y, A, x)
        got = y.local_values()[0][: ind.n_own]
        assert np.array_equal(got, 32.0 * co[rank][: ind.n_own])
        xl = x.local_values()[0]
        assert np.array_equal(xl, 32.0 * xg[ind.local_to_global - 1]), "consistent! side effect of mul! (ghosts)"
        # assemble!: ghosts flow back to the owners
        v = pa.pfill(1.0, A.cols)
        v.assemble_().wait()
        vo = [np.ones(i.n_local) for i in part]
        o.assemble(vo, part, plan)
        assert np.array_equal(v.local_values()[0], vo[rank])
        # reductions across processes (NCCL all-reduce of the per-part partials)
        d = x.dot(x)
        want = float(np.dot(32.0 * xg, 32.0 * xg))
        assert abs(d - want) <= 1e-12 * want, (d, want)
        # CG
        if bo is None:
            ones = [np.ones(i.n_local) for i in part]
            bo = [np.zeros(i.n_local) for i in part]
            o.pmul(Ao, ones, plan, bo)
        mats = [(i.n_own, i.n_local, Ao.local[p].rowptr.astype(np.int64) - 1, Ao.local[p].colval.astype(np.int32) - 1, Ao.local[p].nzval)
                for p, i in enumerate(part)]
        prob = c_oracle.CGProblem(mats, plan, bo, [np.zeros(i.n_local) for i in part])
        it_o, hist_o, _ = prob.cg(12, 0.0)
        for flags, strategy in ((0, -1), (pa.PA_CG_REFERENCE_OPS, -1), (0, 3)):  # 3: consistent! fused into the SpMV kernel
            backend.set_knob("spmv_strategy", strategy)
            xs = pa.pzeros(A.cols)
            res = pa.ref_cg_(xs, A, rhs, tolerance=0.0, maxiter=12, flags=flags)
            np.testing.assert_allclose(res.history, hist_o, rtol=1e-8, atol=1e-12 * hist_o[0])
            np.testing.assert_allclose(xs.local_values()[0][: ind.n_own], prob.x[rank][: ind.n_own], rtol=1e-8, atol=1e-10)
            xs.free()
        backend.set_knob("spmv_strategy", -1)
        for obj in (x, y, v, rhs):
            obj.free()
        A.free()
    backend.sync()
    dist.barrier()
    if rank == 0:
        print("GPU_WORKER_OK")
    backend.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
