"""Multi-process paths: world_size-2 gloo on CPU (host logic) and, on a multi-GPU box, one part per GPU."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc, mode, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_dist_worker.py"), mode]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_gloo_world2_plan_logic():
    r = _torchrun(2, "gloo", 300)
    assert r.returncode == 0 and "GLOO_WORKER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.gpu
def test_one_part_per_gpu_against_oracle():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    nproc = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = _torchrun(nproc, "gpu", 600)
    assert r.returncode == 0 and "GPU_WORKER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-6000:]


@pytest.mark.gpu
def test_one_process_drives_all_gpus():
    """pa_ctx_create_multi / with_cuda_multi: ONE process, one host thread and one context per GPU (the DebugArray model
    across the box): mul! and CG equal the oracle, as in the one-process-per-GPU model."""
    import numpy as np
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    import pa_b200 as pa
    from oracle import c_oracle, pa_oracle as o

    nd = 8 if n >= 8 else (4 if n >= 4 else 2)
    shape = pa.compute_optimal_shape_xyz(nd)
    nloc = (8, 6, 5)
    gn = tuple(a * b for a, b in zip(shape, nloc))
    I, J, V, rows, cols = o.laplacian_fdm(gn, shape)
    Ao = o.psparse(I, J, V, rows, cols, assembled=True)
    part = Ao.col_partition
    plan = o.assembly_plan(part)
    xg = o.hash_uniform(np.arange(1, int(np.prod(gn)) + 1), 9)
    xo = o.pvector_from_global(xg, part, ghosts=False)
    co = [np.zeros(i.n_local) for i in Ao.row_partition]
    o.mul_no_lat(Ao, xo, plan, co)
    ones = [np.ones(i.n_local) for i in part]
    bo = [np.zeros(i.n_local) for i in part]
    o.pmul(Ao, ones, plan, bo)
    mats = [(i.n_own, i.n_local, Ao.local[p].rowptr.astype(np.int64) - 1, Ao.local[p].colval.astype(np.int32) - 1, Ao.local[p].nzval)
            for p, i in enumerate(part)]
    prob = c_oracle.CGProblem(mats, plan, bo, [np.zeros(i.n_local) for i in part])
    it_o, hist_o, _ = prob.cg(12, 0.0)

    def spmd(backend):
        k = backend.parts[0] - 1
        A, rhs = pa.stencil_matrix(7, gn, shape, backend)
        x = pa.fill_hash(pa.PVector(A.cols), 9)
        y = pa.pzeros(A.rows)
        out = {}
        for flags in (pa.PA_SPMV_DEFAULT, pa.PA_SPMV_FUSED_EXCHANGE):
            y.fill_(-7.0)
            pa.mul_(y, A, x, flags=flags)
            out[flags] = y.local_values()[0][: part[k].n_own].copy()
        xs = pa.pzeros(A.cols)
        res = pa.ref_cg_(xs, A, rhs, tolerance=0.0, maxiter=12)
        out["hist"] = res.history.copy()
        out["dot"] = x.dot(x)
        for v in (x, y, xs, rhs):
            v.free()
        A.free()
        return out

    results = pa.with_cuda_multi(spmd, list(range(nd)), arena_bytes=64 << 20)
    for k, out in enumerate(results):
        for flags in (pa.PA_SPMV_DEFAULT, pa.PA_SPMV_FUSED_EXCHANGE):
            assert np.array_equal(out[flags], co[k][: part[k].n_own]), (k, flags)
        np.testing.assert_allclose(out["hist"], hist_o, rtol=1e-8, atol=1e-12 * hist_o[0])
        assert abs(out["dot"] - float(np.dot(xg, xg))) <= 1e-12 * float(np.dot(xg, xg))
