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
