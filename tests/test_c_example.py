"""The C ABI from plain C: examples/hpcg_cg.c links against libpa_b200.so (no Python, no torch in the product path).
Without a GPU the program must stop with the library's "no CUDA device" error (exit code 2); on a GPU it solves."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    import pa_b200

    so = pa_b200.build.build()
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "hpcg_cg")
    libdir = os.path.dirname(so)
    subprocess.run([gcc, "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "hpcg_cg.c"), "-L", libdir,
                    "-lpa_b200", f"-Wl,-rpath,{libdir}", "-lm", "-o", exe], check=True)
    return exe


def test_c_example_builds_and_fails_loudly_without_gpu(tmp_path):
    import torch

    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, "8", "5"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "no CUDA device" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_c_example_solves_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "32", "25"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    m = re.search(r"25 CG iterations, \|\|r\|\|/\|\|r0\|\| = ([0-9.e+-]+)", r.stdout)
    assert m and float(m.group(1)) < 1.0, r.stdout
