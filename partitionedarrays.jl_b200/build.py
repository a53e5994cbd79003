"""Builds libpa_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels to the GPU box)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
SO = os.path.join(LIBDIR, "libpa_b200.so")
SOURCES = ["pa_runtime.cu", "pa_vector.cu", "pa_spmv.cu", "pa_cg.cu", "pa_mg.cu", "pa_assembly.cu", "pa_prims.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _source_hash() -> str:
    """Content hash of every source the library is built from (mtimes do not survive a snapshot copy)."""
    import hashlib

    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "pa_b200.h"), __file__]
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode() + b"\0" + fh.read())
    return h.hexdigest()


STAMP = os.path.join(LIBDIR, ".build_hash")


def needs_build() -> bool:
    if not os.path.exists(SO) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return SO
    import fcntl

    os.makedirs(LIBDIR, exist_ok=True)
    with open(os.path.join(LIBDIR, ".build_lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)  # several ranks may import at once: one builds, the others wait
        try:
            if not force and not needs_build():
                return SO
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose: bool) -> str:
    objs = []
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"] + ARCH
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc()] + common + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(f"--- {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    tmp = SO + ".tmp"
    link = [nvcc(), "-shared", "-o", tmp] + objs + ARCH + ["-lcudart_static", "-ldl", "-lpthread", "-lrt"]
    subprocess.run(link, check=True)
    os.replace(tmp, SO)  # atomic: a concurrent loader never sees a half-written library
    with open(STAMP, "w") as f:
        f.write(_source_hash())
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
