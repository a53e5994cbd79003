"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol include/pa_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import pa_b200

    lib = pa_b200._capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in pa_b200.h but not exported"
    assert lib.pa_abi_version() == 1
    # every declared symbol has a ctypes signature in the binding
    assert set(syms) <= set(pa_b200._capi.SIGNATURES), set(syms) - set(pa_b200._capi.SIGNATURES)


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly (no oracle / CPU route)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pa_b200

    with pytest.raises(pa_b200.PAError, match="no CUDA device"):
        pa_b200.CUDAArray(1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "partitionedarrays.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "README.md", f"{f} mentions the oracle"


def test_sass_has_no_fma_in_spmv_accumulate():
    """The SpMV kernel must use separate DMUL/DADD (bit parity with spmv_csr!'s `bi += aij*xj`)."""
    import shutil
    import subprocess

    import pa_b200

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump missing")
    out = subprocess.run([cuobjdump, "-sass", "-fun", "_Z13k_spmv_streamIiLi256ELi2048ELb0EEv8SpmvArgsIT_E", pa_b200.build.SO],
                         capture_output=True, text=True).stdout
    assert "DMUL" in out and "DADD" in out and "DFMA" not in out
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", pa_b200.build.SO], capture_output=True, text=True).stdout
    # the production kernel: TMA bulk copies (UBLKCP) + mbarriers (SYNCS), separate DMUL/DADD, no DFMA
    sass = subprocess.run([cuobjdump, "-sass", pa_b200.build.SO], capture_output=True, text=True).stdout
    blocks = sass.split("Function : ")
    tma = [b for b in blocks if b.startswith("_Z10k_spmv_tmaIiLi0ELi8E")]
    assert tma, "k_spmv_tma<int,0,8> not found in the library"
    assert "UBLKCP" in tma[0] and "SYNCS" in tma[0] and "DMUL" in tma[0] and "DADD" in tma[0] and "DFMA" not in tma[0]
    # the row-pattern kernels (round 2), every instantiation: TMA ring, separate DMUL/DADD, no DFMA in the products
    pat = [b for b in blocks if b.startswith("_Z10k_spmv_patI")]
    assert len(pat) >= 6, "k_spmv_pat instantiations missing"
    for b in pat:
        assert "UBLKCP" in b and "SYNCS" in b and "DMUL" in b and "DADD" in b and "DFMA" not in b, b.split("\n")[0]
    # the multi-colour Gauss-Seidel kernel streams its slices with TMA too (its DFMAs are the division sequence)
    col = [b for b in blocks if b.startswith("_Z14k_gs_color_tmaILi27E")]
    assert col and all("UBLKCP" in b and "SYNCS" in b and "DMUL" in b and "DADD" in b for b in col)
    gs = [b for b in blocks if b.startswith("_Z9k_gs_flowIiE")]
    # Gauss-Seidel sweeps: separate DMUL / DADD for s -= a*x (the only DFMAs belong to the IEEE division sequence of __ddiv_rn)
    assert gs and "DMUL" in gs[0] and "DADD" in gs[0] and "MUFU.RCP64H" in gs[0]
