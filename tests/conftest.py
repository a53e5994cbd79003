import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The row-pattern compression of the TMA SpMV is built for parts of >= 4096 rows by default; the parity suite runs on
# small operators, so it lowers the threshold: every mul! below goes through the pattern kernel wherever rows repeat
# patterns (tests/test_gpu_patterns.py compares it with the plain column stream explicitly).
os.environ.setdefault("PA_SPMV_PATTERN_MIN_ROWS", "1")
os.environ.setdefault("PA_GS_PATTERN_MIN_ROWS", "1")  # same for the multi-colour Gauss-Seidel kernel


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")
