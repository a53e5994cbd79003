"""Q1 FEM assembly -> mul! on one GPU at a size where the host stages take their long-array paths (device sort of the triplets,
dense gid -> lid table): A * u_exact == rhs to rounding (Q1 reproduces x1 + x2 exactly), both shipping routes give the same CSR."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pa_b200 as pa  # noqa: E402
from pa_b200 import fem_example as fe  # noqa: E402

nd = int(sys.argv[1]) if len(sys.argv) > 1 else 500
b = pa.CUDAArray(2, arena_bytes=256 << 20)
lay = fe.Q1Layout((2, 1), (nd + 1, nd + 1), (2.0, 2.0))
trip = [fe.q1_part(lay, p) for p in b.parts]
rows = pa.variable_partition(b, lay.n_own_dofs, lay.n_global_dofs)
mats = []
for ship in ("device", "host"):
    A = pa.psparse([t[0] for t in trip], [t[1] for t in trip], [t[2] for t in trip], rows, rows, assembled=False, local_format="csr",
                   compress="device", ship=ship)
    mats.append([A.download_csr(k) for k in range(2)])
    if ship == "device":
        rhs = pa.pvector_from_triplets([t[3] for t in trip], [t[4] for t in trip], rows)
        xe = pa.PVector(A.cols).set_local_values([np.concatenate([lay.exact_own(p), np.zeros(i.n_ghost)]) for p, i in zip(b.parts, A.cols.indices)])
        y = pa.pzeros(A.rows)
        pa.mul_(y, A, xe)
        resid = max(float(np.abs(a - c).max()) for a, c in zip(y.own_values(), rhs.own_values()))
        bound = 64 * np.finfo(float).eps * float(np.abs(lay.Ae).max() * 4.0)
        print(f"triplets per part {len(trip[0][0])}, |A*u_exact - rhs|_inf = {resid:.3e} (bound {bound:.3e})")
        assert resid < bound
for k in range(2):
    for a, c in zip(mats[0][k], mats[1][k]):
        assert np.array_equal(a, c)
print("FEM_QUICK_OK")
b.close()
