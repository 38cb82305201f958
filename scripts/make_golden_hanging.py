"""Generates tests/golden/hanging_cases.npz from oracle/hanging_oracle.py: operator application, right-hand side and CG
iteration count on small locally refined meshes (seeded inputs, seed = degree).  The fixture pins the hanging-node oracle
against regressions (tests/test_hanging_oracle.py) and the CUDA paths on the GPU box (tests/test_gpu_hanging_nodes.py).
The reference has no mesh with hanging nodes, so there is no reference-side vector to pin to.

    python scripts/make_golden_hanging.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle as O
from hanging_oracle import HangingMesh

CASES = [(1, (3, 3, 3), (1, 1, 1), (2, 2, 2), 0.0), (2, (4, 3, 3), (1, 0, 1), (3, 2, 2), 0.1),
         (3, (3, 3, 2), (0, 0, 0), (2, 2, 1), 0.1), (4, (3, 2, 2), (1, 1, 0), (2, 2, 2), 0.05)]
out = {}
for p, cells, lo, hi, eps in CASES:
    for quad in (O.GAUSS, O.GLL):
        hm = HangingMesh(p, cells, lo, hi, quad=quad, upper=(1., 1., 1.), deform=1 if eps else 0, eps=eps)
        u = np.random.default_rng(p).standard_normal(hm.n_dofs)
        b = hm.rhs()
        _, its, _ = hm.cg(b, tol=1e-8 * np.linalg.norm(b), max_its=1000)
        key = f"p{p}_q{quad}"
        out[key + "_spec"] = np.array([*cells, *lo, *hi], dtype=np.int64)
        out[key + "_eps"] = np.array(eps)
        out[key + "_n"] = np.array([hm.n_dofs, hm.n_cells], dtype=np.int64)
        out[key + "_Au"] = hm.vmult(u)
        out[key + "_b"] = b
        out[key + "_its"] = np.array(its)
        if quad == O.GAUSS:
            out[key + "_Hu"] = hm.vmult(u, kind=O.HELMHOLTZ)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hanging_cases.npz"), **out)
print("wrote", len(out), "arrays")
