"""Generates tests/golden/*.npz|json from the ORACLE (the reference cannot run here: it needs
deal.II).  The fixtures pin (a) the oracle against regressions and (b) the CUDA path on the
GPU box without needing anything but numpy.  External anchors (upstream deal.II step-64
tutorial output) are hard-coded in tests/test_oracle_known_answers.py, not generated here.

    python scripts/make_golden.py
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as O
from conftest import ladder

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)

# 1) operator application on seeded inputs (seed = degree), small ragged meshes
cases = {}
for p in range(1, 9):
    cells = (3, 2, 4) if p <= 4 else (2, 3, 2) if p <= 6 else (2, 1, 2)
    for quad in (O.GAUSS, O.GLL):
        for kind in (O.POISSON, O.HELMHOLTZ):
            for deform in (0, 1):
                m = O.OracleMesh(p, cells, quad=quad, deform=deform, eps=0.1)
                u = np.random.default_rng(p).standard_normal(m.n_dofs)
                v = m.vmult(u, kind=kind)
                key = f"p{p}_q{quad}_k{kind}_d{deform}"
                cases[key + "_cells"] = np.array(cells)
                cases[key + "_out"] = v
np.savez_compressed(os.path.join(out, "vmult_cases.npz"), **cases)

# 2) BP5 ladder, degree 5 (the reference's default degree, bp5/step-64.cu:725), merged CG
lad = {}
for cyc in (7, 8, 12, 13):
    cells, upper = ladder(cyc)
    for quad, qn in ((O.GAUSS, "gauss"), (O.GLL, "gll")):
        m = O.OracleMesh(5, cells, quad=quad, upper=upper)
        b = m.rhs(); tol = 1e-6 * np.linalg.norm(b)
        x, its, res, hist, ok = m.cg(b, variant=1, control=0, tol=tol, max_its=200)
        x0, its0, *_ = m.cg(b, variant=0, control=0, tol=tol, max_its=200)
        lad[f"cycle{cyc}_{qn}"] = dict(cells=cells, upper=upper, n_dofs=int(m.n_dofs), its_merged=int(its), its_standard=int(its0),
                                        x_l2=float(np.linalg.norm(x)), b_l2=float(np.linalg.norm(b)), res=float(res),
                                        history=[float(h) for h in hist], solution_l2=float(m.l2_norm(x)))
json.dump(lad, open(os.path.join(out, "bp5_ladder_p5.json"), "w"), indent=1)

# 3) config 1 (BP5 p=4, 32^3 cells, 2,146,689 DoFs): scalars only
m = O.OracleMesh(4, (32, 32, 32), quad=O.GAUSS, upper=(1., 1., 1.))
b = m.rhs(); tol = 1e-6 * np.linalg.norm(b)
x, its, res, hist, ok = m.cg(b, variant=1, control=0, tol=tol, max_its=200)
u = np.random.default_rng(4).standard_normal(m.n_dofs); u[m.boundary_mask()] = 0
v = m.vmult(u)
json.dump(dict(n_dofs=int(m.n_dofs), its=int(its), b_l2=float(np.linalg.norm(b)), x_l2=float(np.linalg.norm(x)),
               rel_res=float(res / np.linalg.norm(b)), history_every_20=[float(h) for h in hist[::20]],
               vmult_seed4_l2=float(np.linalg.norm(v)), vmult_seed4_sum=float(v.sum()),
               vmult_seed4_sample=[float(t) for t in v[:: m.n_dofs // 16][:16]]),
          open(os.path.join(out, "config1_p4_32.json"), "w"), indent=1)
print("golden written")
