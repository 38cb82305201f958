"""Diagnostic: vmult parity vs the oracle for every degree / quadrature / operator (GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import dealceed_b200 as dc
import oracle as O

ctx = dc.Context(0)
worst = 0.0
for p in range(1, 9):
    cells = (3, 2, 4) if p <= 5 else (2, 3, 2)
    for quad in (dc.QUAD_GAUSS, dc.QUAD_GLL):
        for kind in (dc.OP_POISSON, dc.OP_HELMHOLTZ):
            for deform in (0, 1):
                t0 = time.time()
                op = dc.PoissonOperator(ctx, dc.make_problem(p, cells, quadrature=quad, operator_kind=kind,
                                                             deformation=deform, eps=0.1))
                m = O.OracleMesh(p, cells, quad=quad, deform=deform, eps=0.1)
                G = op.coefficients(); Go = m.metric()
                gerr = np.abs(G - Go).max() / np.abs(Go).max()
                rng = np.random.default_rng(p)
                u = rng.standard_normal(m.n_dofs)
                src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
                src.import_host(u)
                op.vmult(dst, src)
                got, ref = dst.to_host(), m.vmult(u, kind=kind)
                err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
                worst = max(worst, err)
                print(f"p={p} quad={quad} kind={kind} deform={deform} dofs={m.n_dofs} metric_err={gerr:.2e} vmult_err={err:.2e} "
                      f"{'OK' if err < 1e-12 else 'FAIL'} ({time.time()-t0:.2f}s)", flush=True)
                src.close(); dst.close(); op.close()
print("worst", worst)
ctx.close()
