"""Small fixed workload for ncu: a few vmults (zero + cell kernel + constrained copy) of one configuration.
usage: python scripts/ncu_target.py <degree> <gll|gauss> <cells_per_dir> [n_vmults] [stored|otf] [deformation eps]
NCU_TARGET_CG=1: n_vmults merged-CG iterations instead (the in-loop kernel with the fused dot product; graph replay off)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dealceed_b200 as dc
p, quad, cells = int(sys.argv[1]), sys.argv[2], int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
geom = dc.GEOM_ON_THE_FLY if len(sys.argv) > 5 and sys.argv[5] == "otf" else dc.GEOM_STORED
eps = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0
ctx = dc.Context(0)
op = dc.PoissonOperator(ctx, dc.make_problem(p, (cells,) * 3, quadrature=dc.QUAD_GLL if quad == "gll" else dc.QUAD_GAUSS,
                                             geometry_mode=geom, deformation=1 if eps else 0, eps=eps))
src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
src.import_host(np.random.default_rng(0).standard_normal(op.n_owned))
if os.environ.get("NCU_TARGET_CG"):
    os.environ["BP5_NO_GRAPH"] = "1"
    op.assemble_rhs(src)
    op.do_zero_out = False
    dst.set(0.0)
    dc.SolverCGFullMerge(dc.IterationNumberControl(reps, 0.0)).solve(op, dst, src, history=False)
else:
    for _ in range(reps):
        op.vmult(dst, src)
ctx.synchronize()
print("done", op.kernel_name, op.n_owned, "algorithmic bytes per vmult", op.algorithmic_bytes()[0])
