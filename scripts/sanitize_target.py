"""Tiny workload for compute-sanitizer (racecheck / memcheck / synccheck): every cell-kernel family once, a few CG
iterations of each solver.  usage: compute-sanitizer --tool racecheck python scripts/sanitize_target.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dealceed_b200 as dc
os.environ.setdefault("BP5_NO_GRAPH", "1")
ctx = dc.Context(0)
CASES = [(3, dc.QUAD_GLL, dc.OP_POISSON, dc.GEOM_STORED, {}), (2, dc.QUAD_GAUSS, dc.OP_HELMHOLTZ, dc.GEOM_STORED, {}),
         (4, dc.QUAD_GLL, dc.OP_POISSON, dc.GEOM_ON_THE_FLY, {}),
         # coloured cell order (eight launches, plain adds) and locally refined meshes (hanging-node exchange)
         (3, dc.QUAD_GLL, dc.OP_POISSON, dc.GEOM_STORED, dict(cell_order=dc.CELL_ORDER_COLORED)),
         (3, dc.QUAD_GLL, dc.OP_POISSON, dc.GEOM_STORED, dict(refine_lo=(1, 0, 1), refine_hi=(2, 2, 2))),
         (2, dc.QUAD_GAUSS, dc.OP_HELMHOLTZ, dc.GEOM_STORED, dict(refine_lo=(0, 0, 0), refine_hi=(2, 1, 1)))]
for p, quad, kind, geom, extra in CASES:
    op = dc.PoissonOperator(ctx, dc.make_problem(p, (3, 2, 2), quadrature=quad, operator_kind=kind, deformation=1, eps=0.1,
                                                 geometry_mode=geom, **extra))
    b, x, y = op.initialize_dof_vector(), op.initialize_dof_vector(), op.initialize_dof_vector()
    op.assemble_rhs(b)
    op.vmult(y, b)
    op.cell_loop(y, b)
    for Solver, zero in ((dc.SolverCGFullMerge, False), (dc.SolverCG, True)):
        op.do_zero_out = zero
        x.set(0.0)
        ctl = dc.IterationNumberControl(6, 0.0)
        Solver(ctl).solve(op, x, b, history=False)
    print("ok", op.kernel_name, x.l2_norm(), flush=True)
    for v in (b, x, y):
        v.close()
    op.close()
ctx.close()
