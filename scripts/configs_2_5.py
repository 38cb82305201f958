"""BASELINE configs 2 and 5 on one B200 (full size), JSON lines:
 config 2: step-64 variable-coefficient Helmholtz, p=4, 64^3 cells = 257^3 = 16,974,593 DoFs, SolverControl(n_dofs, 1e-12|b|)
 config 5: BP5 p=5 on a smoothly deformed (curvilinear) mesh, stored metric tensor, 60^3 cells = 27.3 M DoFs
Each: merged CG + standard CG (iteration parity), throughput, fraction of the measured HBM peak, symmetry check."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import dealceed_b200 as dc
HBM = 6548.2
ctx = dc.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)

def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); fn(); e1.record(stream); e1.synchronize()
    return e0.elapsed_time(e1) * 1e-3

def run(name, prob, control_factory):
    op = dc.PoissonOperator(ctx, prob)
    n = op.n_owned
    bytes_v, bytes_cg = op.algorithmic_bytes()
    b, x, y = op.initialize_dof_vector(), op.initialize_dof_vector(), op.initialize_dof_vector()
    op.assemble_rhs(b)
    out = dict(config=name, dofs=n, kernel=op.kernel_name, b_l2=b.l2_norm())
    for Solver, key, zero in ((dc.SolverCGFullMerge, "merged", False), (dc.SolverCG, "standard", True)):
        ctl = control_factory(b.l2_norm(), n)
        op.do_zero_out = zero
        x.set(0.0); Solver(ctl).solve(op, x, b, history=False)           # warm-up
        x.set(0.0)
        t = timed(lambda: Solver(ctl).solve(op, x, b, history=False))
        out[key] = dict(its=ctl.last_step(), residual=ctl.last_value(), seconds=t, gdofs=n * ctl.last_step() / t / 1e9,
                        frac=bytes_cg * ctl.last_step() / t / 1e9 / HBM, x_l2=x.l2_norm())
    op.do_zero_out = True
    for _ in range(3): op.vmult(y, x)
    t = timed(lambda: [op.vmult(y, x) for _ in range(20)]) / 20
    out["vmult"] = dict(ms=t * 1e3, gdofs=n / t / 1e9, gbs=bytes_v / t / 1e9, frac=bytes_v / t / 1e9 / HBM)
    # symmetry on vectors vanishing on the boundary: x (solution) and b (rhs) both do
    Ab = op.initialize_dof_vector()
    op.vmult(y, x); op.vmult(Ab, b)
    out["symmetry_rel"] = abs(b.dot_local(y) - x.dot_local(Ab)) / abs(b.dot_local(y))
    print(json.dumps(out), flush=True)
    for v in (b, x, y, Ab): v.close()
    op.close()

run("config2_helmholtz_p4_64", dc.make_problem(4, (64, 64, 64), operator_kind=dc.OP_HELMHOLTZ, upper=(1., 1., 1.)),
    lambda bn, n: dc.SolverControl(n, 1e-12 * bn))
run("config5_bp5_p5_deformed_stored_metric_gauss", dc.make_problem(5, (60, 60, 60), deformation=1, eps=0.1),
    lambda bn, n: dc.IterationNumberControl(200, 1e-6 * bn))
run("config5_bp5_p5_deformed_stored_metric_gll", dc.make_problem(5, (60, 60, 60), quadrature=dc.QUAD_GLL, deformation=1, eps=0.1),
    lambda bn, n: dc.IterationNumberControl(200, 1e-6 * bn))
run("config5_bp5_p5_deformed_on_the_fly_geometry_gll",
    dc.make_problem(5, (60, 60, 60), quadrature=dc.QUAD_GLL, deformation=1, eps=0.1, geometry_mode=dc.GEOM_ON_THE_FLY),
    lambda bn, n: dc.IterationNumberControl(200, 1e-6 * bn))
ctx.close()
