"""Tuning probe for the on-the-fly-geometry kernels on a deformed mesh (one JSON line per degree and kernel):
vmult and merged-CG rates of the stored metric, the collocation + Poisson kernel (apply_otf.cuh) and the general
kernel (apply_otfg.cuh).  BP5_LIB selects a tuning build (scripts/build_variant.sh), argv[1] labels the lines."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dealceed_b200 as dc
label = sys.argv[1] if len(sys.argv) > 1 else "default"
degrees = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4, 5, 6]
CELLS = {1: 200, 2: 112, 3: 80, 4: 56, 5: 48, 6: 42, 7: 32, 8: 28}
ctx = dc.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); fn(); e1.record(stream); e1.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def run(p, name, general, **kw):
    os.environ["BP5_OTF_GENERAL"] = "1" if general else "0"     # honoured by tuning builds (-DBP5_TUNING_ENV) only
    nc = CELLS[p]
    op = dc.PoissonOperator(ctx, dc.make_problem(p, (nc, nc, nc), deformation=1, eps=0.1, **kw))
    n = op.n_owned
    b, x, y = op.initialize_dof_vector(), op.initialize_dof_vector(), op.initialize_dof_vector()
    op.assemble_rhs(b)
    ctl = dc.IterationNumberControl(50, 1e-30)
    op.do_zero_out = False
    x.set(0.0); dc.SolverCGFullMerge(ctl).solve(op, x, b, history=False)
    x.set(0.0)
    t = timed(lambda: dc.SolverCGFullMerge(ctl).solve(op, x, b, history=False))
    op.do_zero_out = True
    for _ in range(2): op.vmult(y, x)
    tv = timed(lambda: [op.vmult(y, x) for _ in range(10)]) / 10
    print(json.dumps(dict(build=label, p=p, case=name, kernel=op.kernel_name, dofs=n, vmult_ms=round(tv * 1e3, 4),
                          vmult_gdofs=round(n / tv / 1e9, 2), cg_gdofs=round(n * ctl.last_step() / t / 1e9, 2),
                          x_l2=x.l2_norm())), flush=True)
    for v in (b, x, y): v.close()
    op.close()


OTF = dict(geometry_mode=dc.GEOM_ON_THE_FLY)
for p in degrees:
    if label == "default":
        run(p, "stored_gll", False, quadrature=dc.QUAD_GLL)
        run(p, "stored_gauss", False)
    run(p, "otf_gll_special", False, quadrature=dc.QUAD_GLL, **OTF)
    run(p, "otf_gll_general", True, quadrature=dc.QUAD_GLL, **OTF)
    run(p, "otf_gauss_general", True, **OTF)
    run(p, "otf_gll_helmholtz_general", True, quadrature=dc.QUAD_GLL, operator_kind=dc.OP_HELMHOLTZ, **OTF)
ctx.close()
