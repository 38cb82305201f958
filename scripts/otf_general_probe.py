"""Geometry on the fly for the reference's default quadrature and for Helmholtz (general kernel, apply_otfg.cuh) at
the sizes of BASELINE configs 5 and 2 on one B200, next to the stored metric: JSON lines with throughput, iteration
parity and the bytes of geometry data each mode keeps in HBM."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dealceed_b200 as dc
HBM = 6548.2
ctx = dc.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); fn(); e1.record(stream); e1.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def run(name, control_factory, **kw):
    res = dict(config=name)
    for mode, key in ((dc.GEOM_STORED, "stored"), (dc.GEOM_ON_THE_FLY, "on_the_fly")):
        op = dc.PoissonOperator(ctx, dc.make_problem(geometry_mode=mode, **kw))
        n = op.n_owned
        bytes_v, bytes_cg = op.algorithmic_bytes()
        b, x, y = op.initialize_dof_vector(), op.initialize_dof_vector(), op.initialize_dof_vector()
        op.assemble_rhs(b)
        ctl = control_factory(b.l2_norm(), n)
        op.do_zero_out = False
        x.set(0.0); dc.SolverCGFullMerge(ctl).solve(op, x, b, history=False)
        x.set(0.0)
        t = timed(lambda: dc.SolverCGFullMerge(ctl).solve(op, x, b, history=False))
        op.do_zero_out = True
        for _ in range(2): op.vmult(y, x)
        tv = timed(lambda: [op.vmult(y, x) for _ in range(10)]) / 10
        res[key] = dict(kernel=op.kernel_name, dofs=n, geometry_bytes=bytes_v - 16.0 * n, its=ctl.last_step(),
                        residual=ctl.last_value(), cg_gdofs=n * ctl.last_step() / t / 1e9, x_l2=x.l2_norm(),
                        vmult_ms=tv * 1e3, vmult_gdofs=n / tv / 1e9)
        for v in (b, x, y): v.close()
        op.close()
    res["otf_over_stored_cg"] = res["on_the_fly"]["cg_gdofs"] / res["stored"]["cg_gdofs"]
    res["x_rel_diff"] = abs(res["on_the_fly"]["x_l2"] - res["stored"]["x_l2"]) / res["stored"]["x_l2"]
    print(json.dumps(res), flush=True)


it200 = lambda bn, n: dc.IterationNumberControl(200, 1e-6 * bn)
run("config5_p5_deformed_gauss", it200, degree=5, cells=(60, 60, 60), deformation=1, eps=0.1)
run("config5_p5_deformed_gll_helmholtz", it200, degree=5, cells=(60, 60, 60), quadrature=dc.QUAD_GLL,
    operator_kind=dc.OP_HELMHOLTZ, deformation=1, eps=0.1)
run("config2_helmholtz_p4_64_gauss", lambda bn, n: dc.SolverControl(n, 1e-12 * bn), degree=4, cells=(64, 64, 64),
    operator_kind=dc.OP_HELMHOLTZ, upper=(1., 1., 1.))
run("p6_deformed_gauss_42", it200, degree=6, cells=(42, 42, 42), deformation=1, eps=0.1)
ctx.close()
