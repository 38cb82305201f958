"""Quick throughput probe (GPU box): vmult and merged CG per degree, CUDA-event timed on the library's stream."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import dealceed_b200 as dc

target = float(sys.argv[1]) if len(sys.argv) > 1 else 30e6
degrees = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(2, 9))
quads = [int(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]
ctx = dc.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
HBM = 6548.2
for p in degrees:
    nc = max(2, round((target ** (1 / 3) - 1) / p))
    for quad in quads:
        geom = int(os.environ.get('PROBE_GEOM', '0'))         # 1: geometry on the fly (gll only)
        eps = float(os.environ.get('PROBE_EPS', '0'))
        if geom and quad != 1:
            continue
        op = dc.PoissonOperator(ctx, dc.make_problem(p, (nc, nc, nc), quadrature=quad, geometry_mode=geom,
                                                     deformation=1 if eps else 0, eps=eps,
                                                     cell_order=int(os.environ.get('PROBE_COLORED', '0')),
                                                     **(dict(refine_lo=(0, 0, 0), refine_hi=((nc,) * 3 if os.environ.get('PROBE_REFINE') == 'full' else (nc // 2,) * 3))
                                                        if os.environ.get('PROBE_REFINE') else {})))
        n = op.n_owned
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        src.import_host(np.random.default_rng(0).standard_normal(n))
        bytes_v, bytes_cg = op.algorithmic_bytes()
        for _ in range(3): op.vmult(dst, src); op.cell_loop(dst, src)
        ctx.synchronize()
        reps = int(os.environ.get('PROBE_REPS', '10'))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps): op.vmult(dst, src)
        e1.record(stream); e1.synchronize()
        t = e0.elapsed_time(e1) / reps * 1e-3
        # cell loop only (no zeroing / constrained copy)
        e0.record(stream)
        for _ in range(reps): op.cell_loop(dst, src)
        e1.record(stream); e1.synchronize()
        tc = e0.elapsed_time(e1) / reps * 1e-3
        # merged CG, fixed 40 iterations
        b, x = op.initialize_dof_vector(), op.initialize_dof_vector()
        op.assemble_rhs(b)
        op.do_zero_out = False
        ctl = dc.IterationNumberControl(40, 0.0)
        dc.SolverCGFullMerge(ctl).solve(op, x, b, history=False)
        x.set(0.0)
        e0.record(stream)
        dc.SolverCGFullMerge(ctl).solve(op, x, b, history=False)
        e1.record(stream); e1.synchronize()
        tcg = e0.elapsed_time(e1) * 1e-3 / ctl.last_step()
        print(json.dumps(dict(p=p, quad="gauss" if quad == 0 else "gll", cells=nc, dofs=n, kernel=op.kernel_name,
                              vmult_ms=round(t * 1e3, 4), vmult_gdofs=round(n / t / 1e9, 3), vmult_gbs=round(bytes_v / t / 1e9, 1),
                              vmult_frac=round(bytes_v / t / 1e9 / HBM, 3),
                              cellloop_ms=round(tc * 1e3, 4), cellloop_gbs=round(bytes_v / tc / 1e9, 1),
                              cg_ms_per_it=round(tcg * 1e3, 4), cg_gdofs=round(n / tcg / 1e9, 3), cg_gbs=round(bytes_cg / tcg / 1e9, 1),
                              cg_frac=round(bytes_cg / tcg / 1e9 / HBM, 3))), flush=True)
        for v in (src, dst, b, x): v.close()
        op.close()
ctx.close()
