"""BASELINE config 3: BP5 degree sweep p=2..8 over problem sizes ~1M..200M DoFs on one B200.
Per (p, size, quadrature): merged CG, IterationNumberControl(200, 1e-6|b|) like the reference's pcg-merged block
(best of `reps`, CUDA events on the library stream), plus the bare operator (vmult, best of reps x 20).
Writes JSON lines.  usage: python scripts/sweep.py [out.jsonl] [quads=gll,gauss] [degrees=2,...,8]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import dealceed_b200 as dc

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "sweep.jsonl")
quads = sys.argv[2].split(",") if len(sys.argv) > 2 else ["gll", "gauss"]
degrees = [int(a) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else list(range(2, 9))
# cells per direction from SURVEY.md 8(d) config 3 (about 1, 4, 16, 64, 200 M DoFs)
CELLS = {2: (49, 79, 125, 199, 292), 3: (33, 53, 84, 133, 195), 4: (25, 39, 63, 100, 146), 5: (20, 32, 50, 80, 117),
         6: (16, 26, 42, 66, 97), 7: (14, 23, 36, 57, 83), 8: (12, 20, 31, 50, 73)}
HBM = 6548.2
ctx = dc.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
reps = 3
with open(out_path, "w") as f:
    for p in degrees:
        for nc in CELLS[p]:
            for qn in quads:
                op = dc.PoissonOperator(ctx, dc.make_problem(p, (nc,) * 3, quadrature=dc.QUAD_GLL if qn == "gll" else dc.QUAD_GAUSS))
                n = op.n_owned
                bytes_v, bytes_cg = op.algorithmic_bytes()
                b, x = op.initialize_dof_vector(), op.initialize_dof_vector()
                op.assemble_rhs(b)
                ctl = dc.IterationNumberControl(200, 1e-6 * b.l2_norm())
                op.do_zero_out = False
                best = None
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for rep in range(reps + 1):
                    x.set(0.0)
                    e0.record(stream)
                    dc.SolverCGFullMerge(ctl).solve(op, x, b, history=False)
                    e1.record(stream); e1.synchronize()
                    t = e0.elapsed_time(e1) * 1e-3
                    if rep > 0: best = t if best is None else min(best, t)
                its = ctl.last_step()
                op.do_zero_out = True
                bestv = None
                for rep in range(reps + 1):
                    e0.record(stream)
                    for _ in range(20): op.vmult(x, b)
                    e1.record(stream); e1.synchronize()
                    t = e0.elapsed_time(e1) * 1e-3 / 20
                    if rep > 0: bestv = t if bestv is None else min(bestv, t)
                rec = dict(p=p, quad=qn, cells=nc, dofs=n, its=its, cg_gdofs=n * its / best / 1e9, cg_ms_per_it=best / its * 1e3,
                           cg_frac=bytes_cg * its / best / 1e9 / HBM, vmult_gdofs=n / bestv / 1e9, vmult_ms=bestv * 1e3,
                           vmult_gbs=bytes_v / bestv / 1e9, vmult_frac=bytes_v / bestv / 1e9 / HBM)
                f.write(json.dumps(rec) + "\n"); f.flush()
                print(json.dumps(rec), flush=True)
                b.close(); x.close(); op.close()
ctx.close()
