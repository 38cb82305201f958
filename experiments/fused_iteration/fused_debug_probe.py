"""Tuning aid (GPU box): per-activity clock ticks of the fused kernel's CTA 0, for a few BP5_FUSE_DEBUG masks.
Needs a -DBP5_FZ_DEBUG build of the library (BP5_LIB=...)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import dealceed_b200 as dc
p = int(os.environ.get("P", "6")); nc = int(os.environ.get("NC", "88"))
ctx = dc.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
op = dc.PoissonOperator(ctx, dc.make_problem(p, (nc,) * 3, quadrature=1))
src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
src.import_host(np.random.default_rng(0).standard_normal(op.n_owned))
lib = dc.bindings.lib()
for _ in range(2): op.vmult(dst, src)
ctx.synchronize()
lib.bp5_debug_fused_ticks(op.h)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(5): op.vmult(dst, src)
e1.record(stream); e1.synchronize()
print("vmult ms", e0.elapsed_time(e1) / 5, "env", {k: v for k, v in os.environ.items() if k.startswith("BP5_FUSE")}, file=sys.stderr)
lib.bp5_debug_fused_ticks(op.h)
b, x = op.initialize_dof_vector(), op.initialize_dof_vector()
op.assemble_rhs(b); op.do_zero_out = False
ctl = dc.IterationNumberControl(20, 0.0)
dc.SolverCGFullMerge(ctl).solve(op, x, b, history=False)
lib.bp5_debug_fused_ticks(op.h)
x.set(0.0)
e0.record(stream)
dc.SolverCGFullMerge(ctl).solve(op, x, b, history=False)
e1.record(stream); e1.synchronize()
print("cg ms/it", e0.elapsed_time(e1) / 20, file=sys.stderr)
lib.bp5_debug_fused_ticks(op.h)
