"""torchrun parity check of the partitioned path on N GPUs against the oracle (small mesh):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py
Prints one OK/FAIL line per check on rank 0; exit code 1 on failure."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, torch.distributed as dist
import dealceed_b200 as dc
from dealceed_b200.distributed import DistributedPoisson
import oracle as O

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local_rank = int(os.environ.get("LOCAL_RANK", rank))
# CHECK_SAME_DEVICE=1: every rank uses cuda:0 (a 1-GPU box still exercises the peer transport: CUDA IPC works
# between processes on one device; the processes time-slice, so the flag waits are slow but finite).  NCCL
# refuses two ranks on one device, so the plumbing (handle exchange, scalar sums of the checks) is gloo then and
# only the peer transport runs.
SAME_DEVICE = os.environ.get("CHECK_SAME_DEVICE", "0") == "1"
if SAME_DEVICE:
    local_rank = 0
torch.cuda.set_device(local_rank)
if SAME_DEVICE:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
fails = 0
CASES = [(2, (3, 2, 2), dc.QUAD_GAUSS, 0, 0), (4, (2, 2, 3), dc.QUAD_GLL, 1, 0), (6, (2, 2, 2), dc.QUAD_GLL, 1, 0),
         (5, (3, 3, 2), dc.QUAD_GAUSS, 0, 0), (5, (2, 3, 2), dc.QUAD_GLL, 1, 1)]      # last: geometry on the fly
# CHECK_TRANSPORTS=peer,nccl  CHECK_CASES=0,1,4  restrict the run (large boxes are charged per GPU)
TRANSPORTS = os.environ.get("CHECK_TRANSPORTS", "peer" if SAME_DEVICE else "peer,nccl").split(",")
if "CHECK_CASES" in os.environ:
    CASES = [CASES[int(i)] for i in os.environ["CHECK_CASES"].split(",")]
for transport, (p, cpg, quad, deform, geom) in [(t, c) for t in TRANSPORTS for c in CASES]:
    P = DistributedPoisson(p, cpg, quadrature=quad, deformation=deform, eps=0.1, device=local_rank, transport=transport,
                           geometry_mode=geom)
    cells = P.part.cells
    m = O.OracleMesh(p, cells, quad=quad, deform=deform, eps=0.1)
    gi = P.op.global_indices()
    own = gi[: P.op.n_owned]
    u = np.random.default_rng(5).standard_normal(m.n_dofs)
    src, dst = P.op.initialize_dof_vector(), P.op.initialize_dof_vector()
    full = np.zeros(P.op.n_owned + P.op.n_ghost); full[: P.op.n_owned] = u[own]
    src.import_host(full)
    P.vmult(dst, src)
    ref = m.vmult(u)
    err = np.linalg.norm(dst.to_host() - ref[own]) ** 2
    tot = P.allreduce_scalar(err)
    rel = np.sqrt(tot) / np.linalg.norm(ref)
    # merged CG
    b, x = P.op.initialize_dof_vector(), P.op.initialize_dof_vector()
    P.op.assemble_rhs(b)
    bo = m.rhs()
    tol = 1e-8 * np.linalg.norm(bo)
    ctl = dc.SolverControl(500, tol)
    P.cg_solve(x, b, ctl, poll_every=3, history=True)
    xo, its, res, hist, ok = m.cg(bo, variant=1, control=1, tol=tol, max_its=500)
    xerr = np.sqrt(P.allreduce_scalar(np.linalg.norm(x.to_host() - xo[own]) ** 2)) / np.linalg.norm(xo)
    # ghost-value semantics on arbitrary vectors (update_ghost_values / compress(add)), and the host-buffer solve
    n_own, n_gh = P.op.n_owned, P.op.n_ghost
    v = P.op.initialize_dof_vector()
    full = np.full(n_own + n_gh, -1.0); full[:n_own] = gi[:n_own].astype(float)
    v.import_host(full)
    P.update_ghost_values(v)
    ghost_ok = bool(np.array_equal(v.to_host(with_ghosts=True), gi.astype(float)))
    full = np.zeros(n_own + n_gh); full[n_own:] = gi[n_own:].astype(float) + 1.0
    v.import_host(full)
    P.compress_add(v)
    lists = [None] * world
    dist.all_gather_object(lists, gi[n_own:].tolist())
    counts = np.bincount(np.concatenate([np.asarray(l, dtype=np.int64) for l in lists]) if any(lists) else np.zeros(0, dtype=np.int64),
                         minlength=m.n_dofs)
    after = v.to_host(with_ghosts=True)
    ghost_ok = ghost_ok and bool(np.array_equal(after[:n_own], counts[own] * (own.astype(float) + 1.0))) and not after[n_own:].any()
    ghost_ok = bool(P.allreduce_scalar(0.0 if ghost_ok else 1.0) == 0.0)
    v.close()
    host_ok = True
    if P.transport == "peer":
        xh = np.empty(n_own); bhh = b.to_host()
        ctl2 = dc.SolverControl(500, tol)
        P.cg_solve_host(xh, bhh, ctl2)
        herr = np.sqrt(P.allreduce_scalar(np.linalg.norm(xh - xo[own]) ** 2)) / np.linalg.norm(xo)
        host_ok = abs(ctl2.last_step() - its) <= 1 and herr <= 1e-7
    good = rel <= 1e-12 and abs(ctl.last_step() - its) <= 1 and xerr <= 1e-7 and ghost_ok and host_ok
    fails += 0 if good else 1
    if rank == 0:
        print(f"{'OK  ' if good else 'FAIL'} transport={transport} world={world} grid={P.part.grid} p={p} cells={cells} quad={quad} deform={deform} geom={geom}: "
              f"vmult rel err {rel:.2e}, CG its {ctl.last_step()} (oracle {its}), x rel err {xerr:.2e}, "
              f"ghost ops {'ok' if ghost_ok else 'WRONG'}, host solve {'ok' if host_ok else 'WRONG'}", flush=True)
    for v in (src, dst, b, x):
        v.close()
    P.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if fails else 0)
