"""N>1 leg of bench.py: one rank per GPU (torchrun).  Weak scaling (default): every GPU holds a cells^3
block of a P_x x P_y x P_z partition (1x1x1, 2x1x1, 2x2x1, 2x2x2); strong scaling (--scaling strong): one
cells^3 mesh split over the GPUs.  Halo exchange inside every operator application and one sum of the 7 CG
scalars per iteration, through peer memory over NVLink (--transport peer, default) or NCCL (--transport nccl)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parity_leg(args, dc, DistributedPoisson, local_rank, quad_id):
    """Outside the timed region (like the cpu_baseline leg of N=1): a small smoothly deformed mesh through the SAME
    DistributedPoisson class / transport as the timed solve, checked against the CPU oracle (test infrastructure,
    used here only as the checker): partitioned vmult <= 1e-12 relative L2, merged-CG iteration count +-1 and the
    solution to 1e-7.  Every rank evaluates the oracle for the (small) global mesh and compares its owned range;
    the error sums travel through the class's own scalar allreduce."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    p, cpg = args.degree, (2, 2, 2)
    P = DistributedPoisson(p, cpg, quadrature=quad_id, deformation=1, eps=0.1, device=local_rank, transport=args.transport)
    m = O.OracleMesh(p, P.part.cells, quad=O.GLL if quad_id == dc.QUAD_GLL else O.GAUSS, deform=1, eps=0.1)
    gi = P.op.global_indices()
    own = gi[: P.op.n_owned]
    u = np.random.default_rng(5).standard_normal(m.n_dofs)
    src, dst = P.op.initialize_dof_vector(), P.op.initialize_dof_vector()
    full = np.zeros(P.op.n_owned + P.op.n_ghost); full[: P.op.n_owned] = u[own]
    src.import_host(full)
    P.vmult(dst, src)
    ref = m.vmult(u)
    rel = float(np.sqrt(P.allreduce_scalar(np.linalg.norm(dst.to_host() - ref[own]) ** 2)) / np.linalg.norm(ref))
    b, x = P.op.initialize_dof_vector(), P.op.initialize_dof_vector()
    P.op.assemble_rhs(b)
    bo = m.rhs()
    tol = 1e-8 * float(np.linalg.norm(bo))
    ctl = dc.SolverControl(500, tol)
    P.cg_solve(x, b, ctl, poll_every=3)
    xo, its, res, hist, ok = m.cg(bo, variant=1, control=1, tol=tol, max_its=500)
    xerr = float(np.sqrt(P.allreduce_scalar(np.linalg.norm(x.to_host() - xo[own]) ** 2)) / np.linalg.norm(xo))
    out = {"partition_mesh": f"p={p}, {P.part.cells[0]}x{P.part.cells[1]}x{P.part.cells[2]} cells, deformed (eps 0.1), "
                             f"{P.part.grid[0]}x{P.part.grid[1]}x{P.part.grid[2]} blocks, transport {P.transport}",
           "partition_vmult_rel_err": rel, "partition_cg_its": ctl.last_step(), "oracle_cg_its": int(its),
           "partition_cg_x_rel_err": xerr,
           "partition_ok": bool(rel <= 1e-12 and abs(ctl.last_step() - its) <= 1 and xerr <= 1e-7)}
    for v in (src, dst, b, x):
        v.close()
    P.close()
    return out


def strong_leg(args, dc, DistributedPoisson, local_rank, quad_id, cells, steps, torch, dist, max_its):
    """A short STRONG-scaling measurement on the same ranks: one cells^3 mesh split over all GPUs, merged CG,
    device-timed like the headline (max over ranks).  Reported under "variants" of the N>1 line; the N=1 line of
    the same mesh is the denominator (cells = --cells: the headline's own N=1 workload)."""
    try:
        P = DistributedPoisson(args.degree, (cells,) * 3, quadrature=quad_id, device=local_rank, transport=args.transport,
                               global_cells=(cells,) * 3)
        op = P.op
        b, x = op.initialize_dof_vector(), op.initialize_dof_vector()
        op.assemble_rhs(b)
        control = dc.IterationNumberControl(max_its, 1e-6 * P.l2_norm(b))
        op.do_zero_out = False
        for _ in range(2):
            x.set(0.0); P.cg_solve(x, b, control)
        P.ctx.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(P.stream)
        its = 0
        for _ in range(steps):
            x.set(0.0); P.cg_solve(x, b, control)
            its += control.last_step()
        e1.record(P.stream); e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out = {"workload": f"strong scaling: BP5 p={args.degree}, {cells}^3 cells = {P.n_global} DoFs in total over "
                           f"{P.part.grid[0]}x{P.part.grid[1]}x{P.part.grid[2]} blocks",
               "dofs_global": P.n_global, "dofs_per_gpu": P.n_global / dist.get_world_size(),
               "value": P.n_global * its / float(t.item()) / 1e9, "unit": "GDoF*it/s",
               "ms_per_iteration": float(t.item()) / max(1, its) * 1e3, "x_l2": P.l2_norm(x)}
        b.close(); x.close(); P.close()
        return out
    except Exception as e:                        # a side measurement must not cost the headline line
        return {"error": f"{type(e).__name__}: {e}"}


def run(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import dealceed_b200 as dc
    from dealceed_b200.distributed import DistributedPoisson
    import bench as single

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    quad_ids = {"gll": dc.QUAD_GLL, "gauss": dc.QUAD_GAUSS}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))

    check_partition = parity_leg(args, dc, DistributedPoisson, local_rank, quad_ids[args.quadrature])

    strong = args.scaling == "strong"
    P = DistributedPoisson(args.degree, (args.cells,) * 3, quadrature=quad_ids[args.quadrature], device=local_rank,
                           transport=args.transport, global_cells=(args.cells,) * 3 if strong else None,
                           deformation=1 if args.deformation else 0, eps=args.deformation,
                           geometry_mode=dc.GEOM_ON_THE_FLY if args.geometry == "otf" else dc.GEOM_STORED)
    op, ctx = P.op, P.ctx
    b, x = op.initialize_dof_vector(), op.initialize_dof_vector()
    op.assemble_rhs(b)
    bnorm = P.l2_norm(b)
    control = dc.IterationNumberControl(single.MAX_ITS, 1e-6 * bnorm)
    op.do_zero_out = False

    def solve():
        x.set(0.0)
        P.cg_solve(x, b, control)

    for _ in range(args.warmup):
        solve()
    ctx.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launch_count
    # timed loop = the shipped path (CUDA-graph replay); the per-launch events come from a separate pass below
    sampler = single.ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(P.stream)
    its_total = 0
    for _ in range(args.steps):
        solve()
        its_total += control.last_step()
    e1.record(P.stream)
    e1.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    secs_local = e0.elapsed_time(e1) * 1e-3
    t = torch.tensor([secs_local], dtype=torch.float64, device=f"cuda:{local_rank}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs = float(t.item())
    launches = ctx.launch_count - launches0
    xnorm = P.l2_norm(x)
    last_value = control.last_value()
    # profiled pass: one more solve with a CUDA event pair around every cell-kernel launch on every rank; the
    # roofline line reports the SLOWEST rank's average launch (that rank sets the step time), not rank 0's
    op.profile(True)
    dist.barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(P.stream)
    solve()
    p1.record(P.stream)
    p1.synchronize()
    prof_secs_local = p0.elapsed_time(p1) * 1e-3
    prof_its = control.last_step()
    k_launches, k_ms = op.profile_result()
    op.profile(False)
    clocks = sampler.stop() if sampler else None
    bytes_vmult_local, _ = op.algorithmic_bytes()
    # per operator application: ranks with a lower ghost layer launch the kernel twice (boundary cells, interior cells)
    k_s_local = k_ms * 1e-3 / max(1, prof_its)
    # [launch seconds, achieved GB/s, profiled-pass seconds] of this rank; keep the rank with the longest launch
    mine = torch.tensor([k_s_local, bytes_vmult_local / max(k_s_local, 1e-12) / 1e9, prof_secs_local, float(rank)],
                        dtype=torch.float64, device=f"cuda:{local_rank}")
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    slow = max(allr, key=lambda v: float(v[0]))
    fast = min(allr, key=lambda v: float(v[0]))

    # end to end: pinned host b -> device, solve, x -> pinned host, every step
    n_loc = op.n_owned
    bh = torch.empty(n_loc, dtype=torch.float64, pin_memory=True)
    xh = torch.empty(n_loc, dtype=torch.float64, pin_memory=True)
    bh.numpy()[:] = b.to_host()
    bview, xview = P.view(b), P.view(x)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_its = 0
    e2e_steps = max(1, min(args.steps, 5))
    use_host_entry = P.transport == "peer"
    xnp, bnp = xh.numpy(), bh.numpy()
    for _ in range(e2e_steps):
        if use_host_entry:
            # the C-ABI host-buffer entry point: pinned b -> device, zero initial guess on the device, solve, x -> host
            P.cg_solve_host(xnp, bnp, control, x0_is_zero=True)
        else:
            with torch.cuda.stream(P.stream):
                bview[:n_loc].copy_(bh, non_blocking=True)
            x.set(0.0)                  # zero initial guess made on the device (bp5/step-64.cu:491): nothing to upload
            P.cg_solve(x, b, control)
            with torch.cuda.stream(P.stream):
                xh.copy_(xview[:n_loc], non_blocking=True)
            ctx.synchronize()
        e2e_its += control.last_step()
    dist.barrier()
    e2e_local = time.perf_counter() - t0
    t = torch.tensor([e2e_local], dtype=torch.float64, device=f"cuda:{local_rank}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_secs = float(t.item())

    bytes_vmult, bytes_cg = op.algorithmic_bytes()
    launches_per_it = launches / max(1, its_total)
    variants = {}
    if not args.no_variants and not strong:
        # the weak-scaling vectors go first: the strong-scaling legs need their own blocks
        ctx.synchronize()
        variants["strong_same_mesh_as_n1"] = strong_leg(args, dc, DistributedPoisson, local_rank, quad_ids[args.quadrature],
                                                        args.cells, 2, torch, dist, single.MAX_ITS)
        variants["strong_small_mesh"] = strong_leg(args, dc, DistributedPoisson, local_rank, quad_ids[args.quadrature],
                                                   max(8, args.cells // 2), 3, torch, dist, single.MAX_ITS)
    if rank == 0:
        n_glob = P.n_global
        k_s = float(slow[0])
        ach = float(slow[1])
        grid = P.part.grid
        out = {
            "metric": single.METRIC, "value": n_glob * its_total / secs / 1e9, "unit": single.UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"BP5 Poisson, p={args.degree}, {args.cells}^3 cells {'in total' if strong else 'per GPU'}, "
                            f"{grid[0]}x{grid[1]}x{grid[2]} blocks = "
                            f"{n_glob} DoFs, {args.quadrature} quadrature, merged CG, IterationNumberControl({single.MAX_ITS}, 1e-6|b|), "
                            f"{its_total / args.steps:.0f} iterations per step",
                "quadrature": args.quadrature, "degree": args.degree, "cells_per_gpu": op.n_cells,
                "geometry": "on-the-fly" if args.geometry == "otf" else "stored metric tensor",
                "deformation_eps": args.deformation,
                "dofs_global": n_glob, "dofs_per_gpu": n_glob / world,
                "parallelism": f"domain decomposition {grid[0]}x{grid[1]}x{grid[2]}, " + (
                    "peer-memory halo (P2P stores over NVLink, interior cells overlap) + mailbox sum of 7 doubles/iteration"
                    if P.transport == "peer" else "NCCL halo (send/recv) + allreduce(7 doubles)/iteration"),
                "transport": P.transport,
                "iterations_per_step": its_total / args.steps,
                "l2": "no flush: vectors and metric are far larger than the 126 MB L2", "kernel": op.kernel_name,
            },
            "clocks": clocks,
            "e2e": {"value": n_glob * e2e_its / e2e_secs / 1e9, "unit": single.UNIT,
                    "h2d_bytes_per_step": n_loc * 8 * world, "d2h_bytes_per_step": n_loc * 8 * world,
                    "steps": e2e_steps,
                    "api": ("per rank: bp5_peer_cg_solve_host (pinned host b -> device, zero initial guess, solve, x -> pinned host)"
                            if use_host_entry else
                            "per rank: pinned host b -> device, zero initial guess, DistributedPoisson.cg_solve, x -> host")},
            "gpu_launches": launches, "launches_per_iteration": launches_per_it,
            "variants": variants,
            "roofline": {"bound": "hbm", "kernel": op.kernel_name, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                         "frac": ach / hbm_peak, "traffic": None, "algorithmic_bytes_per_launch": bytes_vmult,
                         "avg_launch_ms": k_s * 1e3, "launches_timed": k_launches, "applications_timed": prof_its,
                         "note": f"slowest rank's cell kernel (rank {int(slow[3])}; interior + boundary launches added per "
                                 f"vmult); fastest rank {int(fast[3])}: {float(fast[0]) * 1e3:.4f} ms. Profiled pass "
                                 "(graph replay off); value / ms_per_step are the graph path",
                         "kernel_share_of_step": float(slow[0]) * prof_its / max(float(slow[2]), 1e-12),
                         "cg_frac_per_gpu": bytes_cg * its_total / secs / 1e9 / hbm_peak,
                         "cg_frac_per_gpu_64B_model": (bytes_cg - 8.0 * n_loc) * its_total / secs / 1e9 / hbm_peak},
            "check": dict({"x_l2": xnorm, "b_l2": bnorm, "last_residual": last_value}, **check_partition),
        }
        print(json.dumps(out))
    # release every torch object that touched the library's stream before that stream goes away
    ctx.synchronize()
    del bview, xview, bh, xh, t
    torch.cuda.synchronize()
    b.close(); x.close()
    P.close()
    dist.destroy_process_group()
    return 0 if check_partition["partition_ok"] else 1
