"""Domain-partitioned multi-GPU layer: one process per GPU, `torch.distributed` for the plumbing.

Replaces what the reference gets from p4est + Utilities::MPI::Partitioner + CUDA-aware MPI
(bp5/step-64.cu:241,310,349; tests/cuda_aware_mpi.cc) and the per-iteration
MPI_Allreduce of seven doubles (bp5/solver.h:493):

  * Cartesian block partition of the structured mesh, blocks numbered x fastest;
  * interface DoFs are owned by the LOWER block; ghosts are grouped per direction mask
    m = 1..7 (bit d set: the owner is the lower neighbour in dimension d), each group one
    contiguous segment -> a halo message is a plain slice, no unpack on the receiving side of
    update_ghost_values and no pack on the sending side of compress(add);
  * transport "peer" (default): every exchange of the iteration is done by the library's own kernels
    storing into the neighbour's memory over NVLink (CUDA IPC mappings, csrc/peer.cu), the interior
    cells overlap the halo, the seven CG scalars are summed through peer mailboxes, and the whole
    loop is one native call (bp5_peer_cg_solve).  torch.distributed only carries the IPC handles;
  * transport "nccl": update_ghost_values / compress(add) as NCCL send/recv groups and an
    all_reduce of the scalars, driven step by step from here (the baseline the peer path is
    measured against, and the path for blocks that are not in one NVLink domain).

`Partition` is pure host logic (numpy) and is what the CPU `gloo` tests exercise.
"""
import numpy as np

from . import bindings as B


def process_grid(world_size):
    """1x1x1, 2x1x1, 2x2x1, 2x2x2; otherwise the most cubic factorisation."""
    best = None
    for pz in range(1, world_size + 1):
        if world_size % pz:
            continue
        for py in range(pz, world_size // pz + 1):
            if (world_size // pz) % py:
                continue
            px = world_size // (pz * py)
            if px < py:
                continue
            key = (px - pz, px)
            if best is None or key < best[0]:
                best = (key, (px, py, pz))
    return best[1]


class Partition:
    """Index arithmetic of one block; mirrors bp5_operator_create / halo.cu (checked against them
    on the GPU by tests/test_gpu_partition.py)."""

    def __init__(self, degree, cells, grid, coord):
        self.p, self.cells, self.grid, self.coord = degree, tuple(cells), tuple(grid), tuple(coord)
        self.rank = self.rank_of(coord)
        self.c0, self.lc, self.ld, self.has_lo, self.has_hi, self.od = [], [], [], [], [], []
        for d in range(3):
            G, P, c = cells[d], grid[d], coord[d]
            c0 = G * c // P
            lc = G * (c + 1) // P - c0
            self.c0.append(c0); self.lc.append(lc); self.ld.append(lc * degree + 1)
            self.has_lo.append(int(c > 0)); self.has_hi.append(int(c < P - 1))
            self.od.append(lc * degree + 1 - int(c > 0))
        self.n_owned = int(np.prod(self.od))
        self.ghost_offset, self.ghost_size = [0] * 8, [0] * 8
        self.send_offset, self.send_count = [0] * 8, [0] * 8
        goff = soff = 0
        for m in range(1, 8):
            gexists = all(self.has_lo[d] for d in range(3) if m >> d & 1)
            sexists = all(self.has_hi[d] for d in range(3) if m >> d & 1)
            size = int(np.prod([self.od[d] for d in range(3) if not m >> d & 1]))
            self.ghost_offset[m], self.ghost_size[m] = goff, size if gexists else 0
            self.send_offset[m], self.send_count[m] = soff, size if sexists else 0
            goff += self.ghost_size[m]; soff += self.send_count[m]
        self.n_ghost, self.n_send = goff, soff
        self.nd_global = tuple(c * degree + 1 for c in cells)
        self.n_global = int(np.prod(self.nd_global))

    def rank_of(self, coord):
        return coord[0] + self.grid[0] * (coord[1] + self.grid[1] * coord[2])

    def lower(self, m):
        """rank that owns ghost group m (None if the group is empty)"""
        if not self.ghost_size[m]:
            return None
        return self.rank_of(tuple(self.coord[d] - (m >> d & 1) for d in range(3)))

    def upper(self, m):
        """rank whose ghost group m this block feeds (None if none)"""
        if not self.send_count[m]:
            return None
        return self.rank_of(tuple(self.coord[d] + (m >> d & 1) for d in range(3)))

    # ---- numpy index maps (CPU emulation / tests) --------------------------------------------
    def global_indices(self):
        """global lexicographic index of every local dof, owned then ghost groups 1..7"""
        p = self.p
        ax = [np.arange(self.has_lo[d], self.ld[d]) + self.c0[d] * p for d in range(3)]
        lo = [np.array([self.c0[d] * p]) for d in range(3)]
        Nx, Ny = self.nd_global[0], self.nd_global[1]

        def lex(a0, a1, a2):
            k, j, i = np.meshgrid(a2, a1, a0, indexing="ij")
            return (i + Nx * (j + Ny * k)).ravel()

        out = [lex(*ax)]
        for m in range(1, 8):
            if self.ghost_size[m]:
                out.append(lex(*[lo[d] if m >> d & 1 else ax[d] for d in range(3)]))
        return np.concatenate(out)

    def send_indices(self, m):
        """owned indices packed for the upper neighbour in direction m (x fastest)"""
        rng = [np.array([self.od[d] - 1]) if m >> d & 1 else np.arange(self.od[d]) for d in range(3)]
        k, j, i = np.meshgrid(rng[2], rng[1], rng[0], indexing="ij")
        return (i + self.od[0] * (j + self.od[1] * k)).ravel()


class HaloExchange:
    """update_ghost_values / compress(add) over torch.distributed point-to-point ops.
    Works on 1D torch tensors laid out [owned | ghost]; `pack(vec, sendbuf)` and
    `unpack_add(vec, recvbuf)` are supplied by the caller (device kernels or numpy emulation)."""

    def __init__(self, part, make_buffer, group=None):
        import torch.distributed as dist
        self.dist, self.part, self.group = dist, part, group
        self.sendbuf = make_buffer(max(part.n_send, 1))
        self.recvbuf = make_buffer(max(part.n_send, 1))

    def _run(self, ops):
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()

    def update_ghost_values(self, vec, pack):
        part, dist = self.part, self.dist
        if part.n_send:
            pack(vec, self.sendbuf)
        ops = []
        for m in range(1, 8):
            if part.send_count[m]:
                s = self.sendbuf[part.send_offset[m]: part.send_offset[m] + part.send_count[m]]
                ops.append(dist.P2POp(dist.isend, s, part.upper(m), self.group))
            if part.ghost_size[m]:
                o = part.n_owned + part.ghost_offset[m]
                ops.append(dist.P2POp(dist.irecv, vec[o: o + part.ghost_size[m]], part.lower(m), self.group))
        self._run(ops)

    def compress_add(self, vec, unpack_add):
        part, dist = self.part, self.dist
        ops = []
        for m in range(1, 8):
            if part.ghost_size[m]:
                o = part.n_owned + part.ghost_offset[m]
                ops.append(dist.P2POp(dist.isend, vec[o: o + part.ghost_size[m]], part.lower(m), self.group))
            if part.send_count[m]:
                r = self.recvbuf[part.send_offset[m]: part.send_offset[m] + part.send_count[m]]
                ops.append(dist.P2POp(dist.irecv, r, part.upper(m), self.group))
        self._run(ops)
        if part.n_send:
            unpack_add(vec, self.recvbuf)
        if part.n_ghost:
            vec[part.n_owned:].zero_()            # compress() leaves the ghosts zero


class _CudaArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def device_view(torch, ptr, n, device):
    """zero-copy torch tensor over library-owned device memory"""
    return torch.as_tensor(_CudaArray(ptr, n), device=f"cuda:{device}")


class DistributedPoisson:
    """One block of the partitioned BP5 problem on this rank's GPU + the exchanges around it."""

    def __init__(self, degree, cells_per_gpu, quadrature=B.QUAD_GLL, operator_kind=B.OP_POISSON, deformation=0,
                 eps=0.0, device=None, transport="peer", global_cells=None, geometry_mode=B.GEOM_STORED):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.device = torch.cuda.current_device() if device is None else device
        grid = process_grid(self.world)
        coord = (self.rank % grid[0], (self.rank // grid[0]) % grid[1], self.rank // (grid[0] * grid[1]))
        if global_cells is not None:                                      # strong scaling: fixed global mesh
            cells = tuple(int(c) for c in global_cells)
        else:                                                             # weak scaling: fixed block per GPU
            cells = tuple(cells_per_gpu[d] * grid[d] for d in range(3))
        self.part = Partition(degree, cells, grid, coord)
        self.ctx = B.Context(self.device)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream)
        prob = B.make_problem(degree, cells, quadrature=quadrature, operator_kind=operator_kind,
                              deformation=deformation, eps=eps, part_grid=grid, part_coord=coord,
                              geometry_mode=geometry_mode)
        self.op = B.PoissonOperator(self.ctx, prob)
        assert (self.op.n_owned, self.op.n_ghost) == (self.part.n_owned, self.part.n_ghost)
        # buffers are allocated on torch's default stream (the caching allocator must not tie them to
        # the library's stream, which is destroyed before they are) and only ever used on self.stream
        self.halo = HaloExchange(self.part, lambda n: torch.zeros(n, dtype=torch.float64, device=f"cuda:{self.device}"))
        self.sums = torch.zeros(8, dtype=torch.float64, device=f"cuda:{self.device}")
        torch.cuda.synchronize(self.device)
        self.n_global = self.part.n_global
        self.transport = transport if self.world > 1 else "nccl"
        if self.transport == "peer":
            # all ranks or none: a rank that cannot map its neighbours (no peer access between two devices, IPC
            # disabled in a container) sends everybody to the NCCL transport
            try:
                self._connect_peers()
                ok = 1
            except B.Bp5Error as e:
                ok, self._peer_error = 0, str(e)
            flag = torch.tensor([ok], dtype=torch.int32, device=self._plumbing_device())
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0 and dist.get_backend() != "nccl":
                raise B.Bp5Error(B.ERR_UNSUPPORTED, "peer-memory transport unavailable and the process group is not "
                                 f"NCCL: {getattr(self, '_peer_error', 'a neighbour failed to connect')}")
            if int(flag.item()) == 0:
                if self.rank == 0:
                    import sys
                    print("dealceed_b200: peer-memory transport unavailable, using NCCL "
                          f"({getattr(self, '_peer_error', 'a neighbour failed to connect')})", file=sys.stderr)
                self.transport = "nccl"
        elif self.transport != "nccl":
            raise ValueError(f"unknown transport {transport!r}")

    def _connect_peers(self):
        """publish this block's IPC handles, map everybody else's (bp5_peer_export / bp5_peer_connect)"""
        lib, C, dist = B.lib(), B.C, self.dist
        info = B.PeerInfo()
        try:
            B._check(lib.bp5_peer_export(self.op.h, self.rank, self.world, C.byref(info)))
            blob = bytes(info)
        except B.Bp5Error as e:          # still take part in the gather: nobody may be left waiting
            blob, self._peer_error = None, str(e)
        blobs = [None] * self.world
        dist.all_gather_object(blobs, blob)
        if any(b is None for b in blobs):
            raise B.Bp5Error(B.ERR_UNSUPPORTED, "a rank could not export its peer buffers")
        infos = (B.PeerInfo * self.world)()
        for r, blob in enumerate(blobs):
            C.memmove(C.byref(infos[r]), blob, C.sizeof(B.PeerInfo))
        as_rank = lambda r: -1 if r is None else int(r)
        upper = (C.c_int32 * 8)(*[as_rank(self.part.upper(m)) if m else -1 for m in range(8)])
        lower = (C.c_int32 * 8)(*[as_rank(self.part.lower(m)) if m else -1 for m in range(8)])
        B._check(lib.bp5_peer_connect(self.op.h, infos, upper, lower))
        self.ctx.synchronize()

    # -- helpers ---------------------------------------------------------------------------------
    def _plumbing_device(self):
        """where the few scalars that travel through torch.distributed live: the GPU under NCCL, the host under
        gloo (ranks sharing one device in tests)"""
        return f"cuda:{self.device}" if self.dist.get_backend() == "nccl" else "cpu"

    def view(self, vec):
        return device_view(self.torch, vec.get_values(), vec.n_owned + vec.n_ghost, self.device)

    def _pack(self, vec):
        return lambda t, buf: B._check(B.lib().bp5_operator_halo_pack(self.op.h, vec.h, buf.data_ptr()))

    def _unpack(self, vec):
        return lambda t, buf: B._check(B.lib().bp5_operator_halo_unpack_add(self.op.h, vec.h, buf.data_ptr()))

    def update_ghost_values(self, vec):
        """LinearAlgebra::distributed::Vector::update_ghost_values [UPSTREAM] on any vector of the block's layout"""
        if self.transport == "peer":
            B._check(B.lib().bp5_vector_update_ghost_values(self.op.h, vec.h))
            return
        with self.torch.cuda.stream(self.stream):
            self.halo.update_ghost_values(self.view(vec), self._pack(vec))

    def compress_add(self, vec):
        """compress(VectorOperation::add) [UPSTREAM]; the ghost entries are zero afterwards"""
        if self.transport == "peer":
            B._check(B.lib().bp5_vector_compress_add(self.op.h, vec.h))
            return
        with self.torch.cuda.stream(self.stream):
            self.halo.compress_add(self.view(vec), self._unpack(vec))

    def cg_solve_host(self, x_host, b_host, control, x0_is_zero=True):
        """merged CG with this block's HOST buffers (owned range): b in, x out -- peer transport, one native call"""
        if self.transport != "peer":
            raise B.Bp5Error(B.ERR_UNSUPPORTED, "cg_solve_host needs the peer transport")
        its, val = B.C.c_int(0), B.C.c_double(0.0)
        rc = B.lib().bp5_peer_cg_solve_host(self.op.h, x_host.ctypes.data, b_host.ctypes.data, x_host.size, int(x0_is_zero),
                                            control.kind, control.tol, control.max_its, B.C.byref(its), B.C.byref(val))
        control._last_step, control._last_value = its.value, val.value
        B._check(rc)

    def allreduce_scalar(self, value, op=None):
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self._plumbing_device())
        self.dist.all_reduce(t, op=op or self.dist.ReduceOp.SUM)
        return float(t.item())

    def l2_norm(self, vec):
        return float(np.sqrt(self.allreduce_scalar(vec.dot_local(vec))))

    # -- operator --------------------------------------------------------------------------------
    def vmult(self, dst, src):
        """PoissonOperator::vmult with the exchanges of MatrixFree::cell_loop (bp5/step-64.cu:263-276)."""
        if self.transport == "peer":
            B._check(B.lib().bp5_peer_vmult(self.op.h, dst.h, src.h))
            return
        self.update_ghost_values(src)
        dst.set(0.0)
        self.op.cell_loop(dst, src)
        self.compress_add(dst)
        self.op.copy_constrained_values(dst, src)

    # -- solver ----------------------------------------------------------------------------------
    def cg_solve(self, x, b, control, diag=None, poll_every=10, history=False):
        """SolverCGFullMerge::solve over the partition (x must be zero on entry)."""
        lib, op = B.lib(), self.op
        torch, dist = self.torch, self.dist
        if self.transport == "peer":
            its, val = B.C.c_int(0), B.C.c_double(0.0)
            hist = np.full(control.max_its + 2, np.nan) if history else None
            rc = lib.bp5_peer_cg_solve(op.h, x.h, b.h, diag.h if diag is not None else None, control.kind, control.tol,
                                       control.max_its, B.C.byref(its), B.C.byref(val),
                                       hist.ctypes.data_as(B._dp) if history else None, len(hist) if history else 0)
            control._last_step, control._last_value = its.value, val.value
            control.history = hist[: its.value + 1] if history else None
            B._check(rc)
            return
        # the stepwise path starts from g = -b (bp5_cg_step_begin): a non-zero initial guess would silently give a
        # wrong answer (the reference forms g = A x - b, solver.h:375-381); the peer path rejects it the same way
        if self.allreduce_scalar(0.0 if x.all_zero() else 1.0) != 0.0:
            raise B.Bp5Error(B.ERR_INVALID, "DistributedPoisson.cg_solve needs x == 0 on entry")
        res0 = self.l2_norm(b)
        hist_len = control.max_its + 2 if history else 0
        state = 1 if res0 <= control.tol else (0 if control.max_its > 0 else (1 if control.kind == 0 else 2))
        if state != 0:
            control._last_step, control._last_value = 0, res0
            if state == 2:
                raise B.NoConvergence(B.ERR_NO_CONVERGENCE, "step 0")
            return
        B._check(lib.bp5_cg_step_begin(op.h, x.h, b.h, diag.h if diag is not None else None, control.kind,
                                       control.tol, control.max_its, res0, hist_len))
        g, d, h = B._vp(), B._vp(), B._vp()
        B._check(lib.bp5_cg_step_vectors(op.h, B.C.byref(g), B.C.byref(d), B.C.byref(h)))
        dv, hv = _Borrowed(d, op), _Borrowed(h, op)
        dview, hview = self.view(dv), self.view(hv)
        sums = self.sums
        st, it_done, res = B.C.c_int(0), B.C.c_int(0), B.C.c_double(0.0)
        with torch.cuda.stream(self.stream):
            for it in range(1, control.max_its + 1):
                B._check(lib.bp5_cg_step_update(op.h, it))
                self.halo.update_ghost_values(dview, self._pack(dv))
                B._check(lib.bp5_cg_step_apply_local(op.h))
                self.halo.compress_add(hview, self._unpack(hv))
                B._check(lib.bp5_cg_step_constrained(op.h))
                B._check(lib.bp5_cg_step_local_dots(op.h, sums.data_ptr()))
                dist.all_reduce(sums)                                   # MPI_Allreduce, solver.h:493
                B._check(lib.bp5_cg_step_scalars(op.h, sums.data_ptr()))
                if it % poll_every == 0 or it == control.max_its:
                    B._check(lib.bp5_cg_step_poll(op.h, B.C.byref(st), B.C.byref(it_done), B.C.byref(res)))
                    if st.value != 0:
                        break
            hist = np.full(hist_len, np.nan) if history else None
            B._check(lib.bp5_cg_step_finish(op.h, hist.ctypes.data_as(B._dp) if history else None))
        B._check(lib.bp5_cg_step_poll(op.h, B.C.byref(st), B.C.byref(it_done), B.C.byref(res)))
        control._last_step, control._last_value = it_done.value, res.value
        if history:
            hist[0] = res0
            control.history = hist[: it_done.value + 1]
        if st.value == 2:
            raise B.NoConvergence(B.ERR_NO_CONVERGENCE, f"step {it_done.value}, residual {res.value}")
        if st.value == 3:
            raise B.Bp5Error(B.ERR_DIVIDE_BY_ZERO, "d.Ad == 0")

    def close(self):
        self.ctx.synchronize()
        self.torch.cuda.synchronize(self.device)
        if self.transport == "peer":
            self.dist.barrier()     # nobody unmaps while a neighbour may still store into this block
        self.halo = None
        self.sums = None
        self.op.close()
        self.ctx.close()


class _Borrowed:
    """non-owning handle of a library-owned vector (CG work vectors)"""

    def __init__(self, handle, op):
        self.h, self.n_owned, self.n_ghost = handle, op.n_owned, op.n_ghost

    def get_values(self):
        return B.lib().bp5_vector_get_values(self.h)
