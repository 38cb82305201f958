"""CPU tests of the N>1 host logic: Partition index arithmetic, and the halo exchange /
allreduce plumbing over a world_size-2 `gloo` group (no GPU).  The per-block cell loop is
emulated with the oracle's bare cell loop on the block's own mesh."""
import itertools
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from dealceed_b200.distributed import HaloExchange, Partition, process_grid


def test_process_grids():
    assert process_grid(1) == (1, 1, 1) and process_grid(2) == (2, 1, 1)
    assert process_grid(4) == (2, 2, 1) and process_grid(8) == (2, 2, 2)
    assert np.prod(process_grid(6)) == 6


@pytest.mark.parametrize("p,cells,grid", [(2, (4, 3, 2), (2, 1, 1)), (3, (4, 4, 4), (2, 2, 2)), (1, (5, 4, 3), (2, 2, 3)),
                                          (4, (3, 2, 2), (3, 2, 1))])
def test_partition_covers_every_dof_once_and_messages_match(p, cells, grid):
    parts = {c: Partition(p, cells, grid, c) for c in itertools.product(*[range(g) for g in grid])}
    owned = np.zeros(parts[(0, 0, 0)].n_global, dtype=int)
    for part in parts.values():
        gi = part.global_indices()
        assert len(gi) == part.n_owned + part.n_ghost
        owned[gi[: part.n_owned]] += 1
    assert np.all(owned == 1)
    by_rank = {part.rank: part for part in parts.values()}
    for part in parts.values():
        gi = part.global_indices()
        for m in range(1, 8):
            if part.send_count[m]:
                up = by_rank[part.upper(m)]
                assert up.lower(m) == part.rank and up.ghost_size[m] == part.send_count[m]
                # what I pack for the upper neighbour is exactly its ghost group m, in order
                mine = gi[part.send_indices(m)]
                o = up.n_owned + up.ghost_offset[m]
                np.testing.assert_array_equal(mine, up.global_indices()[o: o + up.ghost_size[m]])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, p, cells, quad, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grid = process_grid(world)
        coord = (rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1]))
        part = Partition(p, cells, grid, coord)
        gi = part.global_indices()
        glob = O.OracleMesh(p, cells, quad=quad)
        u = np.random.default_rng(3).standard_normal(glob.n_dofs)
        ref = glob.vmult(u, semantics=2)                      # unpartitioned cell loop

        halo = HaloExchange(part, lambda n: torch.zeros(n, dtype=torch.float64))
        src = torch.zeros(part.n_owned + part.n_ghost, dtype=torch.float64)
        src[: part.n_owned] = torch.from_numpy(u[gi[: part.n_owned]])

        def pack(vec, buf):
            for m in range(1, 8):
                if part.send_count[m]:
                    buf[part.send_offset[m]: part.send_offset[m] + part.send_count[m]] = vec[torch.from_numpy(part.send_indices(m))]

        def unpack_add(vec, buf):
            for m in range(1, 8):
                if part.send_count[m]:
                    vec.index_add_(0, torch.from_numpy(part.send_indices(m)),
                                   buf[part.send_offset[m]: part.send_offset[m] + part.send_count[m]])

        halo.update_ghost_values(src, pack)                    # owner -> ghost
        ok_ghost = bool(np.array_equal(src.numpy(), u[gi]))

        # this block's cells as a mesh of their own (unit cells, same spacing), bare cell loop
        lo = tuple(float(c) for c in part.c0)
        hi = tuple(float(part.c0[d] + part.lc[d]) for d in range(3))
        blk = O.OracleMesh(p, part.lc, quad=quad, lower=lo, upper=hi)
        # block-lexicographic index -> local [owned | ghost] index
        pos = {g: i for i, g in enumerate(gi)}
        k, j, i = np.meshgrid(*[np.arange(part.ld[d]) + part.c0[d] * p for d in (2, 1, 0)], indexing="ij")
        lex_global = (i + part.nd_global[0] * (j + part.nd_global[1] * k)).ravel()
        l2l = np.array([pos[g] for g in lex_global])
        dst_lex = blk.vmult(src.numpy()[l2l], semantics=2)
        dst = torch.zeros_like(src)
        dst[torch.from_numpy(l2l)] = torch.from_numpy(dst_lex)
        halo.compress_add(dst, unpack_add)                     # ghost -> owner, add, ghosts zeroed
        own = gi[: part.n_owned]
        err = float(np.linalg.norm(dst.numpy()[: part.n_owned] - ref[own]) / np.linalg.norm(ref))
        ghosts_zero = bool(torch.all(dst[part.n_owned:] == 0))

        # the CG scalars: allreduce of seven partial sums
        sums = torch.tensor([float((src[: part.n_owned] ** 2).sum())] * 7, dtype=torch.float64)
        dist.all_reduce(sums)
        ok_sum = abs(sums[0].item() - float(u @ u)) <= 1e-12 * float(u @ u)
        out[rank] = (ok_ghost, err, ghosts_zero, ok_sum)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p,cells,quad", [(2, (4, 3, 2), O.GAUSS), (3, (3, 2, 2), O.GLL)])
def test_halo_exchange_over_gloo_world_size_2(p, cells, quad):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), p, cells, quad, out), nprocs=world, join=True)
    assert len(out) == world
    for rank in range(world):
        ok_ghost, err, ghosts_zero, ok_sum = out[rank]
        assert ok_ghost and ghosts_zero and ok_sum
        assert err <= 1e-13
