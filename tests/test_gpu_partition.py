"""-m gpu: domain-partitioned layout emulated on ONE device: every block of a
Cartesian partition is applied separately with its own [owned | ghost] numbering; ghost values are
filled from / ghost contributions are added to the global vector by global index (what
update_ghost_values / compress(add) do between ranks).  The sum must equal the unpartitioned operator."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p,cells,grid,quad,deform", [(2, (4, 3, 2), (2, 1, 1), 0, 0), (3, (4, 4, 3), (2, 2, 1), 1, 1),
                                                      (4, (4, 2, 4), (2, 2, 2), 0, 1), (5, (3, 4, 2), (3, 2, 1), 1, 0),
                                                      (6, (2, 2, 2), (2, 2, 2), 1, 1), (1, (5, 4, 3), (2, 2, 3), 0, 0)])
def test_partitioned_blocks_sum_to_global_operator(gpu_ctx, p, cells, grid, quad, deform):
    import dealceed_b200 as dc
    import oracle as O
    m = O.OracleMesh(p, cells, quad=quad, deform=deform, eps=0.1)
    u = np.random.default_rng(9).standard_normal(m.n_dofs)
    ref = m.vmult(u)
    bm = m.boundary_mask()
    acc = np.zeros(m.n_dofs)
    owned_seen = np.zeros(m.n_dofs, dtype=int)
    coords_ref = m.dof_coords()
    for coord in itertools.product(*[range(g) for g in grid]):
        op = dc.PoissonOperator(gpu_ctx, dc.make_problem(p, cells, quadrature=quad, deformation=deform, eps=0.1,
                                                         part_grid=grid, part_coord=coord))
        gi = op.global_indices()
        assert len(gi) == op.n_owned + op.n_ghost
        owned_seen[gi[: op.n_owned]] += 1
        np.testing.assert_allclose(op.dof_coordinates(), coords_ref[gi], atol=1e-13)
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        src.import_host(u[gi])                 # owned values + update_ghost_values
        op.cell_loop(dst, src)                 # dst (zero) += local cells' contributions
        np.add.at(acc, gi, dst.to_host(with_ghosts=True))   # compress(add)
        # vmult on the block: zero, cell loop, Dirichlet copy on the owned range
        op.vmult(dst, src)
        loc = dst.to_host(with_ghosts=True)
        own = gi[: op.n_owned]
        assert np.array_equal(loc[: op.n_owned][bm[own]], u[own][bm[own]])
        # right-hand side needs no exchange: owned entries are complete
        b = op.initialize_dof_vector()
        op.assemble_rhs(b)
        np.testing.assert_allclose(b.to_host(), m.rhs()[own], rtol=1e-12, atol=1e-15)
        for v in (src, dst, b):
            v.close()
        op.close()
    assert np.all(owned_seen == 1)             # every dof owned by exactly one block
    acc[bm] = u[bm]
    assert np.linalg.norm(acc - ref) <= 1e-12 * np.linalg.norm(ref)


@pytest.mark.parametrize("p,cells,grid,quad,kind", [(2, (4, 3, 2), (2, 1, 1), 0, 0), (3, (4, 4, 3), (2, 2, 1), 1, 1),
                                                    (4, (4, 2, 4), (2, 2, 2), 0, 1), (5, (3, 4, 2), (3, 2, 1), 1, 0),
                                                    (6, (2, 2, 2), (2, 2, 2), 0, 0)])
def test_partitioned_blocks_with_geometry_on_the_fly(gpu_ctx, p, cells, grid, quad, kind):
    """the same emulation with geometry_mode = BP5_GEOM_ON_THE_FLY on a deformed mesh: the kernels gather the nodal
    coordinates of ghost DoFs through the index tables of the cells on a block's lower faces (both on-the-fly
    kernels, both operators); the blocks must sum to the oracle's global operator"""
    import dealceed_b200 as dc
    import oracle as O
    m = O.OracleMesh(p, cells, quad=quad, deform=1, eps=0.1)
    u = np.random.default_rng(11).standard_normal(m.n_dofs)
    ref = m.vmult(u, kind=kind)
    bm = m.boundary_mask()
    acc = np.zeros(m.n_dofs)
    for coord in itertools.product(*[range(g) for g in grid]):
        op = dc.PoissonOperator(gpu_ctx, dc.make_problem(p, cells, quadrature=quad, operator_kind=kind, deformation=1, eps=0.1,
                                                         part_grid=grid, part_coord=coord,
                                                         geometry_mode=dc.GEOM_ON_THE_FLY))
        assert "on-the-fly" in op.kernel_name
        gi = op.global_indices()
        src, dst = op.initialize_dof_vector(), op.initialize_dof_vector()
        src.import_host(u[gi])
        op.cell_loop(dst, src)
        np.add.at(acc, gi, dst.to_host(with_ghosts=True))
        for v in (src, dst):
            v.close()
        op.close()
    acc[bm] = u[bm]
    assert np.linalg.norm(acc - ref) <= 1e-12 * np.linalg.norm(ref)
