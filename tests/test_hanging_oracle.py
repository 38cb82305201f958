"""CPU: the hanging-node restatement (oracle/hanging_oracle.py) -- the `constraint_mask` slot of the reference's
evaluator (bp5/fe_evaluation_gl.h:88,150,167), which no mesh of the reference exercises and no fixture pins.  It is
anchored on the pinned oracle in the two conforming limits and on the mathematics of a conforming space."""
import numpy as np
import pytest
import scipy.sparse.linalg as sla

import oracle as O
from hanging_oracle import HangingMesh


@pytest.mark.parametrize("p", [1, 2, 3])
@pytest.mark.parametrize("quad", [O.GAUSS, O.GLL])
def test_conforming_limits_reproduce_the_pinned_oracle(p, quad):
    cells = (3, 2, 2)
    rng = np.random.default_rng(p)
    # empty refinement box: the coarse mesh
    m0 = O.OracleMesh(p, cells, quad=quad, deform=1, eps=0.1)
    h0 = HangingMesh(p, cells, (0, 0, 0), (0, 0, 0), quad=quad, deform=1, eps=0.1)
    assert h0.n_dofs == m0.n_dofs and np.abs(h0.dof_coords() - m0.dof_coords()).max() <= 1e-14
    u = rng.standard_normal(m0.n_dofs)
    for kind in (O.POISSON, O.HELMHOLTZ):
        ref = m0.vmult(u, kind=kind)
        assert np.linalg.norm(h0.vmult(u, kind) - ref) <= 1e-12 * np.linalg.norm(ref)
    assert np.linalg.norm(h0.rhs() - m0.rhs()) <= 1e-13 * np.linalg.norm(m0.rhs())
    assert h0.l2_norm(u) == pytest.approx(m0.l2_norm(u), rel=1e-12)
    # box = whole domain: the mesh with twice the cells (no interior interface, no hanging node)
    m1 = O.OracleMesh(p, tuple(2 * c for c in cells), quad=quad, deform=1, eps=0.1, upper=tuple(float(c) for c in cells))
    h1 = HangingMesh(p, cells, (0, 0, 0), cells, quad=quad, deform=1, eps=0.1)
    assert h1.n_dofs == m1.n_dofs and h1.n_cells == 8 * np.prod(cells)
    assert np.abs(h1.dof_coords() - m1.dof_coords()).max() <= 1e-14
    u = rng.standard_normal(m1.n_dofs)
    ref = m1.vmult(u)
    assert np.linalg.norm(h1.vmult(u) - ref) <= 1e-12 * np.linalg.norm(ref)


@pytest.mark.parametrize("p,lo,hi", [(1, (1, 1, 1), (3, 2, 2)), (2, (1, 0, 1), (3, 2, 2)), (3, (0, 0, 0), (2, 2, 1)),
                                     (4, (1, 1, 0), (2, 2, 2)), (2, (1, 1, 1), (2, 2, 2))])
def test_constrained_space_is_conforming(p, lo, hi):
    """a harmonic polynomial of degree <= p lies in the constrained space; then (A u)_i = int grad u . grad phi_i = 0
    for every free interior DoF exactly when phi_i is continuous across the coarse-fine faces (affine mesh, Gauss
    quadrature exact).  Wrong hanging-node weights or slots leave O(1) residuals on the interface."""
    hm = HangingMesh(p, (4, 3, 3), lo, hi, quad=O.GAUSS, upper=(1., 1., 1.))
    x, y, z = hm.dof_coords().T
    harmonic = {1: x + 2 * y - z + x * y, 2: x * x - y * y + x * z, 3: x ** 3 - 3 * x * y * y + y * z,
                4: x ** 4 - 6 * x * x * y * y + y ** 4 + x * y * z}
    A = hm.matrix()
    interior = ~hm.boundary_mask()
    for d in range(1, p + 1):
        assert np.abs((A @ harmonic[d])[interior]).max() <= 1e-12 * np.abs(A).max() * np.abs(harmonic[d]).max()
    assert abs(A - A.T).max() <= 1e-13 * np.abs(A).max()
    assert np.abs(A @ np.ones(hm.n_dofs)).max() <= 1e-12 * np.abs(A).max()      # constants are in the kernel
    # the refinement adds DoFs: strictly between the coarse and the globally refined count
    n_coarse = np.prod([c * p + 1 for c in (4, 3, 3)]); n_fine = np.prod([2 * c * p + 1 for c in (4, 3, 3)])
    assert n_coarse < hm.n_dofs < n_fine


@pytest.mark.parametrize("p", [2, 3])
def test_manufactured_solution_converges_with_order_p_plus_1_on_refined_corner(p):
    u = lambda x: np.prod(np.sin(np.pi * x), axis=-1)
    f = lambda x: 3 * np.pi ** 2 * u(x)
    errs = []
    for c in (2, 4, 8):
        hm = HangingMesh(p, (c, c, c), (0, 0, 0), (c // 2,) * 3, quad=O.GAUSS, upper=(1., 1., 1.), deform=1, eps=0.05)
        free = ~hm.boundary_mask()
        A = hm.matrix()[free][:, free].tocsc()
        xs = np.zeros(hm.n_dofs)
        xs[free] = sla.spsolve(A, hm.rhs(f)[free])
        errs.append(hm.l2_error(xs, u))
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert rates[1] >= p + 1 - 0.3, (errs, rates)


def test_facade_counts_dofs_and_cells_of_a_locally_refined_mesh(tmp_path):
    """DoFHandler::n_dofs / Triangulation::n_global_active_cells of the header-only facade (no CUDA needed)"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "probe.cc"
    src.write_text('#include <iostream>\n#include "dealii_b200/dealii_b200.h"\n'
                   'int main(int, char **argv) { using namespace dealii; Triangulation<3> t; Point<3> p2; p2[0]=p2[1]=p2[2]=1.;\n'
                   '  GridGenerator::subdivided_hyper_rectangle(t, {4u, 3u, 3u}, Point<3>(), p2);\n'
                   '  t.refine_cells_in_box({1u, 0u, 1u}, {3u, 2u, 2u}); DoFHandler<3> dh(t); dh.distribute_dofs(FE_Q<3>(std::atoi(argv[1])));\n'
                   '  std::cout << dh.n_dofs() << " " << t.n_global_active_cells() << std::endl; return 0; }\n')
    exe = tmp_path / "probe"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    for p in (1, 2, 5):
        n_dofs, n_cells = [int(v) for v in subprocess.check_output([str(exe), str(p)], text=True).split()]
        hm = HangingMesh(p, (4, 3, 3), (1, 0, 1), (3, 2, 2), upper=(1., 1., 1.)) if p < 5 else None
        if hm is not None:
            assert (n_dofs, n_cells) == (hm.n_dofs, hm.n_cells)
        assert n_cells == 36 - 4 + 32


def test_hanging_oracle_reproduces_its_committed_fixture():
    """tests/golden/hanging_cases.npz (scripts/make_golden_hanging.py): regression pin of the restatement itself"""
    import os
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "hanging_cases.npz"))
    for key in sorted(k[:-5] for k in gold.files if k.endswith("_spec")):
        p, quad = int(key[1]), int(key[4])
        spec = gold[key + "_spec"]
        eps = float(gold[key + "_eps"])
        hm = HangingMesh(p, spec[0:3], spec[3:6], spec[6:9], quad=quad, upper=(1., 1., 1.), deform=1 if eps else 0, eps=eps)
        assert (hm.n_dofs, hm.n_cells) == tuple(gold[key + "_n"])
        u = np.random.default_rng(p).standard_normal(hm.n_dofs)
        assert np.linalg.norm(hm.vmult(u) - gold[key + "_Au"]) <= 1e-13 * np.linalg.norm(gold[key + "_Au"])
        assert np.linalg.norm(hm.rhs() - gold[key + "_b"]) <= 1e-13 * np.linalg.norm(gold[key + "_b"])
