"""Oracle vs an independent dense numpy/scipy restatement, plus structural identities (CPU)."""
import numpy as np
import pytest

import oracle as O
from bp5_numpy import NumpyBP5, gauss01, lobatto01, textbook_cg


@pytest.mark.parametrize("p,cells,quad,deform", [(1, (2, 3, 2), "gauss", 0), (2, (2, 2, 2), "gauss", 0), (3, (2, 1, 2), "gll", 0),
                                                 (3, (2, 2, 1), "gauss", 1), (4, (1, 2, 1), "gll", 1), (5, (1, 1, 2), "gauss", 1)])
def test_operator_rhs_norm_match_dense_assembly(p, cells, quad, deform):
    m = O.OracleMesh(p, cells, quad=O.GAUSS if quad == "gauss" else O.GLL, deform=deform, eps=0.1)
    nb = NumpyBP5(p, cells, quad=quad, deform=deform, eps=0.1)
    u = np.random.default_rng(0).standard_normal(m.n_dofs)
    for helm in (False, True):
        ref = nb.assemble(helmholtz=helm) @ u
        got = m.vmult(u, kind=int(helm))
        assert np.linalg.norm(ref - got) <= 1e-12 * np.linalg.norm(ref)
    assert np.linalg.norm(nb.rhs() - m.rhs()) <= 1e-12 * np.linalg.norm(m.rhs())
    assert nb.l2_norm(u) == pytest.approx(m.l2_norm(u), rel=1e-6)   # float-rounded cell norms


@pytest.mark.parametrize("n", range(2, 10))
def test_quadrature_rules(n):
    for rule_np, kind in ((gauss01, 0), (lobatto01, 1)):
        x, w = rule_np(n)
        sh = O.shape(n - 1, kind)
        np.testing.assert_allclose(sh["xq"], x, atol=1e-14)
        np.testing.assert_allclose(sh["wq"], w, atol=1e-14)
        deg = 2 * n - 1 if kind == 0 else 2 * n - 3
        for k in range(deg + 1):
            assert np.dot(sh["wq"], sh["xq"] ** k) == pytest.approx(1.0 / (k + 1), abs=1e-13)


@pytest.mark.parametrize("p", [2, 4, 7])
def test_kronecker_form_on_cartesian_cells(p):
    """QGauss(p+1) integrates mass and stiffness exactly on Cartesian cells, so the cell matrix is
    K(x)M(x)M + M(x)K(x)M + M(x)M(x)K."""
    m = O.OracleMesh(p, (1, 1, 1), quad=O.GAUSS, upper=(0.5, 2.0, 1.25))
    sh = O.shape(p, O.GAUSS)
    B, D, w = sh["B"], sh["Dg"], sh["wq"]
    M1 = B.T @ np.diag(w) @ B
    K1 = D.T @ np.diag(w) @ D
    h = [0.5, 2.0, 1.25]
    Mx, My, Mz = [M1 * hh for hh in h]
    Kx, Ky, Kz = [K1 / hh for hh in h]
    A = np.kron(Mz, np.kron(My, Kx)) + np.kron(Mz, np.kron(Ky, Mx)) + np.kron(Kz, np.kron(My, Mx))
    u = np.random.default_rng(1).standard_normal(m.n_dofs)
    bm = m.boundary_mask()
    ref = A @ u
    ref[bm] = u[bm]
    got = m.vmult(u)
    assert np.linalg.norm(ref - got) <= 1e-12 * np.linalg.norm(ref)


@pytest.mark.parametrize("quad", [O.GAUSS, O.GLL])
def test_symmetry_nullspace_and_semantics(quad):
    m = O.OracleMesh(3, (3, 2, 2), quad=quad, deform=1, eps=0.1)
    rng = np.random.default_rng(2)
    bm = m.boundary_mask()
    u, v = rng.standard_normal(m.n_dofs), rng.standard_normal(m.n_dofs)
    u[bm] = 0; v[bm] = 0
    Au, Av = m.vmult(u), m.vmult(v)
    assert abs(v @ Au - u @ Av) <= 1e-12 * abs(v @ Au)
    assert u @ Au > 0
    one = np.ones(m.n_dofs)
    r = m.vmult(one)
    assert np.abs(r[~bm]).max() <= 1e-12 and np.all(r[bm] == 1.0)   # A*1 = 0 away from the constrained rows
    # device semantics [A_ii A_ib; 0 I] vs CPU-MatrixFree semantics [A_ii 0; 0 I] agree when src_b = 0
    w = rng.standard_normal(m.n_dofs)
    assert np.linalg.norm(m.vmult(w, semantics=0) - m.vmult(w, semantics=1)) > 1e-3
    assert np.array_equal(m.vmult(u, semantics=0), m.vmult(u, semantics=1))


def test_cg_matches_scipy_free_textbook_cg():
    nb = NumpyBP5(2, (3, 2, 2), quad="gauss")
    m = O.OracleMesh(2, (3, 2, 2), quad=O.GAUSS)
    A = nb.assemble()
    b = m.rhs()
    tol = 1e-8 * np.linalg.norm(b)
    xr, itr, hr = textbook_cg(A, b, tol, 500)
    x, its, res, hist, ok = m.cg(b, variant=1, control=1, tol=tol, max_its=500)
    assert ok and abs(its - itr) <= 1
    assert np.linalg.norm(x - xr) <= 1e-7 * np.linalg.norm(xr)


def test_timing_variant_of_the_cell_operator_equals_the_general_one():
    """orc_set_fast_path(1) (bench.py's CPU arm only: collocation derivatives with unit-stride loops) changes nothing
    but the speed: <= 1e-13 against the general evaluator every test uses, p = 1..8, deformed mesh"""
    L = O.lib()
    try:
        for p in range(1, 9):
            m = O.OracleMesh(p, (3, 2, 2), quad=O.GLL, deform=1, eps=0.1)
            u = np.random.default_rng(p).standard_normal(m.n_dofs)
            L.orc_set_fast_path(0)
            ref = m.vmult(u)
            L.orc_set_fast_path(1)
            fast = m.vmult(u)
            assert np.linalg.norm(fast - ref) <= 1e-13 * np.linalg.norm(ref)
            # Gauss quadrature and Helmholtz are not affected by the switch
            mg = O.OracleMesh(p, (2, 2, 1), quad=O.GAUSS)
            ug = np.random.default_rng(p).standard_normal(mg.n_dofs)
            a = mg.vmult(ug, kind=O.HELMHOLTZ)
            L.orc_set_fast_path(0)
            assert np.array_equal(a, mg.vmult(ug, kind=O.HELMHOLTZ))
    finally:
        L.orc_set_fast_path(0)
