"""Analytic anchor of the oracle (CPU): manufactured solution u = prod sin(pi x_d) on the unit cube, affine and
smoothly deformed meshes, both quadratures.  The L2 error of the discrete solution must fall like h^(p+1) --
mathematics, not a recalled reference output, pins the discretisation (basis, quadrature, metric, Dirichlet
treatment, CG).  Same integrals as assemble_rhs / integrate_difference (bp5/step-64.cu:372-418, 604-615)."""
import numpy as np
import pytest

import oracle as O

SIZES = {1: (4, 8, 16), 2: (4, 8, 16), 3: (2, 4, 8), 4: (2, 4, 8), 5: (2, 4), 6: (2, 4)}


def observed_orders(p, quad, deform, solve):
    errs = []
    for n in SIZES[p]:
        M = O.Manufactured(p, (n, n, n), deform=deform, eps=0.1)
        b = M.rhs()
        errs.append(M.l2_error(solve(p, n, quad, deform, b)))
    return errs, [float(np.log2(errs[i] / errs[i + 1])) for i in range(len(errs) - 1)]


def check_orders(p, deform, errs, orders):
    # asymptotic rate p+1; the deformed meshes are pre-asymptotic at these sizes (measured 5.5 .. 7.3 at p = 5, 6)
    lo, hi = (p + 1 - 0.2, p + 1 + 0.2) if not deform else (p + 1 - 0.55, p + 1 + 0.45)
    assert lo <= orders[-1] <= hi, (p, deform, errs, orders)
    assert errs[-1] < errs[0]


def oracle_solve(p, n, quad, deform, b):
    m = O.OracleMesh(p, (n, n, n), quad=quad, lower=(0., 0., 0.), upper=(1., 1., 1.), deform=deform, eps=0.1)
    x, its, res, hist, ok = m.cg(b, variant=0, control=1, tol=1e-13 * np.linalg.norm(b), max_its=5000)
    assert ok
    return x


@pytest.mark.parametrize("p", range(1, 7))
@pytest.mark.parametrize("quad", [O.GAUSS, O.GLL])
@pytest.mark.parametrize("deform", [0, 1])
def test_oracle_l2_error_converges_with_order_p_plus_1(p, quad, deform):
    errs, orders = observed_orders(p, quad, deform, oracle_solve)
    check_orders(p, deform, errs, orders)


def test_manufactured_rhs_matches_constant_rhs_integrals():
    """the numpy assembly used above reproduces the oracle's own b_i = int phi_i (f = 1) when fed f = 1"""
    M = O.Manufactured(3, (3, 2, 2), deform=1, eps=0.1)
    m = O.OracleMesh(3, (3, 2, 2), quad=O.GAUSS, lower=(0., 0., 0.), upper=(1., 1., 1.), deform=1, eps=0.1)
    n = M.mesh.n
    loc = np.einsum("ai,bj,ck,ncba->nkji", M.B, M.B, M.B, M.jxw.reshape(-1, n, n, n)).reshape(-1, n ** 3)
    b = np.zeros(m.n_dofs)
    np.add.at(b, M.l2g, loc)
    b[m.boundary_mask()] = 0.0
    np.testing.assert_allclose(b, m.rhs(), rtol=1e-12, atol=1e-15)
