"""CPU: the geometry algorithm of the on-the-fly kernels (csrc/apply_otfg.cuh, apply_otf.cuh), restated in numpy and
checked against the oracle's stored coefficient.  The kernels never see the analytic mapping: they gather the nodal
coordinates of the degree-p mapped cell (MappingQGeneric(p) support points, bp5/step-64.cu:234), interpolate them to
the quadrature points, take the COLLOCATION derivative there, and form
    G = w_q / det(J) adj(J) adj(J)^T   (== JxW J^-1 J^-T, JacobianFunctor, bp5/step-64.cu:84-114)
in the plane order xx, yy, zz, xy, xz, yz (bp5/step-64.cu:108-112).  Test infrastructure only."""
import numpy as np
import pytest

import oracle as O
from oracle import _cell_dofs, _interp3


def _collocation_derivative(xq):
    """D[q][r] = l_r'(x_q), l_r the Lagrange basis through the quadrature points themselves"""
    n = len(xq)
    D = np.zeros((n, n))
    for r in range(n):
        others = [m for m in range(n) if m != r]
        denom = np.prod([xq[r] - xq[m] for m in others])
        for q in range(n):
            D[q, r] = sum(np.prod([xq[q] - xq[m] for m in others if m != k]) for k in others) / denom
    return D


def _on_the_fly_metric(m):
    """what the kernel computes per cell and quadrature point from the nodal coordinates alone"""
    s = O.shape(m.p, m.quad)
    n = m.n
    B, wq = s["B"], s["wq"]
    D = _collocation_derivative(s["xq"])
    X = m.dof_coords()[_cell_dofs(m)]                                   # [cells, n^3, 3] nodal coordinates
    J = np.zeros((m.n_cells, n ** 3, 3, 3))
    xq = np.zeros((m.n_cells, n ** 3, 3))
    for d in range(3):
        v = _interp3(B, X[:, :, d]).reshape(-1, n, n, n)                # values at the q-points, [c, k, j, i]
        xq[:, :, d] = v.reshape(-1, n ** 3)
        J[:, :, d, 0] = np.einsum("ai,nkji->nkja", D, v).reshape(-1, n ** 3)
        J[:, :, d, 1] = np.einsum("bj,nkji->nkbi", D, v).reshape(-1, n ** 3)
        J[:, :, d, 2] = np.einsum("ck,nkji->ncji", D, v).reshape(-1, n ** 3)
    det = np.linalg.det(J)
    adj = np.linalg.inv(J) * det[..., None, None]                      # rows: d xi_d / d x_f times det
    w = (wq[:, None, None] * wq[None, :, None] * wq[None, None, :]).ravel()
    sc = w[None, :] / det
    G = np.empty((6, m.n_cells, n ** 3))
    for pl, (d, e) in enumerate(((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))):
        G[pl] = sc * np.einsum("nqf,nqf->nq", adj[:, :, d, :], adj[:, :, e, :])
    return G, w[None, :] * det, xq


@pytest.mark.parametrize("p", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("quad", [O.GAUSS, O.GLL])
def test_metric_from_nodal_coordinates_equals_stored_metric(p, quad):
    for cells, deform, upper in (((2, 2, 1), 1, None), ((2, 1, 2), 0, (1.5, 2.0, 0.75))):
        m = O.OracleMesh(p, cells, quad=quad, deform=deform, eps=0.1, upper=upper)
        G, jxw, _ = _on_the_fly_metric(m)
        ref = m.metric()
        assert np.abs(G - ref).max() <= 1e-11 * np.abs(ref).max(), (p, quad, deform)
        np.testing.assert_allclose(jxw, m.jxw(), rtol=1e-11)


def test_quadrature_points_from_nodal_coordinates_feed_the_helmholtz_coefficient():
    """a(x_q) = 10 / (0.05 + 2 |x_q|^2) (VaryingCoefficientFunctor, step-64/step-64.cu:100-118) is evaluated at the
    interpolated quadrature points; with collocation they are the nodes themselves"""
    m = O.OracleMesh(4, (2, 1, 1), quad=O.GLL, deform=1, eps=0.1)
    _, _, xq = _on_the_fly_metric(m)
    np.testing.assert_allclose(xq, m.dof_coords()[_cell_dofs(m)], atol=1e-14)
    a = 10.0 / (0.05 + 2.0 * (xq ** 2).sum(axis=2))
    assert a.min() > 0 and a.max() <= 200.0
