"""CPU: the assumption behind the even-odd contractions of the cell kernels (csrc/apply.cuh, eo_matvec): every 1D
matrix of the path obeys M[n-1-i][n-1-m] = s * M[i][m], s = +1 for values, -1 for derivatives, because the FE_Q
support points and both quadrature rules are symmetric about the cell centre; and the packed E | O | C form
reproduces the full matrix-vector product."""
import numpy as np
import pytest

import oracle as O


def collocation_derivative(xq):
    """D[q][r] = l_r'(x_q), l_r the Lagrange basis through the points xq"""
    n = len(xq)
    D = np.zeros((n, n))
    for q in range(n):
        for r in range(n):
            if q == r:
                D[q, r] = sum(1.0 / (xq[q] - xq[k]) for k in range(n) if k != q)
            else:
                num = np.prod([xq[q] - xq[k] for k in range(n) if k not in (q, r)])
                den = np.prod([xq[r] - xq[k] for k in range(n) if k != r])
                D[q, r] = num / den
    return D


def eo_matvec(M, v, s):
    """the device algorithm, in numpy"""
    n = len(v); H, H1 = n // 2, (n + 1) // 2
    E = 0.5 * (M[:H1, :H] + M[:H1, ::-1][:, :H]); Od = 0.5 * (M[:H1, :H] - M[:H1, ::-1][:, :H])
    C = M[:H1, H] if n % 2 else np.zeros(H1)
    e = v[:H] + v[::-1][:H]; o = v[:H] - v[::-1][:H]
    w = np.zeros(n)
    for i in range(H):
        pe = E[i] @ e + (C[i] * v[H] if n % 2 else 0.0); po = Od[i] @ o
        w[i] = pe + po
        w[n - 1 - i] = pe - po if s > 0 else po - pe
    if n % 2:
        w[H] = (E[H] @ e + C[H] * v[H]) if s > 0 else Od[H] @ o
    return w


@pytest.mark.parametrize("p", range(1, 9))
@pytest.mark.parametrize("quad", [O.GAUSS, O.GLL])
def test_matrices_are_centro_symmetric_and_even_odd_form_is_exact(p, quad):
    t = O.shape(p, quad)
    B, xq = t["B"], t["xq"]
    D = collocation_derivative(xq)
    flip = lambda M: M[::-1, ::-1]
    assert np.abs(flip(B) - B).max() <= 1e-13
    assert np.abs(flip(t["Dg"]) + t["Dg"]).max() <= 1e-11 * max(1.0, np.abs(t["Dg"]).max())
    assert np.abs(flip(D) + D).max() <= 1e-11 * max(1.0, np.abs(D).max())
    v = np.random.default_rng(p).standard_normal(p + 1)
    for M, s in ((B, 1), (B.T, 1), (D, -1), (D.T, -1)):
        ref = M @ v
        assert np.abs(eo_matvec(M, v, s) - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max())
