"""Independent numpy/scipy restatement of the BP5 / step-64 operator.

TEST INFRASTRUCTURE ONLY (see oracle/bp5_oracle.cpp header).  Written along a
different route from the C++ oracle so that agreement between the two means
something: 1D tables come from numpy's Legendre module (not Newton on a
recurrence), cell matrices are assembled DENSE from explicit 3D shape-function
gradients (no sum factorisation), and the global matrix is a scipy sparse
matrix.  Only practical for small meshes.

Follows: FE_Q on Gauss-Lobatto nodes (bp5/step-64.cu:312), QGauss(p+1) or
QGaussLobatto(p+1) (:243-247), MappingQGeneric(p) (:234), weak form
(grad v, grad u) [+ (v, a u) for step-64, step-64/step-64.cu:154-160],
Dirichlet rows replaced by identity (device semantics of vmult, :275).
"""
import numpy as np
import scipy.sparse as sp
from numpy.polynomial import legendre as L


def gauss01(n):
    x, w = L.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def lobatto01(n):
    N = n - 1
    cN = np.zeros(N + 1); cN[N] = 1.0
    interior = np.sort(L.legroots(L.legder(cN))) if N > 1 else np.array([])
    x = np.concatenate([[-1.0], interior, [1.0]])
    w = 2.0 / (N * (N + 1) * L.legval(x, cN) ** 2)
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_tables(nodes, pts):
    """V[q,a] = phi_a(pts[q]),  D[q,a] = phi_a'(pts[q]) via the Vandermonde route."""
    n = len(nodes)
    # monomial coefficients of each Lagrange polynomial: solve V c = e_a
    V = np.vander(nodes, n, increasing=True)
    coef = np.linalg.solve(V, np.eye(n))           # coef[:, a] of phi_a
    P = np.vander(pts, n, increasing=True)
    val = P @ coef
    dP = np.zeros_like(P)
    for k in range(1, n):
        dP[:, k] = k * pts ** (k - 1)
    der = dP @ coef
    return val, der


class NumpyBP5:
    def __init__(self, p, cells, quad="gauss", lower=(0, 0, 0), upper=None, deform=0, eps=0.0):
        self.p, self.n = p, p + 1
        self.cells = tuple(cells)
        self.lower = np.array(lower, float)
        self.upper = np.array(self.cells if upper is None else upper, float)
        self.deform, self.eps = deform, eps
        n = self.n
        self.xi, _ = lobatto01(n)
        self.xq, self.wq = gauss01(n) if quad == "gauss" else lobatto01(n)
        self.B, self.D = lagrange_tables(self.xi, self.xq)
        self.nd = tuple(c * p + 1 for c in self.cells)
        self.n_dofs = int(np.prod(self.nd))

    def dof(self, gx, gy, gz):
        return gx + self.nd[0] * (gy + self.nd[1] * gz)

    def boundary_mask(self):
        m = np.zeros(self.nd[::-1], bool)
        m[0, :, :] = m[-1, :, :] = True
        m[:, 0, :] = m[:, -1, :] = True
        m[:, :, 0] = m[:, :, -1] = True
        return m.ravel()

    def map_point(self, x):
        if self.deform == 0:
            return x
        s = np.prod(np.sin(np.pi * (x - self.lower) / (self.upper - self.lower)), axis=-1, keepdims=True)
        return x + self.eps * (self.upper - self.lower) * s

    def _tables3d(self, B, D):
        # shape values / reference gradients of all n^3 basis functions at all q-points
        # index order: q = (qz,qy,qx) x fastest, a = (k,j,i) x fastest
        val = np.einsum("ck,bj,ai->cbakji", B, B, B)
        g0 = np.einsum("ck,bj,ai->cbakji", B, B, D)
        g1 = np.einsum("ck,bj,ai->cbakji", B, D, B)
        g2 = np.einsum("ck,bj,ai->cbakji", D, B, B)
        nq = B.shape[0] ** 3; nb = B.shape[1] ** 3
        return val.reshape(nq, nb), np.stack([g.reshape(nq, nb) for g in (g0, g1, g2)], axis=1)  # [q,3,a]

    def cell_nodes(self, cx, cy, cz):
        n = self.n
        h = (self.upper - self.lower) / np.array(self.cells)
        k, j, i = np.meshgrid(range(n), range(n), range(n), indexing="ij")
        ref = np.stack([self.xi[i], self.xi[j], self.xi[k]], axis=-1).reshape(-1, 3)
        x = self.lower + h * (np.array([cx, cy, cz]) + ref)
        return self.map_point(x)                     # [a,3]

    def cell_matrix(self, cx, cy, cz, helmholtz=False, w3=None, val=None, gref=None):
        X = self.cell_nodes(cx, cy, cz)
        J = np.einsum("qea,ad->qde", gref, X)        # J[q,d,e] = dx_d/dxi_e
        det = np.linalg.det(J)
        Jinv = np.linalg.inv(J)                      # Jinv[q,e,d] = dxi_e/dx_d
        gphys = np.einsum("qed,qea->qda", Jinv, gref)  # grad_x phi_a
        jxw = det * w3
        A = np.einsum("q,qda,qdb->ab", jxw, gphys, gphys)
        if helmholtz:
            xq = val @ X
            a = 10.0 / (0.05 + 2.0 * np.sum(xq * xq, axis=1))
            A = A + np.einsum("q,qa,qb->ab", jxw * a, val, val)
        return A, jxw, val

    def assemble(self, helmholtz=False, apply_bc=True):
        n, p = self.n, self.p
        val, gref = self._tables3d(self.B, self.D)
        w3 = np.einsum("c,b,a->cba", self.wq, self.wq, self.wq).ravel()
        rows, cols, vals = [], [], []
        k, j, i = np.meshgrid(range(n), range(n), range(n), indexing="ij")
        for cz in range(self.cells[2]):
            for cy in range(self.cells[1]):
                for cx in range(self.cells[0]):
                    A, _, _ = self.cell_matrix(cx, cy, cz, helmholtz, w3, val, gref)
                    idx = self.dof(cx * p + i, cy * p + j, cz * p + k).ravel()
                    rows.append(np.repeat(idx, n ** 3)); cols.append(np.tile(idx, n ** 3)); vals.append(A.ravel())
        A = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(self.n_dofs, self.n_dofs))
        if apply_bc:
            # device semantics (bp5/step-64.cu:275): rows of constrained DoFs -> identity,
            # columns are kept ([A_ii A_ib; 0 I]).
            bm = self.boundary_mask()
            keep = sp.diags((~bm).astype(float))
            A = keep @ A + sp.diags(bm.astype(float))
        return A.tocsr()

    def rhs(self):
        """b_i = int phi_i with QGauss(p+1), boundary rows zero (bp5/step-64.cu:372-418)."""
        n, p = self.n, self.p
        xq, wq = gauss01(n)
        B, D = lagrange_tables(self.xi, xq)
        val, gref = self._tables3d(B, D)
        w3 = np.einsum("c,b,a->cba", wq, wq, wq).ravel()
        b = np.zeros(self.n_dofs)
        k, j, i = np.meshgrid(range(n), range(n), range(n), indexing="ij")
        for cz in range(self.cells[2]):
            for cy in range(self.cells[1]):
                for cx in range(self.cells[0]):
                    X = self.cell_nodes(cx, cy, cz)
                    J = np.einsum("qea,ad->qde", gref, X)
                    jxw = np.linalg.det(J) * w3
                    idx = self.dof(cx * p + i, cy * p + j, cz * p + k).ravel()
                    np.add.at(b, idx, val.T @ jxw)
        b[self.boundary_mask()] = 0.0
        return b

    def l2_norm(self, u):
        n, p = self.n, self.p
        xq, wq = gauss01(p + 2)
        B, D = lagrange_tables(self.xi, xq)
        val, gref = self._tables3d(B, D)
        w3 = np.einsum("c,b,a->cba", wq, wq, wq).ravel()
        tot = 0.0
        k, j, i = np.meshgrid(range(n), range(n), range(n), indexing="ij")
        for cz in range(self.cells[2]):
            for cy in range(self.cells[1]):
                for cx in range(self.cells[0]):
                    X = self.cell_nodes(cx, cy, cz)
                    J = np.einsum("qea,ad->qde", gref, X)
                    jxw = np.linalg.det(J) * w3
                    idx = self.dof(cx * p + i, cy * p + j, cz * p + k).ravel()
                    uq = val @ u[idx]
                    tot += float(np.float32(np.sqrt(np.sum(uq * uq * jxw)))) ** 2
        return np.sqrt(tot)


def textbook_cg(A, b, tol, max_its):
    """deal.II SolverCG with identity preconditioner; returns x, its, residual history."""
    x = np.zeros_like(b)
    g = -b.copy()
    res = np.linalg.norm(g)
    hist = [res]
    if res <= tol:
        return x, 0, hist
    d = -g
    gh = res * res
    it = 0
    while True:
        it += 1
        h = A @ d
        alpha = gh / (d @ h)
        x += alpha * d
        g += alpha * h
        res = np.linalg.norm(g)
        hist.append(res)
        if res <= tol or it >= max_its:
            break
        beta = gh
        gh = g @ g
        beta = gh / beta
        d = beta * d - g
    return x, it, hist
