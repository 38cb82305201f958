"""TEST INFRASTRUCTURE (oracle), not product code: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import it.

Locally refined meshes with hanging nodes -- the `constraint_mask` / `resolve_hanging_nodes` slot of the reference's
evaluator (bp5/fe_evaluation_gl.h:88,150,167), which no mesh of the reference exercises (every mesh there is
`subdivided_hyper_rectangle` + `refine_global`, bp5/step-64.cu:656-663).  PARITY UNPINNED BY THE REFERENCE: deal.II is
not in /root/reference and no fixture covers a non-conforming mesh.  The restatement is pinned instead against
  * the pinned oracle (oracle.py / bp5_oracle.cpp) on the two conforming limits: empty refinement box (the coarse
    mesh) and the box = whole domain (the mesh with twice the cells), to 1e-12;
  * the mathematics: a harmonic polynomial of degree <= p lies in the constrained space, so (A u)_i = 0 at every
    free interior DoF if and only if the constraints make the space conforming; symmetry; order p+1 convergence of
    the manufactured solution.

Mesh: a structured mesh of `cells` coarse cells on the box [lower, upper], whose coarse cells with indices in
[refine_lo, refine_hi) are replaced by their eight children (one level, so the mesh is 2:1 balanced by construction).
FE_Q(p) on every active cell.  The DoFs of a child face that lies on a face of an unrefined neighbour are not degrees
of freedom: they take the values of the neighbour's face polynomial (hanging-node constraints).

This file forms the constraint rows from the full 3D parent-to-child prolongation (a Kronecker product of 1D
interpolation matrices restricted to the hanging rows) and assembles dense cell matrices into a sparse matrix; the CUDA
path applies 1D interpolations face line by face line in shared memory.  Different formulations, same operator.

Numbering (shared with the library so that vectors compare entry by entry; the tests also compare DoF coordinates):
first the nodes of the coarse lattice that belong to at least one unrefined cell, lexicographic (x fastest); then the
nodes of the fine lattice over the refined box that are not hanging, lexicographic."""
import numpy as np
import scipy.sparse as sp

import oracle as O


def lagrange_at(xi, x):
    """[len(x), n] values of the Lagrange basis on the nodes xi at the points x"""
    xi = np.asarray(xi, dtype=np.float64); x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    n = len(xi)
    out = np.ones((len(x), n))
    for i in range(n):
        for m in range(n):
            if m != i:
                out[:, i] *= (x - xi[m]) / (xi[i] - xi[m])
    return out


def lagrange_deriv_at(xi, x):
    """[len(x), n] derivatives of the Lagrange basis on the nodes xi at the points x"""
    xi = np.asarray(xi, dtype=np.float64); x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    n = len(xi)
    out = np.zeros((len(x), n))
    for i in range(n):
        for j in range(n):
            if j == i:
                continue
            term = np.full(len(x), 1.0 / (xi[i] - xi[j]))
            for m in range(n):
                if m != i and m != j:
                    term *= (x - xi[m]) / (xi[i] - xi[m])
            out[:, i] += term
    return out


class HangingMesh:
    def __init__(self, p, cells, refine_lo, refine_hi, quad=O.GAUSS, lower=(0., 0., 0.), upper=None, deform=0, eps=0.0):
        self.p, self.n, self.quad = p, p + 1, quad
        self.cells = np.array(cells, dtype=int)
        self.r0 = np.array(refine_lo, dtype=int); self.r1 = np.array(refine_hi, dtype=int)
        self.lower = np.array(lower, dtype=float)
        self.upper = np.array(upper if upper is not None else cells, dtype=float)
        self.deform, self.eps = int(deform), float(eps)
        self.sh = O.shape(p, quad)
        self.xi = self.sh["xi"]
        n, n3 = self.n, self.n ** 3
        # 1D parent-to-child interpolation: row a = child node a of child s, column = parent node
        self.I1 = [lagrange_at(self.xi, 0.5 * (s + self.xi)) for s in (0, 1)]
        self._matrices = {}
        self._number()
        self._cells()
        self._geometry()

    # ---------------------------------------------------------------- numbering
    def _number(self):
        p, c, r0, r1 = self.p, self.cells, self.r0, self.r1
        refined = np.all(r1 > r0)
        self.nc = c * p + 1
        kz, ky, kx = np.meshgrid(*[np.arange(self.nc[d]) for d in (2, 1, 0)], indexing="ij")
        K = [kx, ky, kz]
        used = np.zeros(kx.shape, dtype=bool)
        # a coarse node is a DoF iff one of the (up to 8) coarse cells around it is unrefined
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    ok = np.ones(kx.shape, dtype=bool); inside = np.ones(kx.shape, dtype=bool)
                    for d, dd in ((0, dx), (1, dy), (2, dz)):
                        # candidate cell index in direction d: floor(k/p) - dd, only distinct when k % p == 0
                        i = K[d] // p - dd
                        valid = (i >= 0) & (i < c[d]) & ((dd == 0) | (K[d] % p == 0))
                        # k = c*p lies in cell c-1 only
                        valid &= ~((dd == 0) & (K[d] == c[d] * p))
                        ok &= valid
                        inside &= (i >= r0[d]) & (i < r1[d])
                    used |= ok & ~(inside & refined)
        self.coarse_id = np.full(kx.shape, -1, dtype=np.int64)
        self.coarse_id[used] = np.arange(used.sum())
        n_coarse = int(used.sum())
        if refined:
            f0 = 2 * r0 * p
            self.nf = 2 * (r1 - r0) * p + 1
            fz, fy, fx = np.meshgrid(*[np.arange(self.nf[d]) for d in (2, 1, 0)], indexing="ij")
            F = [fx, fy, fz]
            hanging = np.zeros(fx.shape, dtype=bool)
            for d in range(3):
                if r0[d] > 0:
                    hanging |= F[d] == 0
                if r1[d] < c[d]:
                    hanging |= F[d] == self.nf[d] - 1
            self.fine_id = np.full(fx.shape, -1, dtype=np.int64)
            self.fine_id[~hanging] = n_coarse + np.arange((~hanging).sum())
            self.f0 = f0
            self.n_dofs = n_coarse + int((~hanging).sum())
        else:
            self.fine_id = None
            self.n_dofs = n_coarse
        self.n_coarse_dofs = n_coarse

    # ---------------------------------------------------------------- active cells, local-to-global with constraints
    def _cells(self):
        p, n, c, r0, r1 = self.p, self.n, self.cells, self.r0, self.r1
        n3 = n ** 3
        a = np.arange(n)
        az, ay, ax = np.meshgrid(a, a, a, indexing="ij")
        ax, ay, az = ax.ravel(), ay.ravel(), az.ravel()     # local node (x fastest)
        refined = np.all(r1 > r0)
        level, index, rows, cols, vals = [], [], [], [], []
        row = 0
        def in_box(i):
            return refined and all(r0[d] <= i[d] < r1[d] for d in range(3))
        # unrefined coarse cells, lexicographic
        for cz in range(c[2]):
            for cy in range(c[1]):
                for cx in range(c[0]):
                    if in_box((cx, cy, cz)):
                        continue
                    ids = self.coarse_id[cz * p + az, cy * p + ay, cx * p + ax]
                    assert ids.min() >= 0
                    rows.append(row + np.arange(n3)); cols.append(ids); vals.append(np.ones(n3))
                    row += n3
                    level.append(0); index.append((cx, cy, cz))
        # children of the refined cells: parents lexicographic, children x fastest
        if refined:
            for cz in range(r0[2], r1[2]):
                for cy in range(r0[1], r1[1]):
                    for cx in range(r0[0], r1[0]):
                        par = (cx, cy, cz)
                        for s in range(8):
                            sd = (s & 1, (s >> 1) & 1, (s >> 2) & 1)
                            Fi = tuple(2 * par[d] + sd[d] for d in range(3))
                            fk = [Fi[d] * p + (ax, ay, az)[d] - self.f0[d] for d in range(3)]
                            ids = self.fine_id[fk[2], fk[1], fk[0]]
                            # hanging rows: the parent's polynomial evaluated at the child's node
                            P3 = np.kron(self.I1[sd[2]], np.kron(self.I1[sd[1]], self.I1[sd[0]]))   # [child, parent]
                            pids = self.coarse_id[par[2] * p + az, par[1] * p + ay, par[0] * p + ax]
                            for loc in range(n3):
                                if ids[loc] >= 0:
                                    rows.append([row + loc]); cols.append([ids[loc]]); vals.append([1.0])
                                else:
                                    w = P3[loc]
                                    nz = np.nonzero(np.abs(w) > 1e-15)[0]
                                    assert pids[nz].min() >= 0, "hanging node depends on a node that is not a DoF"
                                    rows.append(np.full(len(nz), row + loc)); cols.append(pids[nz]); vals.append(w[nz])
                            row += n3
                            level.append(1); index.append(Fi)
        self.level = np.array(level); self.index = np.array(index)
        self.n_cells = len(level)
        self.C = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                               shape=(self.n_cells * n3, self.n_dofs))

    # ---------------------------------------------------------------- geometry and cell matrices
    def _map(self, x):
        if not self.deform:
            return x
        L = self.upper - self.lower
        s = np.prod(np.sin(np.pi * (x - self.lower) / L), axis=-1, keepdims=True)
        return x + self.eps * L * s

    def _geometry(self):
        n, n3 = self.n, self.n ** 3
        B, Dg, wq, xi = self.sh["B"], self.sh["Dg"], self.sh["wq"], self.xi
        h0 = (self.upper - self.lower) / self.cells
        h = h0[None, :] / (2.0 ** self.level)[:, None]                         # [cells, 3]
        a = np.arange(n)
        az, ay, ax = np.meshgrid(a, a, a, indexing="ij")
        ref = np.stack([xi[ax.ravel()], xi[ay.ravel()], xi[az.ravel()]], axis=-1)   # [n3, 3]
        X = self._map(self.lower + h[:, None, :] * (self.index[:, None, :] + ref[None, :, :]))   # nodes [cells, n3, 3]
        Xr = X.reshape(self.n_cells, n, n, n, 3)                              # [c, k, j, i, d]
        # J[d][e] = d x_d / d xi_e at the quadrature points
        J = np.empty((self.n_cells, n3, 3, 3))
        J[..., 0] = np.einsum("ai,bj,ck,nkjid->ncbad", Dg, B, B, Xr).reshape(self.n_cells, n3, 3)
        J[..., 1] = np.einsum("ai,bj,ck,nkjid->ncbad", B, Dg, B, Xr).reshape(self.n_cells, n3, 3)
        J[..., 2] = np.einsum("ai,bj,ck,nkjid->ncbad", B, B, Dg, Xr).reshape(self.n_cells, n3, 3)
        self.xq = np.einsum("ai,bj,ck,nkjid->ncbad", B, B, B, Xr).reshape(self.n_cells, n3, 3)
        det = np.linalg.det(J)
        w3 = (wq[:, None, None] * wq[None, :, None] * wq[None, None, :]).ravel()
        self.jxw = det * w3[None, :]
        Ji = np.linalg.inv(J)                                                  # Ji[e][d] = d xi_e / d x_d
        self.G = np.einsum("nqed,nqfd->nqef", Ji, Ji) * self.jxw[..., None, None]
        self.X = X
        # reference gradients of the n3 local basis functions at the n3 quadrature points: [q, e, i]
        g = np.empty((n3, 3, n3))
        g[:, 0, :] = np.kron(B, np.kron(B, Dg))
        g[:, 1, :] = np.kron(B, np.kron(Dg, B))
        g[:, 2, :] = np.kron(Dg, np.kron(B, B))
        self.gref = g
        self.Bq = np.kron(B, np.kron(B, B))                                    # [q, i]

    def cell_matrices(self, kind=O.POISSON):
        n3 = self.n ** 3
        g2 = self.gref.reshape(3 * n3, n3)                                     # [(q, e), i]
        K = np.empty((self.n_cells, n3, n3))
        for c in range(self.n_cells):
            W = np.einsum("qef,qfj->qej", self.G[c], self.gref).reshape(3 * n3, n3)
            K[c] = g2.T @ W
            if kind == O.HELMHOLTZ:
                r2 = np.sum(self.xq[c] ** 2, axis=-1)
                a = 10.0 / (0.05 + 2.0 * r2)                                   # step-64/step-64.cu:100-118
                K[c] += self.Bq.T @ ((a * self.jxw[c])[:, None] * self.Bq)
        return K

    def matrix(self, kind=O.POISSON):
        """C^T blockdiag(K_cell) C, WITHOUT the Dirichlet rows replaced"""
        if kind not in self._matrices:
            K = self.cell_matrices(kind)
            Kb = sp.block_diag([sp.csr_matrix(K[c]) for c in range(self.n_cells)], format="csr")
            self._matrices[kind] = (self.C.T @ Kb @ self.C).tocsr()
        return self._matrices[kind]

    # ---------------------------------------------------------------- DoF data
    def dof_coords(self):
        """real coordinates of the DoFs (through the mapping of a cell that holds them)"""
        n3 = self.n ** 3
        out = np.full((self.n_dofs, 3), np.nan)
        Cc = self.C.tocoo()
        direct = np.abs(Cc.data - 1.0) < 1e-14
        # rows with a single unit entry are the unconstrained local DoFs
        counts = np.bincount(Cc.row, minlength=self.C.shape[0])
        sel = direct & (counts[Cc.row] == 1)
        Xf = self.X.reshape(-1, 3)
        out[Cc.col[sel]] = Xf[Cc.row[sel]]
        assert not np.isnan(out).any()
        return out

    def boundary_mask(self):
        """DoFs on the boundary of the (undeformed) box; the deformation keeps the boundary"""
        p, c = self.p, self.cells
        m = np.zeros(self.n_dofs, dtype=bool)
        kz, ky, kx = np.meshgrid(*[np.arange(self.nc[d]) for d in (2, 1, 0)], indexing="ij")
        onb = (kx == 0) | (kx == c[0] * p) | (ky == 0) | (ky == c[1] * p) | (kz == 0) | (kz == c[2] * p)
        sel = self.coarse_id >= 0
        m[self.coarse_id[sel]] = onb[sel]
        if self.fine_id is not None:
            fz, fy, fx = np.meshgrid(*[np.arange(self.nf[d]) for d in (2, 1, 0)], indexing="ij")
            F = [fx + self.f0[0], fy + self.f0[1], fz + self.f0[2]]
            onb = np.zeros(fx.shape, dtype=bool)
            for d in range(3):
                onb |= (F[d] == 0) | (F[d] == 2 * c[d] * p)
            sel = self.fine_id >= 0
            m[self.fine_id[sel]] = onb[sel]
        return m

    # ---------------------------------------------------------------- the operator, as the reference's vmult does it
    def vmult(self, src, kind=O.POISSON, A=None):
        """dst = (unconstrained cell loop with hanging-node constraints) src ; dst[c] = src[c] on Dirichlet DoFs
        (bp5/step-64.cu:263-276)"""
        A = self.matrix(kind) if A is None else A
        dst = A @ src
        bm = self.boundary_mask()
        dst[bm] = src[bm]
        return dst

    def rhs(self, f=None):
        """b_i = int f phi_i (f = 1: bp5/step-64.cu:372-418), constrained, Dirichlet rows zero"""
        if self.quad != O.GAUSS:       # the reference integrates the right-hand side with QGauss(p+1) in either mode
            g = HangingMesh(self.p, self.cells, self.r0, self.r1, quad=O.GAUSS, lower=self.lower, upper=self.upper,
                            deform=self.deform, eps=self.eps)
            return g.rhs(f)
        fq = np.ones_like(self.jxw) if f is None else f(self.xq)
        loc = np.einsum("qi,nq->ni", self.Bq, fq * self.jxw).reshape(-1)
        b = self.C.T @ loc
        b[self.boundary_mask()] = 0.0
        return b

    def l2_error(self, uh, u):
        loc = (self.C @ uh).reshape(self.n_cells, -1)
        uq = loc @ self.Bq.T
        return float(np.sqrt(np.sum((uq - u(self.xq)) ** 2 * self.jxw)))

    def l2_norm(self, uh):
        """||u_h||_L2 with QGauss(p+2), per-cell norms rounded to float before they are squared and added, as
        output_results does it in the reference (bp5/step-64.cu:603-615)"""
        n, nq = self.n, self.n + 1
        xg, wg = np.polynomial.legendre.leggauss(nq)
        xg, wg = 0.5 * (xg + 1.0), 0.5 * wg
        B2, D2 = lagrange_at(self.xi, xg), lagrange_deriv_at(self.xi, xg)
        loc = (self.C @ uh).reshape(self.n_cells, n, n, n)
        Xr = self.X.reshape(self.n_cells, n, n, n, 3)
        val = np.einsum("ai,bj,ck,nkji->ncba", B2, B2, B2, loc).reshape(self.n_cells, -1)
        J = np.empty((self.n_cells, nq ** 3, 3, 3))
        J[..., 0] = np.einsum("ai,bj,ck,nkjid->ncbad", D2, B2, B2, Xr).reshape(self.n_cells, -1, 3)
        J[..., 1] = np.einsum("ai,bj,ck,nkjid->ncbad", B2, D2, B2, Xr).reshape(self.n_cells, -1, 3)
        J[..., 2] = np.einsum("ai,bj,ck,nkjid->ncbad", B2, B2, D2, Xr).reshape(self.n_cells, -1, 3)
        w3 = (wg[:, None, None] * wg[None, :, None] * wg[None, None, :]).ravel()
        cell_sq = np.sum(val ** 2 * np.linalg.det(J) * w3[None, :], axis=1)
        cell_norm = np.sqrt(np.maximum(cell_sq, 0.0)).astype(np.float32).astype(np.float64)
        return float(np.sqrt(np.sum(cell_norm ** 2)))

    def cg(self, b, kind=O.POISSON, tol=0.0, max_its=200):
        """textbook CG on the operator of vmult (Dirichlet rows identity), zero start: iteration count and solution"""
        A = self.matrix(kind)
        x = np.zeros_like(b); r = b.copy(); d = r.copy()
        rr = r @ r
        hist = [np.sqrt(rr)]
        it = 0
        while it < max_its and np.sqrt(rr) > tol:
            h = self.vmult(d, kind, A)
            alpha = rr / (d @ h)
            x += alpha * d; r -= alpha * h
            rr_new = r @ r
            d = r + (rr_new / rr) * d
            rr = rr_new
            it += 1
            hist.append(np.sqrt(rr))
        return x, it, np.array(hist)
