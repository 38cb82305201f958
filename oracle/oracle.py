"""ctypes loader for the CPU oracle (oracle/bp5_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def build(force=False):
    so = os.path.join(_HERE, "libbp5_oracle.so")
    src = os.path.join(_HERE, "bp5_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libbp5_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libbp5_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_mesh_create.restype = C.c_void_p
        L.orc_mesh_create.argtypes = [C.c_int] * 5 + [C.c_double] * 6 + [C.c_int, C.c_double]
        L.orc_mesh_destroy.argtypes = [C.c_void_p]
        L.orc_n_dofs.restype = C.c_int64
        L.orc_n_dofs.argtypes = [C.c_void_p]
        L.orc_n_cells.restype = C.c_int64
        L.orc_n_cells.argtypes = [C.c_void_p]
        L.orc_metric.argtypes = [C.c_void_p, _dp]
        L.orc_jxw.argtypes = [C.c_void_p, _dp]
        L.orc_inv_jacobian.argtypes = [C.c_void_p, _dp]
        L.orc_dof_coords.argtypes = [C.c_void_p, _dp]
        L.orc_boundary_mask.argtypes = [C.c_void_p, C.POINTER(C.c_uint8)]
        L.orc_vmult.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _dp, _dp]
        L.orc_rhs.argtypes = [C.c_void_p, _dp]
        L.orc_l2_norm.restype = C.c_double
        L.orc_l2_norm.argtypes = [C.c_void_p, _dp]
        L.orc_cg.restype = C.c_int
        L.orc_cg.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, _dp, _dp, _dp,
                             C.POINTER(C.c_int), _dp, _dp, C.c_int]
        L.orc_shape.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]
        L.orc_gauss01.argtypes = [C.c_int, _dp, _dp]
        L.orc_lobatto01.argtypes = [C.c_int, _dp, _dp]
        L.orc_set_fast_path.argtypes = [C.c_int]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


GAUSS, GLL = 0, 1
POISSON, HELMHOLTZ = 0, 1


def shape(p, quad=GAUSS):
    n = p + 1
    B = np.zeros((n, n)); Dg = np.zeros((n, n)); xq = np.zeros(n); wq = np.zeros(n); xi = np.zeros(n)
    lib().orc_shape(p, quad, _ptr(B), _ptr(Dg), _ptr(xq), _ptr(wq), _ptr(xi))
    return dict(B=B, Dg=Dg, xq=xq, wq=wq, xi=xi)


class OracleMesh:
    """Structured hex mesh + FE_Q(p) + quadrature + geometry, global lexicographic DoFs."""

    def __init__(self, p, cells, quad=GAUSS, lower=(0., 0., 0.), upper=None, deform=0, eps=0.0):
        self.p, self.n, self.quad = p, p + 1, quad
        self.cells = tuple(int(c) for c in cells)
        if upper is None:
            upper = tuple(float(c) for c in self.cells)  # unit cells, like the reference ladder
        self.lower, self.upper = tuple(lower), tuple(upper)
        self.h = lib().orc_mesh_create(p, quad, *self.cells, *[float(v) for v in self.lower],
                                       *[float(v) for v in self.upper], int(deform), float(eps))
        if not self.h:
            raise ValueError("bad mesh spec")
        self.n_dofs = lib().orc_n_dofs(self.h)
        self.n_cells = lib().orc_n_cells(self.h)
        self.nd = tuple(c * p + 1 for c in self.cells)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_mesh_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def metric(self):
        G = np.zeros((6, self.n_cells, self.n ** 3))
        lib().orc_metric(self.h, _ptr(G))
        return G

    def jxw(self):
        out = np.zeros((self.n_cells, self.n ** 3))
        lib().orc_jxw(self.h, _ptr(out))
        return out

    def inv_jacobian(self):
        out = np.zeros((self.n_cells, 9, self.n ** 3))
        lib().orc_inv_jacobian(self.h, _ptr(out))
        return out

    def dof_coords(self):
        out = np.zeros((self.n_dofs, 3))
        lib().orc_dof_coords(self.h, _ptr(out))
        return out

    def boundary_mask(self):
        out = np.zeros(self.n_dofs, dtype=np.uint8)
        lib().orc_boundary_mask(self.h, out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.astype(bool)

    def vmult(self, src, kind=POISSON, semantics=0, dst=None):
        src = np.ascontiguousarray(src, dtype=np.float64)
        zero = dst is None
        if dst is None:
            dst = np.zeros(self.n_dofs)
        lib().orc_vmult(self.h, kind, semantics, int(zero), _ptr(src), _ptr(dst))
        return dst

    def rhs(self):
        b = np.zeros(self.n_dofs)
        lib().orc_rhs(self.h, _ptr(b))
        return b

    def l2_norm(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        return lib().orc_l2_norm(self.h, _ptr(u))

    def cg(self, b, x0=None, kind=POISSON, variant=0, control=0, tol=0.0, max_its=200, diag=None):
        """returns (x, its, res, history, ok)"""
        x = np.zeros(self.n_dofs) if x0 is None else np.array(x0, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        its = C.c_int(0); res = C.c_double(0.0)
        hist = np.full(max_its + 2, np.nan)
        d = None if diag is None else np.ascontiguousarray(diag, dtype=np.float64)
        rc = lib().orc_cg(self.h, kind, variant, control, float(tol), int(max_its), _ptr(d), _ptr(x), _ptr(b),
                          C.byref(its), C.byref(res), _ptr(hist), len(hist))
        return x, its.value, res.value, hist[: its.value + 1], rc == 0


# ---------------------------------------------------------------------------------------------------
# Manufactured-solution helpers (numpy, test infrastructure): an analytic anchor for the discretisation that
# does not depend on any recalled reference output.  The right-hand side b_i = int phi_i f and the L2 error
# ||u_h - u|| are integrated with QGauss(p+1) on the (possibly deformed) mesh, the way assemble_rhs
# (bp5/step-64.cu:372-418) and integrate_difference (bp5/step-64.cu:604-615) do it in the reference.
def _cell_dofs(mesh):
    """[n_cells, n^3] global lexicographic DoF index of every cell-local DoF (x fastest)."""
    p, n = mesh.p, mesh.n
    nx, ny, nz = mesh.cells
    ndx, ndy, _ = mesh.nd
    i = np.arange(n)
    loc = (i[None, None, :] + ndx * (i[None, :, None] + ndy * i[:, None, None])).ravel()
    cx, cy, cz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    # cells are numbered x fastest
    cxs, cys, czs = [a.transpose(2, 1, 0).ravel() for a in (cx, cy, cz)]
    base = p * (cxs + ndx * (cys + ndy * czs))
    return base[:, None] + loc[None, :]


def _interp3(B, vals):
    """(B x B x B) applied to [cells, n^3] nodal values (x fastest) -> values at the quadrature points."""
    n = B.shape[0]
    v = vals.reshape(-1, n, n, n)                       # [c, k, j, i]
    return np.einsum("ai,bj,ck,nkji->ncba", B, B, B, v).reshape(-1, n ** 3)


class Manufactured:
    """u = prod_d sin(pi x_d) on the unit cube (zero on the boundary; the smooth deformation keeps the cube),
    f = -Laplace u = 3 pi^2 u."""

    def __init__(self, p, cells, deform=0, eps=0.0):
        self.mesh = OracleMesh(p, cells, quad=GAUSS, lower=(0., 0., 0.), upper=(1., 1., 1.), deform=deform, eps=eps)
        sh = shape(p, GAUSS)
        self.B = sh["B"]
        self.l2g = _cell_dofs(self.mesh)
        X = self.mesh.dof_coords()
        self.xq = np.stack([_interp3(self.B, X[:, d][self.l2g]) for d in range(3)], axis=-1)   # [cells, n^3, 3]
        self.jxw = self.mesh.jxw()

    @staticmethod
    def u(x):
        return np.prod(np.sin(np.pi * x), axis=-1)

    def rhs(self):
        m = self.mesh
        n = m.n
        fq = 3.0 * np.pi ** 2 * self.u(self.xq) * self.jxw                      # [cells, n^3]
        loc = np.einsum("ai,bj,ck,ncba->nkji", self.B, self.B, self.B, fq.reshape(-1, n, n, n)).reshape(-1, n ** 3)
        b = np.zeros(m.n_dofs)
        np.add.at(b, self.l2g, loc)
        b[m.boundary_mask()] = 0.0
        return b

    def l2_error(self, uh):
        uq = _interp3(self.B, np.asarray(uh)[self.l2g])
        return float(np.sqrt(np.sum((uq - self.u(self.xq)) ** 2 * self.jxw)))
