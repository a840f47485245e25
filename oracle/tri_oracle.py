"""CPU ORACLE for H1 Lagrange TRIANGLES (test infrastructure, numpy / scipy only; PARITY UNPINNED like the rest
of oracle/: MFEM is not vendored).  Restates what the reference runs on its shipped Gmsh meshes
(Input/input_2d.yaml:1-2 -> Mesh/unit_square.msh, order 3):

    H1_FECollection(order, 2) on triangles            linear_convection_diffusion_2D.cpp:311-312
    Diffusion + Convection + Mass, a.Assemble()       :335-339
    DomainLFIntegrator, ProjectBdrCoefficient         :341-347
    FormLinearSystem / GMRES                          :351-374   (via pyoracle.CSR)
    ComputeL2Error                                    :383-392

Formulation, deliberately different from the CUDA path (which integrates numerically with dense basis tables):
element matrices are EXACT -- the nodal basis is expanded in monomials, int_T x^a y^b = a! b! / (a + b + 2)!, and
the affine map is applied in closed form:  K_e = kappa |J| sum_ab (J^-1 J^-T)_ab Khat_ab,  C_e = alpha |J| sum_a
(J^-1 c)_a Chat_a,  M_e = s |J| Mhat.  Only forms with non-polynomial data (linear form, error norm) use quadrature.
Node set, native dof order and numbering: [MFEM-upstream, from memory], see host_simplex.cpp.
"""
from math import factorial

import numpy as np
import scipy.sparse as sp
from numpy.polynomial.legendre import leggauss

from . import pyoracle as orc

EDGES = ((0, 1), (1, 2), (2, 0))


def gll(p):
    return orc.gauss_lobatto(p + 1)


def tri_nodes(p):
    cp = gll(p)
    pts = [(cp[0], cp[0]), (cp[p], cp[0]), (cp[0], cp[p])]
    pts += [(cp[i], cp[0]) for i in range(1, p)]
    pts += [(cp[p - i], cp[i]) for i in range(1, p)]
    pts += [(cp[0], cp[p - i]) for i in range(1, p)]
    for j in range(1, p):
        for i in range(1, p - j):
            w = cp[i] + cp[j] + cp[p - i - j]
            pts.append((cp[i] / w, cp[j] / w))
    return np.array(pts)


def monomials(p):
    return [(s - b, b) for s in range(p + 1) for b in range(s + 1)]


def basis_coeffs(p):
    """C[k, j]: coefficient of monomial k in nodal basis function j.  The monomial Vandermonde matrix is mildly
    ill-conditioned (1e5 at p = 4): solved in double, then refined three times with residuals in extended precision, and
    everything downstream is carried in np.longdouble, so that the oracle is good to ~1e-16 and not to cond * eps."""
    X = tri_nodes(p).astype(np.longdouble)
    V = np.array([[x ** a * y ** b for a, b in monomials(p)] for x, y in X], dtype=np.longdouble)
    Vd = V.astype(np.float64)
    C = np.linalg.solve(Vd, np.eye(len(X))).astype(np.longdouble)
    for _ in range(3):
        R = np.eye(len(X), dtype=np.longdouble) - V @ C
        C = C + np.linalg.solve(Vd, R.astype(np.float64)).astype(np.longdouble)
    return C


def eval_basis(p, xy):
    C = basis_coeffs(p)
    xy = np.asarray(xy, dtype=np.longdouble)
    M = np.array([[x ** a * y ** b for a, b in monomials(p)] for x, y in xy], dtype=np.longdouble)
    return (M @ C).astype(np.float64)


def ref_matrices(p):
    """exact reference matrices: M[i,j] = int phi_i phi_j, K[a][b][i,j] = int d_a phi_i d_b phi_j, Cc[a][i,j] = int phi_i d_a phi_j"""
    mono = monomials(p)
    C = basis_coeffs(p)
    nm = len(mono)
    integ = lambda a, b: 0.0 if (a < 0 or b < 0) else factorial(a) * factorial(b) / factorial(a + b + 2)

    def deriv(k, axis):                      # d/dx_axis of monomial k -> (coefficient, (a, b))
        a, b = mono[k]
        return (a, (a - 1, b)) if axis == 0 else (b, (a, b - 1))
    ld = np.longdouble
    integ = lambda a, b: ld(0) if (a < 0 or b < 0) else ld(factorial(a) * factorial(b)) / ld(factorial(a + b + 2))
    Mm = np.array([[integ(mono[k][0] + mono[l][0], mono[k][1] + mono[l][1]) for l in range(nm)] for k in range(nm)], dtype=ld)
    M = (C.T @ Mm @ C).astype(np.float64)
    K = [[None, None], [None, None]]
    Cc = [None, None]
    for a in range(2):
        Ca = np.zeros((nm, nm), dtype=ld)
        for k in range(nm):
            for l in range(nm):
                cl, (la, lb) = deriv(l, a)
                Ca[k, l] = cl * integ(mono[k][0] + la, mono[k][1] + lb) if cl else 0.0
        Cc[a] = (C.T @ Ca @ C).astype(np.float64)
        for b in range(2):
            Kab = np.zeros((nm, nm), dtype=ld)
            for k in range(nm):
                ck, (ka, kb) = deriv(k, a)
                for l in range(nm):
                    cl, (la, lb) = deriv(l, b)
                    Kab[k, l] = ck * cl * integ(ka + la, kb + lb) if (ck and cl) else 0.0
            K[a][b] = (C.T @ Kab @ C).astype(np.float64)
    return M, K, Cc


def tri_rule(n):
    """collapsed Gauss-Legendre rule, n x n points (same definition as host_simplex.cpp, written independently)"""
    x, w = leggauss(n)
    x, w = 0.5 * (x + 1), 0.5 * w
    u, v = np.meshgrid(x, x, indexing="xy")            # u fastest
    wu, wv = np.meshgrid(w, w, indexing="xy")
    return np.stack([(u * (1 - v)).ravel(), v.ravel()], axis=1), (wu * wv * (1 - v)).ravel()


class TriProblem:
    def __init__(self, p, vx, ev, bv, battr, kappa=0.1, vel=(1.0, -2.0), alpha=1.0, mass=1.0, ess_attrs="all"):
        self.p, self.vx, self.ev, self.bv, self.battr = p, np.asarray(vx, float), np.asarray(ev), np.asarray(bv), np.asarray(battr)
        self.kappa, self.vel, self.alpha, self.mass = kappa, vel, alpha, mass
        self.ne, self.nv = len(self.ev), len(self.vx)
        self.nd = (p + 1) * (p + 2) // 2
        self._number()
        marker = set(np.unique(self.battr)) if ess_attrs == "all" else set(ess_attrs)
        mark = np.zeros(self.ndof, np.uint8)
        for b in range(len(self.bv)):
            if self.battr[b] in marker:
                mark[self.bdr_dofs[b]] = 1
        self.ess_mark = mark
        self.ess = np.flatnonzero(mark).astype(np.int32)

    def _number(self):
        p, pm1 = self.p, self.p - 1
        nint = (p - 1) * (p - 2) // 2
        edges = {}
        e_edge = np.zeros((self.ne, 3), np.int64)
        for e in range(self.ne):
            for k, (a, b) in enumerate(EDGES):
                key = tuple(sorted((int(self.ev[e, a]), int(self.ev[e, b]))))
                e_edge[e, k] = edges.setdefault(key, len(edges))
        self.nedges = len(edges)
        off_e, off_i = self.nv, self.nv + self.nedges * pm1
        self.ndof = off_i + self.ne * nint
        g = np.zeros((self.ne, self.nd), np.int32)
        for e in range(self.ne):
            g[e, :3] = self.ev[e]
            o = 3
            for k, (a, b) in enumerate(EDGES):
                base = off_e + e_edge[e, k] * pm1
                fwd = self.ev[e, a] < self.ev[e, b]
                for i in range(pm1):
                    g[e, o] = base + (i if fwd else pm1 - 1 - i)
                    o += 1
            g[e, o:] = off_i + e * nint + np.arange(nint)
        self.elem_dof = g
        self.bdr_dofs = []
        for b in range(len(self.bv)):
            key = tuple(sorted((int(self.bv[b, 0]), int(self.bv[b, 1]))))
            k = edges[key]
            self.bdr_dofs.append(np.concatenate([self.bv[b], off_e + k * pm1 + np.arange(pm1)]).astype(np.int64))

    def jac(self):
        X = self.vx[self.ev]                                    # (ne, 3, 2)
        J = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=2)   # J[e][a][b] = d x_a / d xi_b
        return X, J, np.linalg.det(J)

    def element_matrices(self):
        M, K, Cc = ref_matrices(self.p)
        X, J, det = self.jac()
        Ji = np.linalg.inv(J)
        A = np.zeros((self.ne, self.nd, self.nd))
        if self.kappa is not None:
            G = np.einsum("eab,ecb->eac", Ji, Ji)               # J^-1 J^-T
            for a in range(2):
                for b in range(2):
                    A += (self.kappa * det * G[:, a, b])[:, None, None] * K[a][b][None]
        if self.vel is not None:
            w = np.einsum("eab,b->ea", Ji, np.asarray(self.vel, float))
            for a in range(2):
                A += (self.alpha * det * w[:, a])[:, None, None] * Cc[a][None]
        if self.mass is not None:
            A += (self.mass * det)[:, None, None] * M[None]
        return A

    def csr(self):
        A = self.element_matrices()
        rows = np.repeat(self.elem_dof, self.nd, axis=1).ravel()
        cols = np.tile(self.elem_dof, (1, self.nd)).ravel()
        S = sp.coo_matrix((A.ravel(), (rows, cols)), shape=(self.ndof, self.ndof)).tocsr()
        S.sum_duplicates()
        S.sort_indices()
        return orc.CSR(S.indptr.astype(np.int64), S.indices.astype(np.int32), S.data.astype(np.float64))

    def coords(self):
        X, J, _ = self.jac()
        N = tri_nodes(self.p)
        xe = X[:, 0][:, None, :] + np.einsum("eab,nb->ena", J, N)
        out = np.zeros((self.ndof, 2))
        out[self.elem_dof.ravel()] = xe.reshape(-1, 2)
        return out

    def rule_coords(self, n):
        X, J, _ = self.jac()
        xy, _ = tri_rule(n)
        return X[:, 0][:, None, :] + np.einsum("eab,qb->eqa", J, xy)

    def domain_lf(self, f_q, n, scale=1.0):
        xy, w = tri_rule(n)
        B = eval_basis(self.p, xy)
        _, _, det = self.jac()
        bE = scale * np.abs(det)[:, None] * np.einsum("qi,q,eq->ei", B, w, np.asarray(f_q).reshape(self.ne, -1))
        b = np.zeros(self.ndof)
        np.add.at(b, self.elem_dof.ravel(), bE.ravel())
        return b

    def l2_error(self, u, uex_q, n):
        xy, w = tri_rule(n)
        B = eval_basis(self.p, xy)
        _, _, det = self.jac()
        uh = np.zeros((self.ne, len(w))) if u is None else np.einsum("qi,ei->eq", B, np.asarray(u)[self.elem_dof])
        ex = 0.0 if uex_q is None else np.asarray(uex_q).reshape(self.ne, -1)
        return float(np.sqrt(np.sum(np.abs(det)[:, None] * w[None] * (uh - ex) ** 2)))
