"""Closed-form golden that does NOT come from oracle/: on a rectilinear (axis-aligned,
possibly graded) hex/quad mesh with constant coefficients the assembled operator

    A = kappa K + alpha C(c) + s M     (linear_convection_diffusion_2D.cpp:335-339)

is a Kronecker sum of 1-D matrices,

    K = Kx (x) My (x) Mz + Mx (x) Ky (x) Mz + Mx (x) My (x) Kz,   M = Mx (x) My (x) Mz,
    C = cx Cx (x) My (x) Mz + cy Mx (x) Cy (x) Mz + cz Mx (x) My (x) Cz,

with the 1-D mass / stiffness / convection matrices of Lagrange elements on Gauss-Lobatto
nodes.  Everything here is numpy only: nodes from the roots of P_p', 1-D integrals with
numpy.polynomial.legendre.leggauss, Lagrange values and derivatives from the
product formulas.  The integrands are polynomials of degree <= 2p per direction, so MFEM's
rule (p+2 points in 3-D, p+1 in 2-D) integrates them exactly and the
matrices agree with any exact rule to round-off.

Used by tests/test_oracle.py (oracle vs this) and tests/test_gpu_parity_at_size.py (CUDA vs
this), at sizes up to BASELINE config 2.
"""
import numpy as np
from numpy.polynomial import legendre as L


def gll_nodes(p):
    """p+1 Gauss-Lobatto points on [0,1], ascending"""
    if p == 1:
        return np.array([0.0, 1.0])
    inner = np.sort(np.real(L.Legendre.basis(p).deriv().roots()))
    return 0.5 * (np.concatenate([[-1.0], inner, [1.0]]) + 1.0)


def ref_matrices(p):
    """reference-interval [0,1] matrices m[i,j] = int l_i l_j, k = int l_i' l_j', c = int l_i l_j'"""
    nodes = gll_nodes(p)
    xg, wg = L.leggauss(p + 3)
    xg, wg = 0.5 * (xg + 1.0), 0.5 * wg
    # Lagrange basis and derivative by the product formulas (no Vandermonde: exact to round-off for p <= 6)
    P, dP = np.zeros((len(xg), p + 1)), np.zeros((len(xg), p + 1))
    for j in range(p + 1):
        others = [k for k in range(p + 1) if k != j]
        den = np.prod([nodes[j] - nodes[k] for k in others])
        P[:, j] = np.prod([xg - nodes[k] for k in others], axis=0) / den
        for mm in others:
            rest = [k for k in others if k != mm]
            dP[:, j] += (np.prod([xg - nodes[k] for k in rest], axis=0) if rest else np.ones_like(xg)) / den
    m = P.T @ (wg[:, None] * P)
    k = dP.T @ (wg[:, None] * dP)
    c = P.T @ (wg[:, None] * dP)
    return nodes, m, k, c


def assemble_1d(p, xv):
    """global 1-D matrices (dense, N = p*n + 1) on the vertex coordinates xv, plus the dof coordinates"""
    nodes, m, k, c = ref_matrices(p)
    n = len(xv) - 1
    N = p * n + 1
    M, K, C = np.zeros((N, N)), np.zeros((N, N)), np.zeros((N, N))
    xs = np.zeros(N)
    for e in range(n):
        h = xv[e + 1] - xv[e]
        s = slice(p * e, p * e + p + 1)
        M[s, s] += h * m
        K[s, s] += k / h
        C[s, s] += c
        xs[s] = xv[e] + h * nodes
    return M, K, C, xs


def _banded(A, p):
    """(A as scipy CSR) -- the 1-D matrices have bandwidth p"""
    import scipy.sparse as sp
    return sp.csr_matrix(A)


class KronOperator:
    """y = A x on the lattice of dofs of a rectilinear mesh; `coords` (ndof, dim) maps a caller's dof
    numbering onto the lattice."""

    def __init__(self, p, axes, kappa=None, vel=None, alpha=1.0, mass=None):
        self.p, self.dim = p, len(axes)
        self.mats = [assemble_1d(p, np.asarray(a, dtype=np.float64)) for a in axes]
        self.kappa, self.vel, self.alpha, self.mass = kappa, vel, alpha, mass
        self.shape = tuple(len(m[3]) for m in self.mats)[::-1]          # (Nz, Ny, Nx)
        self._sp = [[_banded(m[i], p) for i in range(3)] for m in self.mats]

    def _along(self, U, d, which):
        """apply the 1-D matrix `which` (0 M, 1 K, 2 C) of direction d (0 = x, fastest lattice axis)"""
        ax = self.dim - 1 - d
        A = self._sp[d][which]
        Um = np.moveaxis(U, ax, 0)
        sh = Um.shape
        out = (A @ Um.reshape(sh[0], -1)).reshape(sh)
        return np.moveaxis(out, 0, ax)

    def apply_lattice(self, U):
        dim = self.dim
        Y = np.zeros_like(U)
        if self.kappa is not None or self.vel is not None:
            for d in range(dim):
                acc_k = self._along(U, d, 1) if self.kappa is not None else None
                acc_c = self._along(U, d, 2) if self.vel is not None else None
                for o in range(dim):
                    if o != d:
                        if acc_k is not None:
                            acc_k = self._along(acc_k, o, 0)
                        if acc_c is not None:
                            acc_c = self._along(acc_c, o, 0)
                if acc_k is not None:
                    Y += self.kappa * acc_k
                if acc_c is not None:
                    Y += self.alpha * self.vel[d] * acc_c
        if self.mass is not None:
            T = U
            for d in range(dim):
                T = self._along(T, d, 0)
            Y += self.mass * T
        return Y

    def lattice_index(self, coords):
        """flat lattice index of every dof from its physical coordinates"""
        idx = np.zeros(coords.shape[0], np.int64)
        stride = 1
        for d in range(self.dim):
            xs = self.mats[d][3]
            i = np.searchsorted(xs, coords[:, d] - 1e-10 * (xs[-1] - xs[0]))
            assert np.all(np.abs(xs[i] - coords[:, d]) < 1e-9 * (xs[-1] - xs[0])), "dof not on the lattice"
            idx += stride * i
            stride *= len(xs)
        return idx

    def mult(self, x, coords):
        li = self.lattice_index(coords)
        U = np.zeros(int(np.prod(self.shape)))
        U[li] = x
        Y = self.apply_lattice(U.reshape(self.shape)).reshape(-1)
        return Y[li]

    def diag(self, coords):
        """diagonal of A on the caller's numbering"""
        dim = self.dim
        dM = [np.diag(m[0]) for m in self.mats]
        dK = [np.diag(m[1]) for m in self.mats]
        dC = [np.diag(m[2]) for m in self.mats]

        def outer(vs):                       # vs in x, y, z order -> lattice (z, y, x)
            out = vs[0]
            for v in vs[1:]:
                out = np.multiply.outer(v, out)
            return out
        D = np.zeros(self.shape)
        for d in range(dim):
            if self.kappa is not None:
                D += self.kappa * outer([dK[o] if o == d else dM[o] for o in range(dim)])
            if self.vel is not None:
                D += self.alpha * self.vel[d] * outer([dC[o] if o == d else dM[o] for o in range(dim)])
        if self.mass is not None:
            D += self.mass * outer(dM)
        return D.reshape(-1)[self.lattice_index(coords)]


def graded_axis(n, grade=0.35):
    """n+1 monotone vertex coordinates on [0,1], cell sizes varying smoothly by about +-grade"""
    t = np.linspace(0.0, 1.0, n + 1)
    return t + grade * np.sin(2 * np.pi * t) / (2 * np.pi)


def rectilinear_vertices(vx_unit, axes):
    """map the vertices of a unit Cartesian mesh with n_d cells per axis onto the graded axes"""
    out = np.zeros_like(vx_unit)
    for d, a in enumerate(axes):
        n = len(a) - 1
        i = np.rint(vx_unit[:, d] * n).astype(np.int64)
        out[:, d] = np.asarray(a)[i]
    return out
