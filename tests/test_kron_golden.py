"""The oracle against a closed-form golden that does not come from it (tests/kron_golden.py):
Kronecker sums of 1-D Gauss-Lobatto Lagrange matrices on graded rectilinear meshes, numpy only.
Pins the oracle's basis, rule, D-tensor algebra, element/dof numbering and both formulations
(PA and assembled CSR) to round-off on affine elements, for every order and both dimensions."""
import numpy as np
import pytest

import kron_golden as kg


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def graded_problem(orc, dim, p, n, shuffle_seed=None, **kw):
    P = orc.Problem(dim, p, n, perturb=0.0, shuffle_seed=shuffle_seed, **kw)
    axes = [kg.graded_axis(k, 0.3 + 0.05 * d) for d, k in enumerate(n)]
    P.vx = kg.rectilinear_vertices(P.vx, axes)
    P.set_coefficients(P.kappa, P.vel, P.alpha, P.mass)
    return P, axes


def test_p1_unit_cube_element_matrices_closed_form():
    """the textbook trilinear matrices: mass = (1/6 [[2,1],[1,2]])^{(x)3}, stiffness from [[1,-1],[-1,1]]"""
    _, m, k, c = kg.ref_matrices(1)
    assert np.allclose(m, np.array([[2, 1], [1, 2]]) / 6.0, atol=1e-15)
    assert np.allclose(k, np.array([[1, -1], [-1, 1]]), atol=1e-15)
    assert np.allclose(c, np.array([[-0.5, 0.5], [-0.5, 0.5]]), atol=1e-15)
    M = np.kron(m, np.kron(m, m))
    assert abs(M.sum() - 1.0) < 1e-15 and abs(M[0, 0] - 1.0 / 27.0) < 1e-16 and abs(M[0, 7] - 1.0 / 216.0) < 1e-16
    K = np.kron(m, np.kron(m, k)) + np.kron(m, np.kron(k, m)) + np.kron(k, np.kron(m, m))
    assert abs(K[0, 0] - 1.0 / 3.0) < 1e-15 and abs(K[0, 7] + 1.0 / 12.0) < 1e-15 and np.allclose(K.sum(1), 0, atol=1e-15)


def test_gauss_tables_match_numpy(orc):
    from numpy.polynomial.legendre import leggauss
    for n in range(1, 9):
        x, w = orc.gauss_legendre(n)
        xr, wr = leggauss(n)
        assert np.allclose(x, 0.5 * (xr + 1), atol=2e-16 * 8) and np.allclose(w, 0.5 * wr, atol=1e-15)
    for p in range(1, 7):
        assert np.allclose(orc.gauss_lobatto(p + 1), kg.gll_nodes(p), atol=1e-15)


def test_p1_single_element_oracle_matrix(orc):
    """one unit cube, order 1: the oracle's assembled matrix is the closed-form one entry by entry"""
    P = orc.Problem(3, 1, 1, perturb=0.0, kappa=0.7, vel=(1.0, -2.0, 0.5), mass=1.3)
    A = P.csr().to_scipy().toarray()
    _, m, k, c = kg.ref_matrices(1)
    I = [m, m, m]
    K = np.kron(m, np.kron(m, k)) + np.kron(m, np.kron(k, m)) + np.kron(k, np.kron(m, m))
    Cx, Cy, Cz = np.kron(m, np.kron(m, c)), np.kron(m, np.kron(c, m)), np.kron(c, np.kron(m, m))
    ref_lex = 0.7 * K + (1.0 * Cx - 2.0 * Cy + 0.5 * Cz) + 1.3 * np.kron(m, np.kron(m, m))
    g = P.elem_dof[0]                              # lexicographic node -> global dof
    assert np.allclose(A[np.ix_(g, g)], ref_lex, atol=1e-15)


@pytest.mark.parametrize("dim,p,n", [(2, 1, [7, 5]), (2, 2, [5, 4]), (2, 3, [4, 5]), (2, 4, [3, 4]), (2, 5, [3, 4]), (2, 6, [2, 3]),
                                     (3, 1, [5, 4, 6]), (3, 2, [4, 3, 5]), (3, 3, [4, 5, 3]), (3, 4, [3, 3, 4]),
                                     (3, 5, [2, 3, 3]), (3, 6, [2, 3, 2])])
def test_oracle_matches_kronecker_golden(orc, dim, p, n):
    vel = (1.0, -2.0, 0.5)[:dim]
    P, axes = graded_problem(orc, dim, p, n, shuffle_seed=3, kappa=0.1, vel=vel, mass=1.0)
    K = kg.KronOperator(p, axes, kappa=0.1, vel=vel, alpha=1.0, mass=1.0)
    x = np.random.default_rng(1).uniform(-1, 1, P.ndof)
    X = P.coords()
    y = K.mult(x, X)
    assert rel(P.pa_apply(x), y) < 1e-13
    assert rel(P.pa_apply_fast(x), y) < 1e-13
    assert rel(P.csr().spmv(x), y) < 1e-13
    assert rel(P.pa_diag(), K.diag(X)) < 1e-13


def test_oracle_matches_kronecker_golden_midsize(orc):
    """3D order 3 on 20 x 19 x 21 graded cells (7980 elements, 211 k dofs): element-order / dof-numbering
    mistakes that only show on a mesh with many distinct element sizes"""
    n = [20, 19, 21]
    P, axes = graded_problem(orc, 3, 3, n, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0)
    K = kg.KronOperator(3, axes, kappa=0.1, vel=(1.0, -2.0, 0.5), alpha=1.0, mass=1.0)
    x = np.random.default_rng(2).uniform(-1, 1, P.ndof)
    X = P.coords()
    assert rel(P.pa_apply_fast(x), K.mult(x, X)) < 1e-13
    assert rel(P.pa_diag(), K.diag(X)) < 1e-13
