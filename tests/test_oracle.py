"""CPU tests of the oracle itself: known-answer checks that pin the restatement
(the reference commits no golden vectors; MFEM/PETSc are not installable here --
"parity unpinned").  The two independent formulations (assembled
CSR = what the app executes, and sum-factorised PA) must agree to round-off."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_gauss_legendre_known_values(orc):
    x, w = orc.gauss_legendre(2)
    assert np.allclose(x, [0.5 - 0.5 / np.sqrt(3), 0.5 + 0.5 / np.sqrt(3)], atol=1e-15)
    assert np.allclose(w, [0.5, 0.5], atol=1e-15)
    x, w = orc.gauss_legendre(3)
    assert np.allclose(x, [0.5 - 0.5 * np.sqrt(0.6), 0.5, 0.5 + 0.5 * np.sqrt(0.6)], atol=1e-15)
    assert np.allclose(w, [5 / 18, 8 / 18, 5 / 18], atol=1e-15)
    for n in range(1, 9):
        x, w = orc.gauss_legendre(n)
        assert abs(w.sum() - 1.0) < 1e-14
        for k in range(2 * n):          # exact for degree <= 2n-1 on [0,1]
            assert abs((w * x ** k).sum() - 1.0 / (k + 1)) < 1e-14


def test_gauss_lobatto_known_values(orc):
    assert np.allclose(orc.gauss_lobatto(2), [0, 1])
    assert np.allclose(orc.gauss_lobatto(3), [0, 0.5, 1], atol=1e-15)
    s5 = 1 / np.sqrt(5)
    assert np.allclose(orc.gauss_lobatto(4), [0, 0.5 - 0.5 * s5, 0.5 + 0.5 * s5, 1], atol=1e-15)
    s37 = np.sqrt(3 / 7)
    assert np.allclose(orc.gauss_lobatto(5), [0, 0.5 - 0.5 * s37, 0.5, 0.5 + 0.5 * s37, 1], atol=1e-15)


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 2), (2, 3), (3, 1), (3, 2), (3, 3), (3, 6)])
def test_basis_partition_of_unity_and_exactness(orc, dim, p):
    q = orc.q1d(dim, p)
    assert q == (p + 2 if dim == 3 else p + 1)
    B, G, w = orc.basis(p, q)
    assert np.allclose(B.sum(1), 1.0, atol=1e-13)
    assert np.allclose(G.sum(1), 0.0, atol=1e-11)
    xn, (xq, _) = orc.gauss_lobatto(p + 1), orc.gauss_legendre(q)
    for k in range(p + 1):               # interpolation / differentiation of x^k is exact
        assert np.allclose(B @ xn ** k, xq ** k, atol=1e-13)
        assert np.allclose(G @ xn ** k, k * xq ** max(k - 1, 0) if k else 0 * xq, atol=1e-11)


def test_cartesian_mesh_conventions(orc):
    vx, ev, bv, battr = orc.cart_mesh(3, [2, 3, 4])
    assert vx.shape == (3 * 4 * 5, 3) and ev.shape == (24, 8)
    assert list(ev[0]) == [0, 1, 4, 3, 12, 13, 16, 15]      # (x,y,z),(x+1,y,z),(x+1,y+1,z),(x,y+1,z), then z+1
    assert np.allclose(vx[1], [0.5, 0, 0]) and np.allclose(vx[3], [0, 1 / 3, 0])
    # boundary attributes: bottom1 front2 right3 back4 left5 top6
    for a, (axis, val) in {1: (2, 0.0), 6: (2, 1.0), 2: (1, 0.0), 4: (1, 1.0), 5: (0, 0.0), 3: (0, 1.0)}.items():
        faces = bv[battr == a]
        assert len(faces) > 0 and np.allclose(vx[faces.reshape(-1)][:, axis], val)
    vx, ev, bv, battr = orc.cart_mesh(2, [3, 2])
    assert list(ev[0]) == [0, 1, 5, 4]
    for a, (axis, val) in {1: (1, 0.0), 2: (0, 1.0), 3: (1, 1.0), 4: (0, 0.0)}.items():
        assert np.allclose(vx[bv[battr == a].reshape(-1)][:, axis], val)


@pytest.mark.parametrize("dim,p,seed", [(2, 2, None), (2, 3, 5), (3, 1, None), (3, 2, 7), (3, 3, None), (3, 3, 11), (3, 4, 2)])
def test_h1_numbering_is_conforming(orc, dim, p, seed):
    P = orc.Problem(dim, p, 3, perturb=0.15, shuffle_seed=seed)
    n1 = 3 * p + 1
    assert P.ndof == n1 ** dim                       # continuous Q_p space on a 3^dim mesh
    xc = orc.node_coords(dim, p, P.ev, P.vx).reshape(-1, dim)
    g = P.elem_dof.reshape(-1)
    ref = np.zeros((P.ndof, dim))
    ref[g] = xc
    assert np.abs(ref[g] - xc).max() < 1e-14         # every sharer sees the dof at the same point
    assert len(np.unique(np.round(ref, 12), axis=0)) == P.ndof
    # entity-major layout: vertices first
    assert set(P.elem_dof[:, [0, p]].reshape(-1)) <= set(range(P.nv))
    # restriction arrays
    assert P.offsets[0] == 0 and P.offsets[-1] == P.ne * P.nd
    assert np.array_equal(np.sort(P.indices), np.arange(P.ne * P.nd))
    assert np.all(g[P.indices] == np.repeat(np.arange(P.ndof), np.diff(P.offsets)))
    for d in range(P.ndof):
        seg = P.indices[P.offsets[d]:P.offsets[d + 1]]
        assert np.all(np.diff(seg) > 0)
    # essential dofs = dofs on the boundary of the unit box
    on_bdr = np.any((np.abs(ref) < 1e-12) | (np.abs(ref - 1) < 1e-12), axis=1)
    assert np.array_equal(on_bdr, P.ess_mark.astype(bool))


def test_essential_dofs_subset_of_attributes(orc):
    P = orc.Problem(2, 3, 4, ess_attrs=[2, 4])        # x-faces only (transient driver, _1D.cpp:214-258)
    X = P.coords()
    expect = (np.abs(X[:, 0]) < 1e-12) | (np.abs(X[:, 0] - 1) < 1e-12)
    assert np.array_equal(expect, P.ess_mark.astype(bool))


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 2), (2, 3), (2, 5), (3, 1), (3, 2), (3, 3), (3, 4)])
def test_csr_and_pa_formulations_agree(orc, dim, p):
    P = orc.Problem(dim, p, 3 if dim == 3 else 5, perturb=0.15, shuffle_seed=p)
    A = P.csr()
    rng = np.random.default_rng(12345)
    x = rng.uniform(-1, 1, P.ndof)
    y1, y2 = A.spmv(x), P.pa_apply(x)
    assert np.linalg.norm(y1 - y2) <= 1e-13 * np.linalg.norm(y1)
    d1, d2 = A.diag(), P.pa_diag()
    assert np.linalg.norm(d1 - d2) <= 1e-13 * np.linalg.norm(d1)


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_fast_cpu_apply_matches_three_pass_apply(orc, p):
    """the fused order-specialised CPU apply timed as cpu_baseline is the same operator"""
    P = orc.Problem(3, p, 3 if p < 5 else 2, perturb=0.12, shuffle_seed=p)
    x = np.random.default_rng(p).uniform(-1, 1, P.ndof)
    a, b = P.pa_apply(x), P.pa_apply_fast(x)
    assert np.linalg.norm(a - b) <= 1e-14 * np.linalg.norm(a)
    for kw in (dict(kappa=None, vel=None, mass=1.0), dict(kappa=0.3, vel=None, mass=None)):
        Q = orc.Problem(3, p, 2, perturb=0.1, **kw)
        x = np.random.default_rng(1).uniform(-1, 1, Q.ndof)
        assert np.linalg.norm(Q.pa_apply(x) - Q.pa_apply_fast(x)) <= 1e-14 * np.linalg.norm(Q.pa_apply(x))


def test_variable_and_matrix_coefficients_agree(orc):
    rng = np.random.default_rng(3)
    for dim in (2, 3):
        P = orc.Problem(dim, 2, 3, perturb=0.1)
        nsym = dim * (dim + 1) // 2
        # SPD matrix coefficient per point, vector velocity per point, scalar mass per point
        L = rng.uniform(-0.3, 0.3, (P.ne, P.nq, dim, dim)) + np.eye(dim)
        M = L @ np.swapaxes(L, -1, -2)
        pairs = [(0, 0), (1, 0), (1, 1)] if dim == 2 else [(0, 0), (1, 0), (2, 0), (1, 1), (2, 1), (2, 2)]
        kap = np.stack([M[..., r, c] for r, c in pairs], axis=-1)
        assert kap.shape[-1] == nsym
        vel = rng.uniform(-1, 1, (P.ne, P.nq, dim))
        mass = rng.uniform(0.5, 1.5, (P.ne, P.nq))
        P.set_coefficients(kap, vel, -1.0, mass)
        x = rng.uniform(-1, 1, P.ndof)
        y1, y2 = P.csr().spmv(x), P.pa_apply(x)
        assert np.linalg.norm(y1 - y2) <= 1e-13 * np.linalg.norm(y1)


def test_operator_known_answers(orc):
    P = orc.Problem(3, 2, 3, perturb=0.15, kappa=0.7, vel=None, mass=None)
    one = np.ones(P.ndof)
    K = P.csr().to_scipy()
    assert np.abs(K @ one).max() < 1e-12                      # constants in the kernel of grad
    assert abs(K - K.T).max() < 1e-13                         # K symmetric
    P = orc.Problem(3, 2, 3, perturb=0.15, kappa=None, vel=None, mass=1.0)
    M = P.csr().to_scipy()
    assert abs(M - M.T).max() < 1e-14
    assert abs(one @ (M @ one) - 1.0) < 1e-13                 # volume of the unit cube
    P = orc.Problem(2, 3, 4, perturb=0.15, kappa=None, vel=(1.0, -2.0), mass=None)
    Cm = P.csr().to_scipy()
    assert np.abs(Cm @ np.ones(P.ndof)).max() < 1e-12         # C 1 = 0
    free = ~P.ess_mark.astype(bool)                            # C + C^T = 0 on interior dofs (constant c)
    S = (Cm + Cm.T)[free][:, free]
    assert abs(S).max() < 1e-13


def test_exact_on_polynomials(orc):
    """a(u, v) for u, v in Q_p on an affine mesh is integrated exactly: compare with calculus."""
    P = orc.Problem(2, 2, 2, perturb=0.0, kappa=0.5, vel=(1.0, -2.0), mass=3.0)
    X = P.coords()
    u = X[:, 0] ** 2 * X[:, 1]            # grad u = (2xy, x^2)
    v = X[:, 0] * X[:, 1] ** 2            # grad v = (y^2, 2xy)
    # int_0^1 int_0^1 [ 0.5 (2xy*y^2 + x^2*2xy) + (2xy - 2x^2) x y^2 + 3 x^3 y^3 ]
    exact = 0.5 * (2 / 2 / 4 + 2 / 4 / 2) + (2 / 3 / 4 - 2 / 4 / 3) + 3 / 16
    assert abs(v @ P.pa_apply(u) - exact) < 1e-13


def mms_2d(P, n=3, m=3, kappa=0.1, c=(1.0, -2.0), s=1.0):
    """app's manufactured solution and forcing (linear_convection_diffusion_2D.cpp:159-215):
    nodal values of u_exact and f at the quadrature points of every element"""
    from oracle import pyoracle as O
    X = P.coords()
    uex = np.sin(n * np.pi * X[:, 0]) * np.sin(m * np.pi * X[:, 1])
    V = P.vx[P.ev]                                     # (ne, 4, 2)
    xg, _ = O.gauss_legendre(P.q1d)
    xq = np.zeros((P.ne, P.nq, 2))
    for q in range(P.nq):
        a, b = xg[q % P.q1d], xg[q // P.q1d]
        N = np.array([(1 - a) * (1 - b), a * (1 - b), a * b, (1 - a) * b])
        xq[:, q, :] = np.einsum("k,ekc->ec", N, V)
    sx, cx = np.sin(n * np.pi * xq[..., 0]), np.cos(n * np.pi * xq[..., 0])
    sy, cy = np.sin(m * np.pi * xq[..., 1]), np.cos(m * np.pi * xq[..., 1])
    f = (kappa * (n * n + m * m) * np.pi ** 2 * sx * sy + c[0] * n * np.pi * cx * sy
         + c[1] * m * np.pi * sx * cy + s * sx * sy)
    return uex, f


@pytest.mark.parametrize("p", [1, 2, 3])
def test_mms_convergence_rate_2d(orc, p):
    """-kappa lap u + c.grad u + s u = f, Dirichlet: L2-ish nodal error decays like h^(p+1)."""
    errs = []
    for n in (4, 8):
        P = orc.Problem(2, p, n, perturb=0.0, kappa=0.1, vel=(1.0, -2.0), mass=1.0)
        uex, f = mms_2d(P)
        # DomainLFIntegrator(f): b_i = sum_q w detJ f(x_q) phi_i(x_q) = (mass form with coefficient f) applied to 1
        Pf = orc.Problem(2, p, n, perturb=0.0, kappa=None, vel=None, mass=f)
        b = Pf.pa_apply(np.ones(P.ndof))
        A = P.csr()
        x0 = np.where(P.ess_mark, uex, 0.0)
        A.eliminate(P.ess_mark, x0, b)
        x, info = A.op().gmres(b, dinv=1.0 / A.diag(), rtol=1e-13, atol=1e-14, max_it=2000, restart=200)
        assert info["converged"]
        errs.append(np.sqrt(np.mean((x - uex) ** 2)))
    rate = np.log2(errs[0] / errs[1])
    assert rate > p + 0.6, (errs, rate)


def test_constrained_operator_matches_eliminated_matrix(orc):
    """Left-Jacobi GMRES iterates coincide for FormLinearSystem on the assembled matrix
    (what the app runs) and ConstrainedOperator + EliminateRHS on the PA operator."""
    P = orc.Problem(3, 2, 3, perturb=0.1)
    rng = np.random.default_rng(5)
    b0 = rng.uniform(-1, 1, P.ndof)
    g = np.where(P.ess_mark, rng.uniform(-1, 1, P.ndof), 0.0)
    A = P.csr()
    b1 = b0.copy()
    A.eliminate(P.ess_mark, g, b1)
    d1 = A.diag()
    x1, i1 = A.op().gmres(b1, dinv=1 / d1)
    op = P.pa_op(constrained=True)
    b2 = b0.copy()
    op.eliminate_rhs(g, b2)
    d2 = np.where(P.ess_mark, 1.0, P.pa_diag())
    x2, i2 = op.gmres(b2, dinv=1 / d2)
    assert i1["converged"] and i2["converged"] and i1["iters"] == i2["iters"]
    assert np.allclose(i1["hist"], i2["hist"], rtol=1e-9, atol=1e-14 * i1["hist"][0])
    assert np.linalg.norm(x1 - x2) <= 1e-10 * np.linalg.norm(x1)
    assert np.allclose(x1[P.ess_mark.astype(bool)], g[P.ess_mark.astype(bool)], atol=1e-12)


def test_gmres_variants_and_cg_solve(orc):
    import scipy.sparse.linalg as spla
    P = orc.Problem(2, 2, 6, perturb=0.1, vel=None)           # SPD: diffusion + mass
    A = P.csr()
    rng = np.random.default_rng(9)
    xs = rng.uniform(-1, 1, P.ndof)
    b = A.spmv(xs)
    for variant in (0, 1):
        x, info = A.op().gmres(b, dinv=1 / A.diag(), variant=variant, rtol=1e-12)
        assert info["converged"] and np.linalg.norm(x - xs) < 1e-8 * np.linalg.norm(xs)
        assert np.all(np.diff(info["hist"]) <= 1e-12)         # GMRES residuals are monotone
    x, info = A.op().cg(b, rtol=1e-12)
    assert info["converged"] and np.linalg.norm(x - xs) < 1e-8 * np.linalg.norm(xs)
    xd = spla.spsolve(A.to_scipy().tocsc(), b)
    assert np.linalg.norm(xd - xs) < 1e-9 * np.linalg.norm(xs)


def test_golden_fixture_matches(orc):
    """tests/golden/oracle_small.json was produced by tests/golden/make_golden.py from
    this oracle: guards the oracle (and, on the GPU box, the CUDA path) against drift."""
    with open(os.path.join(GOLD, "oracle_small.json")) as fh:
        gold = json.load(fh)
    for case in gold["cases"]:
        P = orc.Problem(case["dim"], case["p"], case["n"], perturb=case["perturb"])
        x = np.sin(1.0 + 0.37 * np.arange(P.ndof))
        y = P.pa_apply(x)
        assert P.ndof == case["ndof"]
        assert np.array_equal(P.elem_dof[:2].reshape(-1), np.array(case["elem_dof_head"], np.int32))
        assert abs(np.linalg.norm(y) - case["y_norm"]) <= 1e-13 * case["y_norm"]
        assert np.allclose(y[:8], case["y_head"], rtol=1e-12, atol=1e-14)
        assert abs(P.pa_diag().sum() - case["diag_sum"]) <= 1e-12 * abs(case["diag_sum"])
        fn = lambda c: 1.0 + np.sin(2.3 * c[..., 0]) * np.cos(1.7 * c[..., 1]) + 0.5 * c[..., -1] ** 2
        lf = P.domain_lf(fn(P.rule_coords(P.p + 1)))
        assert abs(np.linalg.norm(lf) - case["lf_norm"]) <= 1e-13 * case["lf_norm"]
        assert np.allclose(lf[:8], case["lf_head"], rtol=1e-12, atol=1e-15)
        assert abs(P.l2_error(x, fn(P.rule_coords(P.p + 2))) - case["l2_error"]) <= 1e-13 * case["l2_error"]


@pytest.mark.parametrize("dim", [2, 3])
def test_identity_ale_map_reduces_to_the_static_mesh_operator(orc, dim):
    """The reference's own verification step for the ALE driver (diffusion_mms_ale_plan.tex:486-491): with the
    identity map (J = 1, CofA = I, phi_hat = 0, div phi_hat = 0) the per-step form
    Mass(J) + Diffusion((alpha dt / J) CofA CofA^T) + Convection(phi_hat, -1) + Mass(-div phi_hat)
    (diffusion_mms_ale.cpp:1017-1023) is exactly the static-mesh backward-Euler form M + alpha dt K of
    diffusion_mms.cpp:301-305.  Per-point coefficient arrays vs constants, both oracle formulations."""
    alpha, dt, p, n = 0.1, 0.05, 2, 3
    static = orc.Problem(dim, p, n, perturb=0.1, kappa=alpha * dt, vel=None, mass=1.0)
    nsym = dim * (dim + 1) // 2
    eye = np.zeros(nsym)
    eye[[0, 2] if dim == 2 else [0, 3, 5]] = 1.0                       # packed symmetric identity
    metric = np.broadcast_to(alpha * dt * eye, (static.ne, static.nq, nsym)).copy()
    phi_hat = np.zeros((static.ne, static.nq, dim))
    jac = np.ones((static.ne, static.nq))                              # J - div(phi_hat) = 1 - 0 in one mass term
    ale = orc.Problem(dim, p, n, perturb=0.1, kappa=metric, vel=phi_hat, alpha=-1.0, mass=jac)
    x = np.sin(1.0 + 0.37 * np.arange(static.ndof))
    ys = static.pa_apply(x)
    assert np.linalg.norm(ale.pa_apply(x) - ys) <= 1e-14 * np.linalg.norm(ys)
    assert np.linalg.norm(ale.csr().spmv(x) - static.csr().spmv(x)) <= 1e-14 * np.linalg.norm(ys)
