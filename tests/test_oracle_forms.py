"""Oracle known-answer tests for the steps either side of the solve: DomainLFIntegrator,
ComputeL2Error / ComputeGlobalLpNorm (linear_convection_diffusion_2D.cpp:341-343, :383-392)."""
import numpy as np
import pytest


@pytest.mark.parametrize("dim,p,n", [(2, 1, 5), (2, 3, 4), (3, 1, 3), (3, 2, 3), (3, 3, 2)])
def test_rule_coords_match_bilinear_map(orc, dim, p, n):
    P = orc.Problem(dim, p, n, perturb=0.1)
    q1d = p + 1
    xq = P.rule_coords(q1d)
    xg, _ = orc.gauss_legendre(q1d)
    V = P.vx[P.ev]                                        # (ne, nvpe, dim)
    for q in (0, q1d ** dim // 2, q1d ** dim - 1):
        a, b = xg[q % q1d], xg[(q // q1d) % q1d]
        if dim == 2:
            N = np.array([(1 - a) * (1 - b), a * (1 - b), a * b, (1 - a) * b])
        else:
            c = xg[q // (q1d * q1d)]
            N = np.array([(1 - a) * (1 - b) * (1 - c), a * (1 - b) * (1 - c), a * b * (1 - c), (1 - a) * b * (1 - c),
                          (1 - a) * (1 - b) * c, a * (1 - b) * c, a * b * c, (1 - a) * b * c])
        assert np.allclose(xq[:, q, :], np.einsum("k,ekc->ec", N, V), atol=1e-14)


@pytest.mark.parametrize("dim,p,n", [(2, 2, 5), (3, 1, 4), (3, 3, 3)])
def test_domain_lf_equals_mass_form_times_one(orc, dim, p, n):
    """b_i = int f phi_i = sum_j int f phi_i phi_j: the LF with MFEM's default rule (p+1 points)
    equals MassIntegrator(f) applied to 1 whenever both rules are exact for the integrand."""
    P = orc.Problem(dim, p, n, perturb=0.1, kappa=None, vel=None, mass=1.0)
    b = P.domain_lf(np.ones((P.ne, (p + 1) ** dim)))
    assert abs(b.sum() - 1.0) < 1e-13                    # volume of the unit square / cube
    ref = P.pa_apply(np.ones(P.ndof))
    assert np.linalg.norm(b - ref) <= 1e-13 * np.linalg.norm(ref)
    # a Q_1 forcing keeps the integrand (degree (dim-1) + 1 + p per direction) inside the exactness
    # range 2p+1 of the LF rule when p >= dim-1
    if p < dim - 1:
        return
    xq = P.rule_coords(p + 1)
    f = 1.0 + xq[..., 0] - 0.5 * xq[..., -1]
    xm = P.rule_coords(P.q1d)
    Pf = orc.Problem(dim, p, n, perturb=0.1, kappa=None, vel=None, mass=1.0 + xm[..., 0] - 0.5 * xm[..., -1])      # (ne, nq) array = per-point coefficient
    ref = Pf.pa_apply(np.ones(P.ndof))
    b = P.domain_lf(f, scale=2.0)
    assert np.linalg.norm(b - 2.0 * ref) <= 1e-12 * np.linalg.norm(ref)


@pytest.mark.parametrize("dim,p", [(2, 2), (3, 2)])
def test_l2_norms_known_answers(orc, dim, p):
    P = orc.Problem(dim, p, 3, perturb=0.1)
    q1d = max(2, 2 * p + 3) // 2 + 1
    assert q1d == p + 2
    xq = P.rule_coords(q1d)
    one = np.ones(xq.shape[:2])
    assert abs(P.l2_error(None, one) - 1.0) < 1e-13                 # ||1|| = sqrt(volume)
    assert abs(P.l2_error(np.ones(P.ndof), None) - 1.0) < 1e-13
    assert P.l2_error(np.ones(P.ndof), one) < 1e-13
    # u_h interpolates Q_p polynomials exactly on an affine mesh
    A = orc.Problem(dim, p, 3, perturb=0.0)
    X, xa = A.coords(), A.rule_coords(q1d)
    poly = lambda x: (1 + x[..., 0] ** p) * (2 - x[..., 1] ** p) * (1.0 if dim == 2 else (1 + 0.5 * x[..., -1]))
    assert A.l2_error(poly(X), poly(xa)) < 1e-13
    # ||x|| over the unit domain = 1/sqrt(3)
    assert abs(A.l2_error(None, xa[..., 0]) - 1 / np.sqrt(3)) < 1e-13


@pytest.mark.parametrize("p", [1, 2])
def test_mms_l2_error_rate_with_lf_rhs(orc, p):
    """The app's steady MMS end to end: DomainLFIntegrator rhs, nodal boundary projection,
    eliminated solve, ComputeL2Error -- L2 error decays like h^(p+1)."""
    kappa, c, s, nn, mm = 0.1, (1.0, -2.0), 1.0, 3, 3
    errs = []
    for n in (4, 8):
        P = orc.Problem(2, p, n, perturb=0.0, kappa=kappa, vel=c, mass=s)
        ex = lambda x: np.sin(nn * np.pi * x[..., 0]) * np.sin(mm * np.pi * x[..., 1])
        xq = P.rule_coords(p + 1)
        sx, cx = np.sin(nn * np.pi * xq[..., 0]), np.cos(nn * np.pi * xq[..., 0])
        sy, cy = np.sin(mm * np.pi * xq[..., 1]), np.cos(mm * np.pi * xq[..., 1])
        f = (kappa * (nn * nn + mm * mm) * np.pi ** 2 * sx * sy + c[0] * nn * np.pi * cx * sy
             + c[1] * mm * np.pi * sx * cy + s * sx * sy)
        b = P.domain_lf(f)
        A = P.csr()
        x0 = np.where(P.ess_mark, ex(P.coords()), 0.0)
        A.eliminate(P.ess_mark, x0, b)
        x, info = A.op().gmres(b, dinv=1.0 / A.diag(), rtol=1e-13, atol=1e-14, max_it=2000, restart=200)
        assert info["converged"]
        errs.append(P.l2_error(x, ex(P.rule_coords(p + 2))))
    rate = np.log2(errs[0] / errs[1])
    assert rate > p + 0.6, (errs, rate)
