"""GPU parity of the steps either side of the solve: DomainLFIntegrator,
ProjectBdrCoefficient, ComputeL2Error -- CUDA path through the C ABI vs the CPU oracle's dense
formulation on the same inputs.  Tolerance: 1e-12 relative (FP64, different summation order)."""
import numpy as np
import pytest

import cdm_b200 as cdm
from test_gpu_parity import Dev, make, rel, torch, ctx  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu
TOL = 1e-12

CASES = [(2, 1, 6), (2, 2, 5), (2, 3, 4), (2, 5, 3), (2, 6, 2),
         (3, 1, 4), (3, 2, 4), (3, 3, 3), (3, 4, 3), (3, 5, 2), (3, 6, 2)]


def forcing(x):
    return 1.0 + np.sin(2.3 * x[..., 0]) * np.cos(1.7 * x[..., 1]) + 0.5 * x[..., -1] ** 2


@pytest.mark.parametrize("dim,p,n", CASES)
def test_rule_coords_and_domain_lf(torch, ctx, orc, dim, p, n):
    P, mesh, sp = make(ctx, orc, dim, p, n, perturb=0.12, shuffle_seed=p)
    d = Dev(torch, ctx)
    for q1d in (0, p + 1, p + 2):
        qq = sp.q1d if q1d == 0 else q1d
        xq = sp.rule_coords(q1d)
        assert np.allclose(xq, P.rule_coords(qq), rtol=0, atol=1e-14)
    # default rule (order 2p -> p+1 points), forcing evaluated by the caller at the rule's points
    xq = sp.rule_coords(p + 1)
    f = forcing(xq)
    b = d.zeros(sp.ndof)
    sp.domain_lf(f, b)
    ref = P.domain_lf(f)
    assert rel(d.down(b), ref) <= TOL
    # accumulate with a scale (rhs += dt * (f, v), diffusion_mms.cpp:437), device-resident f, explicit rule
    xq2 = sp.rule_coords(p + 2)
    f2 = forcing(xq2)
    sp.domain_lf(d.up(f2.reshape(-1)), b, q1d=p + 2, scale=0.05, accumulate=True)
    ref2 = P.domain_lf(f2, q1d=p + 2, scale=0.05, b=ref.copy())
    assert rel(d.down(b), ref2) <= TOL
    # device-resident rule coordinates
    xdev = torch.zeros((sp.ne, (p + 1) ** dim, dim), dtype=torch.float64, device="cuda")
    sp.rule_coords(p + 1, out=xdev)
    ctx.sync()
    assert np.array_equal(xdev.cpu().numpy(), xq)


@pytest.mark.parametrize("dim,p,n", CASES)
def test_l2_error_and_norms(torch, ctx, orc, dim, p, n):
    P, mesh, sp = make(ctx, orc, dim, p, n, perturb=0.12, shuffle_seed=p + 1)
    d = Dev(torch, ctx)
    rng = np.random.default_rng(17 + p)
    u = rng.uniform(-1, 1, sp.ndof)
    xq = sp.rule_coords(p + 2)
    uex = forcing(xq)
    ud = d.up(u)
    e_ref = P.l2_error(u, uex)
    assert abs(sp.l2_error(ud, uex) - e_ref) <= TOL * e_ref
    assert abs(sp.l2_error(None, uex) - P.l2_error(None, uex)) <= TOL * e_ref            # ComputeGlobalLpNorm
    assert abs(sp.l2_error(ud, None) - P.l2_error(u, None)) <= TOL * e_ref
    # another rule + device-resident exact values; reproducible bit for bit
    xq3 = sp.rule_coords(p + 1)
    ex3 = d.up(forcing(xq3).reshape(-1))
    r1 = sp.l2_error(ud, ex3, q1d=p + 1)
    r2 = sp.l2_error(ud, ex3, q1d=p + 1)
    assert r1 == r2
    assert abs(r1 - P.l2_error(u, forcing(xq3), q1d=p + 1)) <= TOL * e_ref
    one = np.ones(xq.shape[:2])
    assert abs(sp.l2_error(None, one) - 1.0) <= 1e-13                                    # sqrt(volume)


def test_project_boundary_values(torch, ctx, orc):
    P, mesh, sp = make(ctx, orc, 3, 3, 3, perturb=0.1)
    d = Dev(torch, ctx)
    ess = sp.essential_dofs(np.ones(6, np.int32))
    X = sp.dof_coords()
    g = forcing(X)
    u = d.zeros(sp.ndof)
    sp.project_dofs(ess, g[ess], u)
    want = np.where(P.ess_mark, g, 0.0)
    assert np.array_equal(d.down(u), want)
    # device-resident index / value arrays
    u2 = d.zeros(sp.ndof)
    sp.project_dofs(torch.from_numpy(ess).cuda(), d.up(g[ess]), u2)
    assert np.array_equal(d.down(u2), want)


@pytest.mark.parametrize("dim", [2, 3])
def test_per_step_update_with_device_resident_coefficients(torch, ctx, orc, dim):
    """ALE-style per-step re-setup (diffusion_mms_ale.cpp:1017-1023): per-point coefficient arrays that
    already live on the device are used in place; host arrays reuse one staging buffer across steps."""
    P, mesh, sp = make(ctx, orc, dim, 2, 3, perturb=0.1)
    d = Dev(torch, ctx)
    op = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0)
    rng = np.random.default_rng(11)
    x = rng.uniform(-1, 1, P.ndof)
    xd, yd = d.up(x), d.zeros(P.ndof)
    for step in range(3):
        kap = rng.uniform(0.5, 1.5, (P.ne, P.nq))
        vel = rng.uniform(-1, 1, (P.ne, P.nq, dim))
        mass = rng.uniform(0.5, 1.5, (P.ne, P.nq))
        P.set_coefficients(kap, vel, -1.0, mass)
        ref = P.pa_apply(x)
        if step % 2 == 0:
            op.update(kappa=d.up(kap), vel=d.up(vel), alpha=-1.0, mass=d.up(mass))      # device tensors
        else:
            op.update(kappa=kap, vel=vel, alpha=-1.0, mass=mass)                        # host arrays
        Dd, Dc, Dm = op.qdata()
        assert rel(Dd, P.Dd) < 1e-13 and rel(Dc, P.Dc) < 1e-13 and rel(Dm, P.Dm) < 1e-13
        op.MultUnconstrained(xd, yd)
        assert rel(d.down(yd), ref) <= TOL


@pytest.mark.parametrize("dim,p,n", [(2, 2, 8), (3, 2, 4)])
def test_steady_mms_end_to_end_on_device(torch, ctx, orc, dim, p, n):
    """linear_convection_diffusion_2D.cpp:311-392 with every step on the device: LF rhs, boundary
    projection, FormLinearSystem (EliminateRHS), GMRES+Jacobi, ComputeL2Error -- against the oracle
    running the reference's assembled path."""
    kappa, c, s = 0.1, (1.0, -2.0, 0.5)[:dim], 1.0
    P, mesh, sp = make(ctx, orc, dim, p, n, perturb=0.0, kappa=kappa, vel=c, mass=s)
    d = Dev(torch, ctx)
    k = 3 * np.pi

    def exact(x):
        r = np.sin(k * x[..., 0]) * np.sin(k * x[..., 1])
        return r * (np.sin(k * x[..., 2]) if dim == 3 else 1.0)

    def rhs(x):
        sx, cx, sy, cy = np.sin(k * x[..., 0]), np.cos(k * x[..., 0]), np.sin(k * x[..., 1]), np.cos(k * x[..., 1])
        if dim == 2:
            return kappa * 2 * k * k * sx * sy + c[0] * k * cx * sy + c[1] * k * sx * cy + s * sx * sy
        sz, cz = np.sin(k * x[..., 2]), np.cos(k * x[..., 2])
        return (kappa * 3 * k * k * sx * sy * sz + c[0] * k * cx * sy * sz + c[1] * k * sx * cy * sz
                + c[2] * k * sx * sy * cz + s * sx * sy * sz)

    # --- CUDA path
    ess = sp.essential_dofs(np.ones(2 * dim, np.int32))
    op = cdm.ConvectionDiffusionOperator(sp, kappa=kappa, vel=c, mass=s, ess_dofs=ess)
    b = d.zeros(sp.ndof)
    sp.domain_lf(rhs(sp.rule_coords(p + 1)), b)
    u = d.zeros(sp.ndof)
    sp.project_dofs(ess, exact(sp.dof_coords())[ess], u)
    op.EliminateRHS(u, b)
    solver = cdm.GMRESSolver(rtol=1e-12, atol=1e-14, max_it=2000)
    solver.SetOperator(op)
    x = d.zeros(sp.ndof)
    solver.Mult(b, x)
    assert solver.GetConverged()
    uex_q = exact(sp.rule_coords(p + 2))
    err = sp.l2_error(x, uex_q)
    nrm = sp.l2_error(None, uex_q)
    # --- oracle: assembled matrix, eliminated, same rhs / norms
    bo = P.domain_lf(rhs(P.rule_coords(p + 1)))
    A = P.csr()
    x0 = np.where(P.ess_mark, exact(P.coords()), 0.0)
    A.eliminate(P.ess_mark, x0, bo)
    xo, info = A.op().gmres(bo, dinv=1.0 / A.diag(), rtol=1e-12, atol=1e-14, max_it=2000)
    assert info["converged"]
    erro = P.l2_error(xo, exact(P.rule_coords(p + 2)))
    assert rel(d.down(x), xo) <= 1e-9
    assert abs(err - erro) <= 1e-8 * erro
    assert err / nrm < 0.2                     # it is a discretisation error, not garbage
