"""Edge cases of the hot path on the GPU: degenerate meshes, the app's "all true dofs are essential"
branch (linear_convection_diffusion_2D.cpp:353-361), zero right-hand sides, hitting max_it
(reported through GetConverged, never an error: :371-374), unaligned / odd-length vectors."""
import numpy as np
import pytest

import cdm_b200 as cdm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch, cdm.Context(0)


def up(torch, a):
    t = torch.from_numpy(np.ascontiguousarray(a, np.float64)).cuda()
    torch.cuda.synchronize()
    return t


@pytest.mark.parametrize("dim,p", [(2, 1), (2, 3), (2, 4), (3, 1), (3, 2), (3, 3), (3, 4), (3, 5), (3, 6)])
def test_single_element_mesh(env, orc, dim, p):
    torch, ctx = env
    P = orc.Problem(dim, p, 1, perturb=0.0)
    mesh = cdm.Mesh.from_arrays(ctx, P.vx, P.ev, P.bv, P.battr)
    sp = cdm.H1Space(mesh, p)
    assert sp.ndof == (p + 1) ** dim and sp.ne == 1
    op = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0)
    x = np.random.default_rng(0).uniform(-1, 1, P.ndof)
    xd, yd = up(torch, x), up(torch, np.zeros(P.ndof))
    op.MultUnconstrained(xd, yd)
    ctx.sync()
    y_ref = P.pa_apply(x)
    assert np.linalg.norm(yd.cpu().numpy() - y_ref) <= 1e-12 * np.linalg.norm(y_ref)


def test_all_dofs_essential_is_the_identity(env, orc):
    """order 1 on one element, or any mesh whose dofs all lie on the boundary: the constrained operator is
    the identity and the solve returns the boundary data (the app skips the solve in this case)"""
    torch, ctx = env
    P = orc.Problem(3, 1, 1)
    assert P.ess.size == P.ndof
    mesh = cdm.Mesh.from_arrays(ctx, P.vx, P.ev, P.bv, P.battr)
    sp = cdm.H1Space(mesh, 1)
    op = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0, ess_dofs=P.ess)
    x = np.arange(1.0, P.ndof + 1)
    xd, yd = up(torch, x), up(torch, np.zeros(P.ndof))
    op.Mult(xd, yd)
    ctx.sync()
    assert np.array_equal(yd.cpu().numpy(), x)
    b = up(torch, np.zeros(P.ndof))
    op.EliminateRHS(xd, b)
    s = cdm.GMRESSolver()
    s.SetOperator(op)
    sol = up(torch, np.zeros(P.ndof))
    s.Mult(b, sol)
    ctx.sync()
    assert s.GetConverged() and s.GetNumIterations() <= 1
    assert np.allclose(sol.cpu().numpy(), x, rtol=0, atol=1e-14)


def test_zero_rhs_and_max_iterations(env, orc):
    torch, ctx = env
    P = orc.Problem(3, 2, 3, perturb=0.1, kappa=0.001)
    mesh = cdm.Mesh.from_arrays(ctx, P.vx, P.ev, P.bv, P.battr)
    sp = cdm.H1Space(mesh, 2)
    op = cdm.ConvectionDiffusionOperator(sp, kappa=0.001, vel=(1.0, -2.0, 0.5), mass=1.0, ess_dofs=P.ess)
    zero = up(torch, np.zeros(P.ndof))
    sol = up(torch, np.ones(P.ndof))
    for solver in (cdm.GMRESSolver(), cdm.CGSolver()):
        solver.SetOperator(op)
        solver.Mult(zero, sol)                       # b = 0, x0 = 0: converged at iteration 0, x = 0
        ctx.sync()
        assert solver.GetConverged() and solver.GetNumIterations() == 0 and solver.GetFinalNorm() == 0.0
        assert not sol.cpu().numpy().any()
    # max_it reached: reported, not raised; history has max_it + 1 entries
    b = up(torch, np.where(P.ess_mark, 0.0, np.random.default_rng(1).uniform(-1, 1, P.ndof)))
    s = cdm.GMRESSolver(cdm.GMRES_PETSC, 0, 3, 1e-14, 0.0, jacobi=True)
    s.SetOperator(op)
    s.Mult(b, sol)
    ctx.sync()
    assert not s.GetConverged() and s.GetNumIterations() == 3 and len(s.history) == 4
    assert s.history[-1] < s.history[0]
    # a restart length of 1 still works (GMRES(1))
    s = cdm.GMRESSolver(cdm.GMRES_PETSC, 1, 5, 1e-14, 0.0, jacobi=True)
    s.SetOperator(op)
    s.Mult(b, sol)
    assert s.GetNumIterations() == 5 and np.all(np.diff(s.history) <= 1e-15)


def test_iterative_mode_uses_the_initial_guess(env, orc):
    torch, ctx = env
    P = orc.Problem(2, 2, 5, perturb=0.1, vel=None)
    mesh = cdm.Mesh.from_arrays(ctx, P.vx, P.ev, P.bv, P.battr)
    sp = cdm.H1Space(mesh, 2)
    op = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, mass=1.0, ess_dofs=P.ess)
    rng = np.random.default_rng(2)
    xs = np.where(P.ess_mark, 0.0, rng.uniform(-1, 1, P.ndof))
    b = up(torch, P.pa_op(True).mult(xs))
    s = cdm.GMRESSolver()
    s.SetOperator(op)
    s.iterative_mode = True
    x0 = up(torch, xs)                               # exact solution as the initial guess: zero iterations
    s.SetAbsTol(1e-10)
    s.Mult(b, x0)
    ctx.sync()
    assert s.GetConverged() and s.GetNumIterations() == 0
    ref_x, info = P.pa_op(True).gmres(P.pa_op(True).mult(xs), dinv=1 / np.where(P.ess_mark, 1.0, P.pa_diag()), x0=0.5 * xs)
    x1 = up(torch, 0.5 * xs)
    s.SetAbsTol(1e-12)
    s.Mult(b, x1)
    ctx.sync()
    assert s.GetNumIterations() == info["iters"]
    assert np.max(np.abs(s.history - info["hist"]) / info["hist"][0]) < 1e-10


def test_unaligned_and_odd_vectors(env):
    torch, ctx = env
    rng = np.random.default_rng(3)
    n, k = 10007, 5
    big = up(torch, rng.uniform(-1, 1, (k + 1) * (n + 1) + 3))
    host = big.cpu().numpy()
    w = big[1:1 + n]                                 # 8-byte but not 16-byte aligned
    V = big[n + 2:n + 2 + k * (n + 1)].view(k, n + 1)[:, :n]     # odd leading dimension
    hw, hV = host[1:1 + n], host[n + 2:n + 2 + k * (n + 1)].reshape(k, n + 1)[:, :n]
    h = ctx.mdot(w, V, k)
    assert np.allclose(h, hV @ hw, rtol=0, atol=1e-11)
    assert abs(ctx.dot(w, V[0]) - hw @ hV[0]) < 1e-11
    assert abs(ctx.norm2(w) - np.linalg.norm(hw)) < 1e-12
    expect = hw - hV.T @ h
    ctx.maxpy(h, V, w)
    ctx.sync()
    assert np.linalg.norm(w.cpu().numpy() - expect) <= 1e-13 * np.linalg.norm(expect)


def test_operator_rejects_bad_input(env, orc):
    torch, ctx = env
    m = cdm.Mesh.cartesian(ctx, 3, 2)
    sp = cdm.H1Space(m, 2)
    with pytest.raises(cdm.CdmError):
        cdm.ConvectionDiffusionOperator(sp)                                  # no integrator
    with pytest.raises(cdm.CdmError):
        cdm.ConvectionDiffusionOperator(sp, kappa=1.0, ess_dofs=np.array([sp.ndof + 5], np.int32))
    with pytest.raises(cdm.CdmError):
        cdm.ConvectionDiffusionOperator(sp, kappa=np.ones(4))                # neither scalar nor symmetric 3x3
    op = cdm.ConvectionDiffusionOperator(sp, kappa=1.0)
    x = up(torch, np.ones(sp.ndof))
    with pytest.raises(cdm.CdmError):
        op.Mult(x, x)                                                        # aliasing
    with pytest.raises(cdm.CdmError):
        op.update(kappa=1.0, mass=1.0)                                       # integrator set must not change
    with pytest.raises(cdm.CdmError):
        op.set_option("no-such-option", 1)
