"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on
the same inputs.  Tolerances (BASELINE.json north_star): operator apply <= 1e-12
relative L2, Krylov residual histories <= 1e-10, final solution <= 1e-10 relative L2,
index maps bit-exact (checked on the CPU in test_host_abi.py and again here)."""
import json
import os

import numpy as np
import pytest

import cdm_b200 as cdm

pytestmark = pytest.mark.gpu
APPLY_TOL = 1e-12
HIST_TOL = 1e-10
SOL_TOL = 1e-10
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch


@pytest.fixture(scope="module")
def ctx(torch):
    c = cdm.Context(0)
    yield c


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


class Dev:
    """device vectors as torch tensors (PyTorch only owns the memory)"""

    def __init__(self, torch, ctx):
        self.t, self.ctx = torch, ctx

    def up(self, a):
        x = self.t.from_numpy(np.ascontiguousarray(a, np.float64)).cuda()
        self.t.cuda.synchronize()
        return x

    def zeros(self, n):
        x = self.t.zeros(n, dtype=self.t.float64, device="cuda")
        self.t.cuda.synchronize()
        return x

    def down(self, x):
        self.ctx.sync()
        return x.cpu().numpy()


def make(ctx, orc, dim, p, n, **kw):
    P = orc.Problem(dim, p, n, **kw)
    mesh = cdm.Mesh.from_arrays(ctx, P.vx, P.ev, P.bv, P.battr)
    sp = cdm.H1Space(mesh, p)
    return P, mesh, sp


def make_op(P, sp, constrained=True):
    return cdm.ConvectionDiffusionOperator(sp, kappa=P.kappa, vel=P.vel, alpha=P.alpha, mass=P.mass,
                                           ess_dofs=P.ess if constrained else None)


CASES = [(2, 1, 5), (2, 2, 4), (2, 3, 5), (2, 4, 3), (2, 5, 3), (2, 6, 2),
         (3, 1, 4), (3, 2, 4), (3, 3, 4), (3, 4, 3), (3, 5, 2), (3, 6, 2)]


@pytest.mark.parametrize("dim,p,n", CASES)
def test_index_maps_and_qdata(ctx, orc, dim, p, n):
    P, mesh, sp = make(ctx, orc, dim, p, n, perturb=0.12, shuffle_seed=p)
    g, o, i = sp.maps()
    assert np.array_equal(g, P.elem_dof) and np.array_equal(o, P.offsets) and np.array_equal(i, P.indices)
    op = make_op(P, sp)
    Dd, Dc, Dm = op.qdata()
    assert rel(Dd, P.Dd) < 1e-13 and rel(Dc, P.Dc) < 1e-13 and rel(Dm, P.Dm) < 1e-13


@pytest.mark.parametrize("scatter", [0, 1])
@pytest.mark.parametrize("dim,p,n", CASES)
def test_apply_matches_oracle(torch, ctx, orc, dim, p, n, scatter):
    """y = A x: CUDA vs oracle PA and vs oracle assembled CSR, seeded uniform[-1,1] input"""
    P, mesh, sp = make(ctx, orc, dim, p, n, perturb=0.12, shuffle_seed=100 + p)
    D = Dev(torch, ctx)
    rng = np.random.default_rng(12345)
    x = rng.uniform(-1, 1, P.ndof)
    op = make_op(P, sp)
    op.set_option("scatter", scatter)
    xd, yd = D.up(x), D.zeros(P.ndof)
    # unconstrained (BilinearForm::Mult)
    op.MultUnconstrained(xd, yd)
    y = D.down(yd)
    y_pa = P.pa_apply(x)
    assert rel(y, y_pa) < APPLY_TOL
    assert rel(y, P.csr().spmv(x)) < APPLY_TOL
    # constrained (ConstrainedOperator::Mult)
    op.Mult(xd, yd)
    yc = D.down(yd)
    assert rel(yc, P.pa_op(True).mult(x)) < APPLY_TOL
    assert np.array_equal(yc[P.ess], x[P.ess])
    # host-buffer entry point
    assert rel(op.mult_host(x), yc) < 1e-13


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_p3_kernel_variants(torch, ctx, orc, variant):
    """3D order 3 (the headline configuration): generic kernel and the bulk-async kernel"""
    P, mesh, sp = make(ctx, orc, 3, 3, 5, perturb=0.12)
    D = Dev(torch, ctx)
    x = np.sin(1.0 + 0.37 * np.arange(P.ndof))
    for scatter in (0, 1):
        op = make_op(P, sp)
        op.set_option("kernel", variant)
        op.set_option("scatter", scatter)
        xd, yd = D.up(x), D.zeros(P.ndof)
        op.MultUnconstrained(xd, yd)
        assert rel(D.down(yd), P.pa_apply(x)) < APPLY_TOL
        op.Mult(xd, yd)
        assert rel(D.down(yd), P.pa_op(True).mult(x)) < APPLY_TOL


@pytest.mark.parametrize("p,n", [(3, 24), (4, 21)])
def test_pipelined_host_apply(torch, ctx, p, n):
    """cdm_operator_mult_host on >= 8192 elements runs the chunked upload / compute / download pipeline:
    same result as the device-vector apply, with and without constraints, pinned or pageable host memory"""
    D = Dev(torch, ctx)
    mesh = cdm.Mesh.cartesian(ctx, 3, n, perturb=0.1)
    sp = cdm.H1Space(mesh, p)
    ess = sp.essential_dofs(np.ones(6, np.int32))
    op = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0, ess_dofs=ess)
    x = np.random.default_rng(4).uniform(-1, 1, sp.ndof)
    xp = torch.from_numpy(x).pin_memory()
    yp = torch.zeros(sp.ndof, dtype=torch.float64).pin_memory()
    xd, yd = D.up(x), D.zeros(sp.ndof)
    for constrained in (True, False):
        (op.Mult if constrained else op.MultUnconstrained)(xd, yd)
        y_dev = D.down(yd)
        for rep in range(2):
            y_pipe = op.mult_host(x, constrained=constrained)
        op.mult_host(xp.numpy(), yp.numpy(), constrained=constrained)
        op.set_option("host_pipeline", 0)
        y_plain = op.mult_host(x, constrained=constrained)
        op.set_option("host_pipeline", 1)
        assert rel(y_pipe, y_dev) < 1e-13 and rel(y_plain, y_dev) < 1e-13 and rel(yp.numpy(), y_dev) < 1e-13
        # both schedules (0: equal chunks, one copy per entity class and chunk; 1: tapered chunks, merged small classes)
        # and odd chunk counts
        for shape, K in ((0, 1), (0, 5), (1, 3), (1, 7)):
            op.set_option("host_pipeline_shape", shape)
            op.set_option("host_pipeline", K)
            assert rel(op.mult_host(x, constrained=constrained), y_dev) < 1e-13, (shape, K)
        op.set_option("host_pipeline_shape", 1)
        op.set_option("host_pipeline", 1)
        if constrained:
            assert np.array_equal(y_pipe[ess], x[ess]) and np.array_equal(yp.numpy()[ess], x[ess])


@pytest.mark.parametrize("p,n", [(1, 5), (2, 4), (3, 4), (4, 3), (5, 3), (6, 2)])
@pytest.mark.parametrize("which", ["full", "mass", "diff+mass", "diff"])
def test_group_kernel_all_orders(torch, ctx, orc, p, n, which):
    """kernel option 4: the register-z / bulk-async group kernel for every order, uneven element counts
    (exercises groups that run out of elements before their warp-mates)"""
    kw = dict(full=dict(), mass=dict(kappa=None, vel=None, mass=1.3), diff=dict(kappa=0.3, vel=None, mass=None))
    kw["diff+mass"] = dict(kappa=0.3, vel=None, mass=2.0)
    P, mesh, sp = make(ctx, orc, 3, p, [n, n + 1, n], perturb=0.12, shuffle_seed=p, **kw[which])
    D = Dev(torch, ctx)
    x = np.random.default_rng(11).uniform(-1, 1, P.ndof)
    for scatter in (0, 1):
        op = make_op(P, sp)
        op.set_option("kernel", 4)
        op.set_option("scatter", scatter)
        xd, yd = D.up(x), D.zeros(P.ndof)
        op.MultUnconstrained(xd, yd)
        assert rel(D.down(yd), P.pa_apply(x)) < APPLY_TOL
        op.Mult(xd, yd)
        assert rel(D.down(yd), P.pa_op(True).mult(x)) < APPLY_TOL


@pytest.mark.parametrize("p,n", [(1, 5), (1, 8), (2, 4), (2, 7)])
@pytest.mark.parametrize("which", ["full", "mass", "diff+mass", "diff"])
def test_subwarp_kernel_low_orders(torch, ctx, orc, p, n, which):
    """kernel option 5: two elements per warp (one per half warp) for orders 1 and 2; odd element counts
    leave the last warp with a single element"""
    kw = dict(full=dict(), mass=dict(kappa=None, vel=None, mass=1.3), diff=dict(kappa=0.3, vel=None, mass=None))
    kw["diff+mass"] = dict(kappa=0.3, vel=None, mass=2.0)
    P, mesh, sp = make(ctx, orc, 3, p, [n, n, n + 2], perturb=0.12, shuffle_seed=p, **kw[which])
    assert P.ne % 2 == 1 or n % 2 == 0
    D = Dev(torch, ctx)
    x = np.random.default_rng(13).uniform(-1, 1, P.ndof)
    for scatter in (0, 1):
        op = make_op(P, sp)
        op.set_option("kernel", 5)
        op.set_option("scatter", scatter)
        xd, yd = D.up(x), D.zeros(P.ndof)
        op.MultUnconstrained(xd, yd)
        assert rel(D.down(yd), P.pa_apply(x)) < APPLY_TOL
        op.Mult(xd, yd)
        assert rel(D.down(yd), P.pa_op(True).mult(x)) < APPLY_TOL


@pytest.mark.parametrize("dim,p,n", [(2, 2, 4), (3, 2, 3), (3, 3, 3)])
@pytest.mark.parametrize("which", ["mass", "diff", "diff+mass", "conv", "heat"])
def test_integrator_subsets(torch, ctx, orc, dim, p, n, which):
    """mass_form (Mass only, _1D.cpp:375-378), heat operator M + a dt K (diffusion_mms.cpp:301-305), ..."""
    kw = dict(mass=dict(kappa=None, vel=None, mass=1.0), diff=dict(kappa=0.3, vel=None, mass=None),
              conv=dict(kappa=None, vel=(1.0, 0.0, 0.0), mass=None, alpha=1e-3),
              heat=dict(kappa=0.1 * 0.05, vel=None, mass=1.0))
    kw["diff+mass"] = dict(kappa=0.3, vel=None, mass=2.0)
    P, mesh, sp = make(ctx, orc, dim, p, n, perturb=0.1, **kw[which])
    D = Dev(torch, ctx)
    x = np.random.default_rng(7).uniform(-1, 1, P.ndof)
    op = make_op(P, sp, constrained=False)
    xd, yd = D.up(x), D.zeros(P.ndof)
    op.MultUnconstrained(xd, yd)
    assert rel(D.down(yd), P.pa_apply(x)) < APPLY_TOL
    dd = D.zeros(P.ndof)
    op.AssembleDiagonal(dd)
    assert rel(D.down(dd), P.pa_diag()) < APPLY_TOL


@pytest.mark.parametrize("dim", [2, 3])
def test_variable_coefficients(torch, ctx, orc, dim):
    """per-point scalar/vector/symmetric-matrix coefficients (diffusion_mms_ale.cpp:1017-1023) and update()"""
    rng = np.random.default_rng(3)
    P, mesh, sp = make(ctx, orc, dim, 2, 3, perturb=0.1)
    D = Dev(torch, ctx)
    L = rng.uniform(-0.3, 0.3, (P.ne, P.nq, dim, dim)) + np.eye(dim)
    M = L @ np.swapaxes(L, -1, -2)
    pairs = [(0, 0), (1, 0), (1, 1)] if dim == 2 else [(0, 0), (1, 0), (2, 0), (1, 1), (2, 1), (2, 2)]
    kap = np.ascontiguousarray(np.stack([M[..., r, c] for r, c in pairs], axis=-1))
    vel = rng.uniform(-1, 1, (P.ne, P.nq, dim))
    mass = rng.uniform(0.5, 1.5, (P.ne, P.nq))
    op = make_op(P, sp, constrained=False)
    x = rng.uniform(-1, 1, P.ndof)
    xd, yd = D.up(x), D.zeros(P.ndof)
    op.update(kappa=kap, vel=vel, alpha=-1.0, mass=mass)
    P.set_coefficients(kap, vel, -1.0, mass)
    Dd, Dc, Dm = op.qdata()
    assert rel(Dd, P.Dd) < 1e-13 and rel(Dc, P.Dc) < 1e-13 and rel(Dm, P.Dm) < 1e-13
    op.MultUnconstrained(xd, yd)
    assert rel(D.down(yd), P.csr().spmv(x)) < APPLY_TOL
    # quadrature-point coordinates handed to Coefficient::Eval
    xq = sp.qpt_coords()
    assert xq.shape == (P.ne, P.nq, dim) and xq.min() > 0 and xq.max() < 1


@pytest.mark.parametrize("dim,p,n", [(2, 2, 4), (3, 2, 3), (3, 3, 3)])
def test_diagonal_and_eliminate_rhs(torch, ctx, orc, dim, p, n):
    P, mesh, sp = make(ctx, orc, dim, p, n, perturb=0.1, shuffle_seed=1)
    D = Dev(torch, ctx)
    op = make_op(P, sp)
    dd = D.zeros(P.ndof)
    op.AssembleDiagonal(dd)
    d_ref = np.where(P.ess_mark, 1.0, P.pa_diag())
    assert rel(D.down(dd), d_ref) < APPLY_TOL
    assert rel(D.down(dd), np.where(P.ess_mark, 1.0, P.csr().diag())) < APPLY_TOL
    rng = np.random.default_rng(5)
    b = rng.uniform(-1, 1, P.ndof)
    g = np.where(P.ess_mark, rng.uniform(-1, 1, P.ndof), 0.0)
    b_ref = b.copy()
    P.pa_op(True).eliminate_rhs(g, b_ref)
    bd, gd = D.up(b), D.up(g)
    op.EliminateRHS(gd, bd)
    assert rel(D.down(bd), b_ref) < APPLY_TOL


def test_vector_kernels(torch, ctx):
    D = Dev(torch, ctx)
    rng = np.random.default_rng(1)
    n, k = 100003, 11
    w = rng.uniform(-1, 1, n)
    V = rng.uniform(-1, 1, (k, n))
    wd, Vd = D.up(w), D.up(V)
    assert abs(ctx.dot(wd, Vd[0]) - w @ V[0]) < 1e-12 * n
    assert abs(ctx.norm2(wd) - np.linalg.norm(w)) < 1e-12 * np.sqrt(n)
    h = ctx.mdot(wd, Vd, k)
    assert np.allclose(h, V @ w, rtol=0, atol=1e-10)
    assert np.array_equal(h, ctx.mdot(wd, Vd, k))              # bit-reproducible
    ctx.maxpy(h, Vd, wd)
    assert rel(D.down(wd), w - V.T @ h) < 1e-13
    y = D.up(w)
    ctx.axpy(-0.5, Vd[1], y)
    assert rel(D.down(y), w - 0.5 * V[1]) < 1e-15
    z = D.zeros(n)
    ctx.add(Vd[2], 2.0, Vd[3], z)
    assert rel(D.down(z), V[2] + 2 * V[3]) < 1e-15


def solve_setup(P, D, op):
    rng = np.random.default_rng(5)
    b0 = rng.uniform(-1, 1, P.ndof)
    g = np.where(P.ess_mark, rng.uniform(-1, 1, P.ndof), 0.0)
    bd, gd = D.up(b0), D.up(g)
    op.EliminateRHS(gd, bd)
    return b0, g, bd


@pytest.mark.parametrize("dim,p,n", [(2, 2, 4), (2, 3, 6), (3, 2, 4), (3, 3, 4)])
def test_gmres_jacobi_history_matches_reference_path(torch, ctx, orc, dim, p, n):
    """GMRES(30)/CGS + Jacobi, rtol 1e-10, atol 1e-12, x0 = 0 (Input/petsc.opts:2-6) on the
    constrained PA operator vs the oracle running the app's path (assembled matrix,
    FormLinearSystem elimination)."""
    P, mesh, sp = make(ctx, orc, dim, p, n, perturb=0.1)
    D = Dev(torch, ctx)
    op = make_op(P, sp)
    b0, g, bd = solve_setup(P, D, op)
    A = P.csr()
    b1 = b0.copy()
    A.eliminate(P.ess_mark, g, b1)
    x_ref, info = A.op().gmres(b1, dinv=1 / A.diag(), variant=0, rtol=1e-10, atol=1e-12, max_it=500)
    assert info["converged"]
    s = cdm.GMRESSolver(cdm.GMRES_PETSC, 0, 500, 1e-10, 1e-12, jacobi=True)
    s.SetOperator(op)
    xd = D.zeros(P.ndof)
    s.Mult(bd, xd)
    assert s.GetConverged() and s.GetNumIterations() == info["iters"]
    h, hr = s.history, info["hist"]
    assert len(h) == len(hr)
    assert np.max(np.abs(h - hr) / hr[0]) < HIST_TOL
    assert rel(D.down(xd), x_ref) < SOL_TOL
    assert abs(s.GetFinalNorm() - info["final_norm"]) < HIST_TOL * hr[0]


def test_gmres_mfem_variant_and_restart(torch, ctx, orc):
    P, mesh, sp = make(ctx, orc, 3, 2, 4, perturb=0.1, kappa=0.01)     # convection dominated: needs restarts
    D = Dev(torch, ctx)
    op = make_op(P, sp)
    b0, g, bd = solve_setup(P, D, op)
    ref_op = P.pa_op(True)
    b2 = b0.copy()
    ref_op.eliminate_rhs(g, b2)
    d = np.where(P.ess_mark, 1.0, P.pa_diag())
    for variant, restart in ((cdm.GMRES_MFEM, 0), (cdm.GMRES_PETSC, 5), (cdm.GMRES_MFEM, 7)):
        x_ref, info = ref_op.gmres(b2, dinv=1 / d, variant=variant, restart=restart, rtol=1e-10, atol=1e-12)
        s = cdm.GMRESSolver(variant, restart, 500, 1e-10, 1e-12, jacobi=True)
        s.SetOperator(op)
        xd = D.zeros(P.ndof)
        s.Mult(bd, xd)
        assert s.GetConverged() == info["converged"] and abs(s.GetNumIterations() - info["iters"]) <= 1
        m = min(len(s.history), len(info["hist"]))
        assert np.max(np.abs(s.history[:m] - info["hist"][:m]) / info["hist"][0]) < HIST_TOL
        assert rel(D.down(xd), x_ref) < 1e-8


def test_cg_history_matches_mfem_cgsolver(torch, ctx, orc):
    """mfem::CGSolver rel 1e-12, abs 0, max 500, no preconditioner (mesh_recession_handler.cpp:270-276)"""
    P, mesh, sp = make(ctx, orc, 3, 2, 4, perturb=0.1, kappa=1.0, vel=None, mass=None)
    D = Dev(torch, ctx)
    op = make_op(P, sp)
    b0, g, bd = solve_setup(P, D, op)
    ref_op = P.pa_op(True)
    b2 = b0.copy()
    ref_op.eliminate_rhs(g, b2)
    x_ref, info = ref_op.cg(b2, rtol=1e-12, atol=0.0, max_it=500)
    s = cdm.CGSolver(500, 1e-12, 0.0, jacobi=False)
    s.SetOperator(op)
    xd = D.zeros(P.ndof)
    s.Mult(bd, xd)
    assert s.GetConverged() and info["converged"] and abs(s.GetNumIterations() - info["iters"]) <= 1
    m = min(len(s.history), len(info["hist"]))
    assert np.max(np.abs(s.history[:m] - info["hist"][:m]) / info["hist"][0]) < HIST_TOL
    assert rel(D.down(xd), x_ref) < SOL_TOL
    # Jacobi-preconditioned variant
    d = np.where(P.ess_mark, 1.0, P.pa_diag())
    x_ref, info = ref_op.cg(b2, dinv=1 / d, rtol=1e-12, atol=0.0, max_it=500)
    s = cdm.CGSolver(500, 1e-12, 0.0, jacobi=True)
    s.SetOperator(op)
    s.Mult(bd, xd)
    assert s.GetConverged() and abs(s.GetNumIterations() - info["iters"]) <= 1
    assert rel(D.down(xd), x_ref) < SOL_TOL


def test_config1_as_shipped_2d_order2(torch, ctx, orc):
    """BASELINE config 1: 2D quads, H1 order 2, kappa=0.1, c=(1,-2), s=1, all-Dirichlet MMS solve
    (linear_convection_diffusion_2D.cpp:311-377): 81 dofs on a 4x4 mesh."""
    P, mesh, sp = make(ctx, orc, 2, 2, 4, perturb=0.0, kappa=0.1, vel=(1.0, -2.0), mass=1.0)
    assert P.ndof == 81
    D = Dev(torch, ctx)
    X = sp.dof_coords()
    uex = np.sin(3 * np.pi * X[:, 0]) * np.sin(3 * np.pi * X[:, 1])
    xq = sp.qpt_coords()
    sx, cx = np.sin(3 * np.pi * xq[..., 0]), np.cos(3 * np.pi * xq[..., 0])
    sy, cy = np.sin(3 * np.pi * xq[..., 1]), np.cos(3 * np.pi * xq[..., 1])
    f = 0.1 * 18 * np.pi ** 2 * sx * sy + 3 * np.pi * cx * sy - 2 * 3 * np.pi * sx * cy + sx * sy
    lf = cdm.ConvectionDiffusionOperator(sp, mass=f)                  # DomainLFIntegrator(f) as a weighted mass form
    one, bd = D.up(np.ones(P.ndof)), D.zeros(P.ndof)
    lf.MultUnconstrained(one, bd)
    op = make_op(P, sp)
    gd = D.up(np.where(P.ess_mark, uex, 0.0))
    op.EliminateRHS(gd, bd)
    s = cdm.GMRESSolver()
    s.SetOperator(op)
    xd = D.zeros(P.ndof)
    s.Mult(bd, xd)
    assert s.GetConverged()
    # same solve on the oracle's assembled path
    Pf = orc.Problem(2, 2, 4, perturb=0.0, kappa=None, vel=None, mass=f)
    b = Pf.pa_apply(np.ones(P.ndof))
    A = P.csr()
    A.eliminate(P.ess_mark, np.where(P.ess_mark, uex, 0.0), b)
    x_ref, info = A.op().gmres(b, dinv=1 / A.diag())
    assert info["iters"] == s.GetNumIterations()
    assert rel(D.down(xd), x_ref) < SOL_TOL
    assert np.max(np.abs(s.history - info["hist"]) / info["hist"][0]) < HIST_TOL


def test_golden_fixture(torch, ctx, orc):
    with open(os.path.join(GOLD, "oracle_small.json")) as fh:
        gold = json.load(fh)
    D = Dev(torch, ctx)
    for case in gold["cases"]:
        m = cdm.Mesh.cartesian(ctx, case["dim"], case["n"], perturb=case["perturb"])
        sp = cdm.H1Space(m, case["p"])
        assert sp.ndof == case["ndof"]
        g, _, _ = sp.maps()
        assert np.array_equal(g[:2].reshape(-1), np.array(case["elem_dof_head"], np.int32))
        op = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0)
        x = np.sin(1.0 + 0.37 * np.arange(sp.ndof))
        xd, yd = D.up(x), D.zeros(sp.ndof)
        op.MultUnconstrained(xd, yd)
        y = D.down(yd)
        assert abs(np.linalg.norm(y) - case["y_norm"]) <= 1e-12 * case["y_norm"]
        assert np.allclose(y[:8], case["y_head"], rtol=1e-11, atol=1e-13)
        dd = D.zeros(sp.ndof)
        op.AssembleDiagonal(dd)
        assert abs(D.down(dd).sum() - case["diag_sum"]) <= 1e-12 * abs(case["diag_sum"])
        # linear form / L2 error (default rules) against the committed vectors
        fn = lambda c: 1.0 + np.sin(2.3 * c[..., 0]) * np.cos(1.7 * c[..., 1]) + 0.5 * c[..., -1] ** 2
        lf = D.zeros(sp.ndof)
        sp.domain_lf(fn(sp.rule_coords(case["p"] + 1)), lf)
        lf = D.down(lf)
        assert abs(np.linalg.norm(lf) - case["lf_norm"]) <= 1e-12 * case["lf_norm"]
        assert np.allclose(lf[:8], case["lf_head"], rtol=1e-11, atol=1e-14)
        assert abs(sp.l2_error(xd, fn(sp.rule_coords(case["p"] + 2))) - case["l2_error"]) <= 1e-12 * case["l2_error"]


@pytest.mark.parametrize("kernel", [0, 1, 2, 3])
def test_full_size_properties_config2(torch, ctx, kernel):
    """BASELINE config 2 size (66^3 hexes, order 3, 7 880 599 dofs): size-independent properties,
    no oracle needed: K 1 = 0, C 1 = 0, 1^T M 1 = volume, symmetry of K + M, linearity."""
    D = Dev(torch, ctx)
    m = cdm.Mesh.cartesian(ctx, 3, 66, perturb=0.1)
    sp = cdm.H1Space(m, 3)
    assert sp.ndof == 199 ** 3 == 7880599
    n = sp.ndof
    one = D.up(np.ones(n))
    y = D.zeros(n)
    rng = np.random.default_rng(0)
    u, v = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    ud, vd = D.up(u), D.up(v)
    for scatter in (0, 1):
        kc = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5))
        kc.set_option("kernel", kernel); kc.set_option("scatter", scatter)
        kc.MultUnconstrained(one, y)
        scale = ctx.norm2(ud)
        kc.MultUnconstrained(ud, D.zeros(n))
        assert ctx.norm2(y) < 1e-10 * scale                        # (K + C) 1 = 0
        del kc
        ms = cdm.ConvectionDiffusionOperator(sp, mass=1.0)
        ms.set_option("kernel", kernel); ms.set_option("scatter", scatter)
        ms.MultUnconstrained(one, y)
        assert abs(ctx.dot(one, y) - 1.0) < 1e-12                  # volume of the unit cube
        del ms
        km = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, mass=1.0)
        km.set_option("kernel", kernel); km.set_option("scatter", scatter)
        yu, yv = D.zeros(n), D.zeros(n)
        km.MultUnconstrained(ud, yu)
        km.MultUnconstrained(vd, yv)
        a, b = ctx.dot(vd, yu), ctx.dot(ud, yv)
        assert abs(a - b) < 1e-11 * max(abs(a), ctx.norm2(yu))      # symmetry
        w = D.up(2.0 * u - 3.0 * v)
        yw = D.zeros(n)
        km.MultUnconstrained(w, yw)
        ctx.axpy(-2.0, yu, yw); ctx.axpy(3.0, yv, yw)
        assert ctx.norm2(yw) < 1e-12 * ctx.norm2(yu)               # linearity
        del km
