"""GPU parity AT SIZE: the CUDA path against the CPU oracle (and against the oracle-independent
Kronecker golden, tests/kron_golden.py) on meshes where every persistent warp / group processes
many elements, i.e. the steady-state loop of the default kernels is what gets compared:
ring reuse (parity ^= 1), mbarrier re-arming, the two-level gather prefetch.

  (a) BASELINE config 2 exactly (66^3 hexes, order 3, 7 880 599 dofs), default kernel, both scatter modes
  (b) a mid-size sweep for every kernel option and order 1..6 with odd element counts, plus the same
      kernels with the persistent grid capped to 3 blocks (option "grid_cap"): 10^2..10^3 elements per warp
  (c) BASELINE config 3 shape (order 2, Dirichlet on the x faces only) at 64^3
  (d) GMRES(30)+Jacobi history on a convection-dominated 32^3 order-3 problem that restarts, both scatter
      modes, with the drift of the atomic (red.add) mode reported
Tolerances: BASELINE.json north_star (apply 1e-12 relative L2, histories / solution 1e-10).
"""
import numpy as np
import pytest

import cdm_b200 as cdm
import kron_golden as kg

pytestmark = pytest.mark.gpu
APPLY_TOL = 1e-12
HIST_TOL = 1e-10
SOL_TOL = 1e-10


@pytest.fixture(scope="module")
def torch():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    return torch


@pytest.fixture(scope="module")
def ctx(torch):
    return cdm.Context(0)


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


class Dev:
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


# --------------------------------------------------------------------------------------- (a) config 2

@pytest.fixture(scope="module")
def config2(ctx, orc):
    P, mesh, sp = make(ctx, orc, 3, 3, 66, perturb=0.1, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0)
    assert P.ndof == 7880599 and P.ne == 287496
    return P, mesh, sp


def test_config2_index_maps_bit_exact(config2):
    P, mesh, sp = config2
    g, o, i = sp.maps()
    assert np.array_equal(g, P.elem_dof) and np.array_equal(o, P.offsets) and np.array_equal(i, P.indices)
    assert np.array_equal(sp.essential_dofs(np.ones(6, np.int32)), P.ess)


@pytest.mark.parametrize("scatter", [0, 1])
def test_config2_apply_matches_oracle(torch, ctx, config2, scatter):
    """66^3, order 3, Diffusion + Convection + Mass, constrained and unconstrained, DEFAULT kernel (3):
    2368 persistent warps x ~121 elements each, against the oracle's fused CPU PA apply"""
    P, mesh, sp = config2
    D = Dev(torch, ctx)
    x = np.random.default_rng(12345).uniform(-1, 1, P.ndof)
    op = make_op(P, sp)
    op.set_option("scatter", scatter)
    xd, yd = D.up(x), D.zeros(P.ndof)
    op.MultUnconstrained(xd, yd)
    y_ref = P.pa_apply_fast(x).copy()
    assert rel(D.down(yd), y_ref) < APPLY_TOL
    op.Mult(xd, yd)
    yc = D.down(yd)
    xz = np.where(P.ess_mark, 0.0, x)
    yc_ref = np.where(P.ess_mark, x, P.pa_apply_fast(xz))
    assert rel(yc, yc_ref) < APPLY_TOL
    assert np.array_equal(yc[P.ess], x[P.ess])
    # every other order-3 kernel option at the same size
    for kernel in (0, 1, 2, 4):
        op.set_option("kernel", kernel)
        op.Mult(xd, yd)
        assert rel(D.down(yd), yc_ref) < APPLY_TOL, kernel
    # the host-buffer entry point (pipelined upload / compute / download at this size)
    op.set_option("kernel", 3)
    assert rel(op.mult_host(x), yc_ref) < APPLY_TOL


def test_config2_diagonal_matches_oracle(torch, ctx, config2):
    P, mesh, sp = config2
    D = Dev(torch, ctx)
    op = make_op(P, sp)
    dd = D.zeros(P.ndof)
    op.AssembleDiagonal(dd)
    assert rel(D.down(dd), np.where(P.ess_mark, 1.0, P.pa_diag())) < APPLY_TOL


def test_config2_size_kronecker_golden(torch, ctx):
    """66^3 graded rectilinear cells, order 3: CUDA against the numpy Kronecker golden (no oracle involved)"""
    D = Dev(torch, ctx)
    n = [66, 66, 66]
    axes = [kg.graded_axis(k, 0.3 + 0.05 * d) for d, k in enumerate(n)]
    unit = cdm.Mesh.cartesian(ctx, 3, n, perturb=0.0)
    vx, ev, bv, battr = unit.arrays()
    mesh = cdm.Mesh.from_arrays(ctx, kg.rectilinear_vertices(vx, axes), ev, bv, battr)
    sp = cdm.H1Space(mesh, 3)
    K = kg.KronOperator(3, axes, kappa=0.1, vel=(1.0, -2.0, 0.5), alpha=1.0, mass=1.0)
    X = sp.dof_coords()
    x = np.random.default_rng(7).uniform(-1, 1, sp.ndof)
    y_ref = K.mult(x, X)
    op = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0)
    xd, yd = D.up(x), D.zeros(sp.ndof)
    for scatter in (0, 1):
        op.set_option("scatter", scatter)
        op.MultUnconstrained(xd, yd)
        assert rel(D.down(yd), y_ref) < APPLY_TOL
    dd = D.zeros(sp.ndof)
    op.AssembleDiagonal(dd)
    assert rel(D.down(dd), K.diag(X)) < APPLY_TOL


# ------------------------------------------------------------------------- (b) every kernel, every order

# (p, n): element counts chosen so that the persistent grids (148 SMs x resident blocks) loop >= 3 times
MID = {1: [33, 31, 32], 2: [27, 25, 26], 3: [21, 19, 20], 4: [17, 15, 16], 5: [15, 13, 14], 6: [13, 11, 12]}
KERNELS = {1: (0, 4, 5), 2: (0, 4, 5), 3: (0, 1, 2, 3, 4), 4: (0, 4), 5: (0, 4), 6: (0, 4)}


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_midsize_every_kernel_matches_oracle(torch, ctx, orc, p):
    P, mesh, sp = make(ctx, orc, 3, p, MID[p], perturb=0.12)
    D = Dev(torch, ctx)
    x = np.random.default_rng(100 + p).uniform(-1, 1, P.ndof)
    y_ref = P.pa_apply_fast(x).copy()
    xz = np.where(P.ess_mark, 0.0, x)
    yc_ref = np.where(P.ess_mark, x, P.pa_apply_fast(xz))
    op = make_op(P, sp)
    xd, yd = D.up(x), D.zeros(P.ndof)
    for kernel in KERNELS[p]:
        for scatter in (0, 1):
            op.set_option("kernel", kernel)
            op.set_option("scatter", scatter)
            op.MultUnconstrained(xd, yd)
            assert rel(D.down(yd), y_ref) < APPLY_TOL, (kernel, scatter)
            op.Mult(xd, yd)
            assert rel(D.down(yd), yc_ref) < APPLY_TOL, (kernel, scatter)


# (4, [5, 5, 5]): an odd element count, so that the last round of the order-4 kernel (three warps for two elements) holds a
# block with one valid and one idle element; (5, [5, 3, 3]) likewise for the packed order-5 variant
@pytest.mark.parametrize("p,n", [(1, [9, 8, 7]), (2, [7, 6, 7]), (3, [6, 7, 5]), (4, [5, 4, 5]), (4, [5, 5, 5]), (5, [4, 5, 3]),
                                 (5, [5, 3, 3]), (6, [3, 4, 3])])
@pytest.mark.parametrize("which", ["full", "mass", "diff+mass", "diff"])
def test_capped_grid_long_loops(torch, ctx, orc, p, n, which):
    """persistent grid capped to 3 blocks: every warp / group walks through tens of elements (including a
    ragged last round) -- integrator subsets, shuffled vertex numbering (all face orientations)"""
    kw = dict(full=dict(), mass=dict(kappa=None, vel=None, mass=1.3), diff=dict(kappa=0.3, vel=None, mass=None))
    kw["diff+mass"] = dict(kappa=0.3, vel=None, mass=2.0)
    P, mesh, sp = make(ctx, orc, 3, p, n, perturb=0.12, shuffle_seed=p, **kw[which])
    D = Dev(torch, ctx)
    x = np.random.default_rng(17).uniform(-1, 1, P.ndof)
    y_ref, yc_ref = P.pa_apply(x), P.pa_op(True).mult(x)
    xd, yd = D.up(x), D.zeros(P.ndof)
    op = make_op(P, sp)
    op.set_option("grid_cap", 3)
    for kernel in KERNELS[p]:
        if kernel == 0:
            continue                                  # the block kernel is not persistent
        for scatter in (0, 1):
            op.set_option("kernel", kernel)
            op.set_option("scatter", scatter)
            op.MultUnconstrained(xd, yd)
            assert rel(D.down(yd), y_ref) < APPLY_TOL, (kernel, scatter)
            op.Mult(xd, yd)
            assert rel(D.down(yd), yc_ref) < APPLY_TOL, (kernel, scatter)


@pytest.mark.parametrize("p", [3, 4, 5, 6])
def test_repeated_applies_are_stable(torch, ctx, orc, p):
    """The exchange stages of the high-order kernels reuse shared memory in place (orders 5, 6) and share warps between
    elements (order 4): a missing barrier would show up as run-to-run differences.  30 applies of the same input: with
    E-vector output (no atomics) every run must be bit-identical, with red.add output every run must match the oracle
    (compute-sanitizer is not available on the GPU pool, so the check is done by repetition)."""
    P, mesh, sp = make(ctx, orc, 3, p, MID[p], perturb=0.12)
    D = Dev(torch, ctx)
    x = np.random.default_rng(300 + p).uniform(-1, 1, P.ndof)
    y_ref = P.pa_apply_fast(x).copy()
    op = make_op(P, sp)
    xd, yd = D.up(x), D.zeros(P.ndof)
    op.set_option("scatter", 0)
    op.MultUnconstrained(xd, yd)
    first = D.down(yd).copy()
    assert rel(first, y_ref) < APPLY_TOL
    for _ in range(30):
        op.MultUnconstrained(xd, yd)
        assert np.array_equal(D.down(yd), first)
    op.set_option("scatter", 1)
    worst = 0.0
    for _ in range(30):
        op.MultUnconstrained(xd, yd)
        worst = max(worst, rel(D.down(yd), y_ref))
    assert worst < APPLY_TOL, worst


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5, 6])
def test_midsize_kronecker_golden_every_order(torch, ctx, p):
    """default kernels against the oracle-independent Kronecker golden on graded rectilinear meshes"""
    D = Dev(torch, ctx)
    n = MID[p]
    axes = [kg.graded_axis(k, 0.3 + 0.05 * d) for d, k in enumerate(n)]
    unit = cdm.Mesh.cartesian(ctx, 3, n, perturb=0.0)
    vx, ev, bv, battr = unit.arrays()
    mesh = cdm.Mesh.from_arrays(ctx, kg.rectilinear_vertices(vx, axes), ev, bv, battr)
    sp = cdm.H1Space(mesh, p)
    K = kg.KronOperator(p, axes, kappa=0.1, vel=(1.0, -2.0, 0.5), alpha=1.0, mass=1.0)
    X = sp.dof_coords()
    x = np.random.default_rng(7).uniform(-1, 1, sp.ndof)
    op = cdm.ConvectionDiffusionOperator(sp, kappa=0.1, vel=(1.0, -2.0, 0.5), mass=1.0)
    xd, yd = D.up(x), D.zeros(sp.ndof)
    op.MultUnconstrained(xd, yd)
    assert rel(D.down(yd), K.mult(x, X)) < APPLY_TOL
    dd = D.zeros(sp.ndof)
    op.AssembleDiagonal(dd)
    assert rel(D.down(dd), K.diag(X)) < APPLY_TOL


@pytest.mark.parametrize("p,n", [(1, 24), (2, 16), (3, 12), (4, 8)])
def test_2d_midsize(torch, ctx, orc, p, n):
    """2-D quads (BASELINE config 1 family) at sizes well beyond one wave of blocks"""
    P, mesh, sp = make(ctx, orc, 2, p, [16 * n + 1, 16 * n - 1], perturb=0.12, vel=(1.0, -2.0))
    D = Dev(torch, ctx)
    x = np.random.default_rng(3).uniform(-1, 1, P.ndof)
    op = make_op(P, sp)
    xd, yd = D.up(x), D.zeros(P.ndof)
    for scatter in (0, 1):
        op.set_option("scatter", scatter)
        op.MultUnconstrained(xd, yd)
        assert rel(D.down(yd), P.pa_apply(x)) < APPLY_TOL
        op.Mult(xd, yd)
        assert rel(D.down(yd), P.pa_op(True).mult(x)) < APPLY_TOL


# --------------------------------------------------------------------------------------- (c) config 3 shape

def test_config3_shape_order2_xface_bcs(torch, ctx, orc):
    """order 2, 64^3, backward-Euler operator M + dt C(beta) + (dt/Pe) K with Dirichlet data on the x faces
    only (linear_convection_diffusion_1D.cpp:391-400, :214-258): default sub-warp kernel at size"""
    dt, pe = 1e-3, 10.0
    P, mesh, sp = make(ctx, orc, 3, 2, 64, perturb=0.1, kappa=dt / pe, vel=(1.0, 0.0, 0.0), alpha=dt, mass=1.0,
                       ess_attrs=(3, 5))
    assert P.ndof == 129 ** 3
    m = np.zeros(6, np.int32); m[[2, 4]] = 1
    assert np.array_equal(sp.essential_dofs(m), P.ess)
    D = Dev(torch, ctx)
    x = np.random.default_rng(9).uniform(-1, 1, P.ndof)
    xz = np.where(P.ess_mark, 0.0, x)
    yc_ref = np.where(P.ess_mark, x, P.pa_apply_fast(xz))
    op = make_op(P, sp)
    xd, yd = D.up(x), D.zeros(P.ndof)
    for scatter in (0, 1):
        op.set_option("scatter", scatter)
        op.Mult(xd, yd)
        assert rel(D.down(yd), yc_ref) < APPLY_TOL
    # the mass form applied every step (mass_form.Mult(c, rhs), :544)
    Pm = orc.Problem(3, 2, 64, perturb=0.1, kappa=None, vel=None, mass=1.0)
    mass = cdm.ConvectionDiffusionOperator(sp, mass=1.0)
    mass.MultUnconstrained(xd, yd)
    assert rel(D.down(yd), Pm.pa_apply_fast(x)) < APPLY_TOL


# ------------------------------------------------------------------------------ (d) GMRES with restarts

def test_gmres_restarts_midsize_both_scatter_modes(torch, ctx, orc):
    """32^3 order 3 (912 673 dofs), kappa = 0.01 (convection dominated -> several GMRES(30) cycles):
    residual history of the CUDA solver against the oracle running the same algorithm on its PA operator"""
    P, mesh, sp = make(ctx, orc, 3, 3, 32, perturb=0.1, kappa=0.01, vel=(1.0, -2.0, 0.5), mass=1.0)
    D = Dev(torch, ctx)
    rng = np.random.default_rng(5)
    b0 = rng.uniform(-1, 1, P.ndof)
    g = np.where(P.ess_mark, rng.uniform(-1, 1, P.ndof), 0.0)
    ref_op = P.pa_op(True)
    b_ref = b0.copy()
    ref_op.eliminate_rhs(g, b_ref)
    d = np.where(P.ess_mark, 1.0, P.pa_diag())
    max_it = 100
    x_ref, info = ref_op.gmres(b_ref, dinv=1 / d, variant=0, restart=30, rtol=1e-10, atol=1e-12, max_it=max_it)
    assert info["iters"] > 60                                   # at least two restarts
    drift = {}
    for scatter in (0, 1):
        op = make_op(P, sp)
        op.set_option("scatter", scatter)
        bd, gd = D.up(b0), D.up(g)
        op.EliminateRHS(gd, bd)
        assert rel(D.down(bd), b_ref) < APPLY_TOL
        s = cdm.GMRESSolver(cdm.GMRES_PETSC, 30, max_it, 1e-10, 1e-12, jacobi=True)
        s.SetOperator(op)
        xd = D.zeros(P.ndof)
        s.Mult(bd, xd)
        assert s.GetConverged() == info["converged"] and abs(s.GetNumIterations() - info["iters"]) <= 1
        m = min(len(s.history), len(info["hist"]))
        drift[scatter] = float(np.max(np.abs(s.history[:m] - info["hist"][:m]) / info["hist"][0]))
        assert drift[scatter] < HIST_TOL, (scatter, drift)
        assert rel(D.down(xd), x_ref) < 1e-8
    print(f"\nGMRES history drift vs oracle over {info['iters']} iterations: deterministic scatter {drift[0]:.2e}, "
          f"fp64 red.add scatter {drift[1]:.2e}")


# ------------------------------------------------------------- integrator-level and prolongation entry points

@pytest.mark.parametrize("dim,p,n", [(3, 3, [9, 8, 7]), (3, 2, [8, 7, 9]), (3, 5, [4, 3, 4]), (2, 3, [20, 21])])
def test_integrator_level_evector_entry_points(torch, ctx, orc, dim, p, n):
    """BilinearFormIntegrator::AddMultPA / AssembleDiagonalPA on E-vectors, and ElementRestriction::Mult /
    MultTranspose: G^T (AddMultPA (G x)) must equal the oracle's apply, with y_E accumulated into"""
    P, mesh, sp = make(ctx, orc, dim, p, n, perturb=0.12, shuffle_seed=4, vel=(1.0, -2.0, 0.5)[:dim])
    D = Dev(torch, ctx)
    x = np.random.default_rng(21).uniform(-1, 1, P.ndof)
    op = make_op(P, sp, constrained=False)
    xd = D.up(x)
    xE, yE = D.zeros(P.ne * P.nd), D.zeros(P.ne * P.nd)
    sp.restrict(xd, xE)
    assert np.array_equal(D.down(xE), x[P.elem_dof.reshape(-1)])
    y0 = np.random.default_rng(22).uniform(-1, 1, P.ne * P.nd)
    yE.copy_(D.up(y0))
    op.AddMultPA(xE, yE)
    yL = D.zeros(P.ndof)
    sp.restrict_transpose(yE, yL)
    y0L = np.zeros(P.ndof)
    np.add.at(y0L, P.elem_dof.reshape(-1), y0)
    assert rel(D.down(yL) - y0L, P.pa_apply(x)) < 1e-11
    dE = D.zeros(P.ne * P.nd)
    op.AssembleDiagonalPA(dE)
    op.AssembleDiagonalPA(dE)                                   # accumulates
    dL = D.zeros(P.ndof)
    sp.restrict_transpose(dE, dL)
    assert rel(D.down(dL), 2.0 * P.pa_diag()) < APPLY_TOL


def test_two_contexts_one_process(torch, orc):
    """two cdm contexts in one process (one per device when the box has two, else two on device 0): every kernel
    instantiation is configured per context, results agree with the oracle on both"""
    ndev = torch.cuda.device_count()
    ctxs = [cdm.Context(0), cdm.Context(1 if ndev > 1 else 0)]
    P = orc.Problem(3, 3, [6, 5, 7], perturb=0.12)
    x = np.random.default_rng(2).uniform(-1, 1, P.ndof)
    y_ref = P.pa_apply(x)
    for i, c in enumerate(ctxs):
        dev = 1 if (i == 1 and ndev > 1) else 0
        with torch.cuda.device(dev):
            mesh = cdm.Mesh.from_arrays(c, P.vx, P.ev, P.bv, P.battr)
            sp = cdm.H1Space(mesh, 3)
            op = make_op(P, sp, constrained=False)
            xd = torch.from_numpy(x).to(f"cuda:{dev}")
            yd = torch.zeros_like(xd)
            torch.cuda.synchronize(dev)
            for kernel in (3, 4, 0):
                op.set_option("kernel", kernel)
                op.MultUnconstrained(xd, yd)
                c.sync()
                assert rel(yd.cpu().numpy(), y_ref) < APPLY_TOL
            del op, sp, mesh


# ------------------------------------------------------------ (e) 2D: the thread-per-element kernel (option 6)

@pytest.mark.parametrize("p", [1, 2, 3, 4])
@pytest.mark.parametrize("which", ["full", "mass", "diff+mass", "diff"])
def test_2d_thread_kernel_long_loops(torch, ctx, orc, p, which):
    """kernel option 6 (default in 2D): a warp owns 32 elements per round; persistent grid capped to 2 blocks so every
    warp re-arms its staging buffers several times and the last chunk is ragged (31 x 33 = 1023 elements); integrator
    subsets, shuffled vertex numbering, both scatter modes, against the oracle and against the block kernel"""
    kw = dict(full=dict(), mass=dict(kappa=None, vel=None, mass=1.3), diff=dict(kappa=0.3, vel=None, mass=None))
    kw["diff+mass"] = dict(kappa=0.3, vel=None, mass=2.0)
    P, mesh, sp = make(ctx, orc, 2, p, [31, 33], perturb=0.12, shuffle_seed=p, **kw[which])
    D = Dev(torch, ctx)
    x = np.random.default_rng(23).uniform(-1, 1, P.ndof)
    y_ref, yc_ref = P.pa_apply(x), P.pa_op(True).mult(x)
    xd, yd = D.up(x), D.zeros(P.ndof)
    op = make_op(P, sp)
    assert op.get_option("kernel") == 6
    for cap in (0, 2):
        op.set_option("grid_cap", cap)
        for scatter in (0, 1):
            op.set_option("scatter", scatter)
            op.MultUnconstrained(xd, yd)
            assert rel(D.down(yd), y_ref) < APPLY_TOL, (cap, scatter)
            op.Mult(xd, yd)
            assert rel(D.down(yd), yc_ref) < APPLY_TOL, (cap, scatter)
    op.set_option("kernel", 0)
    op.Mult(xd, yd)
    assert rel(D.down(yd), yc_ref) < APPLY_TOL


@pytest.mark.parametrize("p,n", [(2, 700), (3, 466)])
def test_2d_at_size_matches_oracle(torch, ctx, orc, p, n):
    """~2 M dofs in 2D (BASELINE config 1's element type at bench size): more chunks than resident warps, default kernel"""
    P, mesh, sp = make(ctx, orc, 2, p, n, perturb=0.1)
    D = Dev(torch, ctx)
    x = np.random.default_rng(5).uniform(-1, 1, P.ndof)
    xz = np.where(P.ess_mark, 0.0, x)
    yc_ref = np.where(P.ess_mark, x, P.pa_apply_fast(xz))
    op = make_op(P, sp)
    xd, yd = D.up(x), D.zeros(P.ndof)
    for scatter in (0, 1):
        op.set_option("scatter", scatter)
        op.Mult(xd, yd)
        assert rel(D.down(yd), yc_ref) < APPLY_TOL
