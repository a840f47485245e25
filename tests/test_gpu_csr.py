"""the reference's literal path on the device -- full assembly into CSR and SpMV as the
operator apply (linear_convection_diffusion_2D.cpp:339,351,368) -- against the oracle's assembled path:
pattern bit-exact, values and products <= 1e-12, GMRES history <= 1e-10."""
import numpy as np
import pytest

import cdm_b200 as cdm
from test_gpu_parity import Dev, make, make_op, rel, torch, ctx  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu
CASES = [(2, 1, 5), (2, 2, 4), (2, 4, 2), (3, 1, 3), (3, 2, 3), (3, 3, 2)]


@pytest.mark.parametrize("dim,p,n", CASES)
def test_assembled_matrix_matches_oracle(torch, ctx, orc, dim, p, n):
    P, mesh, sp = make(ctx, orc, dim, p, n, perturb=0.12, shuffle_seed=p)
    op = make_op(P, sp)
    rowptr, colind, vals = op.assemble_csr()
    A = P.csr()
    assert np.array_equal(rowptr, A.rowptr) and np.array_equal(colind, A.colind)          # index work: bit-exact
    assert np.max(np.abs(vals - A.vals)) <= 1e-12 * np.max(np.abs(A.vals))
    d = Dev(torch, ctx)
    x = np.random.default_rng(3).uniform(-1, 1, P.ndof)
    xd, yd = d.up(x), d.zeros(P.ndof)
    op.set_option("assembly", 1)
    op.MultUnconstrained(xd, yd)
    assert rel(d.down(yd), A.spmv(x)) <= 1e-12
    op.Mult(xd, yd)                                                                      # essential rows / columns eliminated
    assert rel(d.down(yd), P.pa_op(True).mult(x)) <= 1e-12
    # and it is the same operator as the matrix-free one
    op.set_option("assembly", 0)
    zd = d.zeros(P.ndof)
    op.Mult(xd, zd)
    assert rel(d.down(zd), d.down(yd)) <= 1e-12


@pytest.mark.parametrize("dim,p,n", [(2, 2, 6), (3, 2, 3)])
def test_gmres_on_the_assembled_matrix_follows_the_reference_path(torch, ctx, orc, dim, p, n):
    """FormLinearSystem on the assembled matrix + GMRES(30)/Jacobi: what the application executes"""
    P, mesh, sp = make(ctx, orc, dim, p, n, perturb=0.1)
    d = Dev(torch, ctx)
    rng = np.random.default_rng(5)
    b0 = rng.uniform(-1, 1, P.ndof)
    g = np.where(P.ess_mark, rng.uniform(-1, 1, P.ndof), 0.0)
    A = P.csr()
    b1 = b0.copy()
    A.eliminate(P.ess_mark, g, b1)
    x1, info = A.op().gmres(b1, dinv=1.0 / A.diag())
    op = make_op(P, sp)
    op.set_option("assembly", 1)
    bd, gd, xd = d.up(b0), d.up(g), d.zeros(P.ndof)
    op.EliminateRHS(gd, bd)
    # same right-hand side away from the essential dofs; there the app's matrix keeps its diagonal (b = a_ii g)
    # while ConstrainedOperator puts one (b = g) -- the Jacobi-preconditioned systems, hence the iterates, coincide
    free = ~P.ess_mark.astype(bool)
    assert rel(d.down(bd)[free], b1[free]) <= 1e-12
    assert np.array_equal(d.down(bd)[~free], g[~free])
    s = cdm.GMRESSolver()
    s.SetOperator(op)
    s.Mult(bd, xd)
    assert s.GetConverged() and s.GetNumIterations() == info["iters"]
    assert np.max(np.abs(s.history - info["hist"]) / info["hist"][0]) <= 1e-10
    assert rel(d.down(xd), x1) <= 1e-10


def test_update_refills_the_assembled_values(torch, ctx, orc):
    P, mesh, sp = make(ctx, orc, 3, 2, 3, perturb=0.1)
    d = Dev(torch, ctx)
    op = make_op(P, sp, constrained=False)
    op.set_option("assembly", 1)
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, P.ndof)
    xd, yd = d.up(x), d.zeros(P.ndof)
    kap, vel, mass = rng.uniform(0.5, 1.5, (P.ne, P.nq)), rng.uniform(-1, 1, (P.ne, P.nq, 3)), rng.uniform(0.5, 1.5, (P.ne, P.nq))
    op.update(kappa=kap, vel=vel, alpha=-1.0, mass=mass)
    P.set_coefficients(kap, vel, -1.0, mass)
    op.MultUnconstrained(xd, yd)
    assert rel(d.down(yd), P.csr().spmv(x)) <= 1e-12
    _, _, vals = op.assemble_csr()
    assert np.max(np.abs(vals - P.csr().vals)) <= 1e-12 * np.max(np.abs(P.csr().vals))


def test_sizes_before_and_after_assembly(torch, ctx, orc):
    P, mesh, sp = make(ctx, orc, 2, 1, 3)
    op = make_op(P, sp)
    import ctypes as C
    a, b = C.c_int64(0), C.c_int64(0)
    assert cdm.lib().cdm_operator_csr_sizes(op.h, C.byref(a), C.byref(b)) == cdm.EINVAL         # not assembled yet
    op.assemble_csr()
    assert cdm.lib().cdm_operator_csr_sizes(op.h, C.byref(a), C.byref(b)) == 0
    assert a.value == P.ndof and b.value == len(P.csr().colind)
