"""BASELINE config 3 semantics at test size: the backward-Euler loop of
linear_convection_diffusion_1D.cpp:375-400,537-572 --
    (M + dt C(beta) + (dt/Pe) K) c^{n+1} = M c^n,   Dirichlet (analytic, time dependent) on x = 0, 1 only,
one mass apply + one BC elimination + one GMRES(30)+Jacobi solve per step and Peclet block --
on the CUDA path vs the oracle running the app's assembled-matrix path."""
import math

import numpy as np
import pytest
from scipy.special import erfc

import cdm_b200 as cdm

pytestmark = pytest.mark.gpu


def exact_concentration(x, t, pe):
    """ExactConcentration (linear_convection_diffusion_1D.cpp:146-166)"""
    if t <= 0.0:
        return np.zeros_like(x)
    diff = t / pe
    root = math.sqrt(diff)
    a1, a2 = (x - t) / (2 * root), (x + t) / (2 * root)
    gauss = -((x - t) ** 2) / (4 * diff)
    t3 = 0.5 * (1 + pe * x + pe * t) * np.exp(np.minimum(pe * x, 700.0)) * erfc(a2)
    c = 0.5 * erfc(a1) + math.sqrt(t * pe / math.pi) * np.exp(gauss) - t3
    return np.where(np.isfinite(c), c, 0.0)


@pytest.mark.parametrize("dim,p,n", [(2, 2, (12, 3)), (3, 2, (6, 2, 2))])
def test_backward_euler_blocks_match_reference_path(orc, dim, p, n):
    import torch
    ctx = cdm.Context(0)
    dt, nsteps, peclets = 1e-2, 4, (1.0, 10.0, 100.0)
    beta = (1.0, 0.0, 0.0)[:dim]
    x_attrs = [2, 4] if dim == 2 else [3, 5]                       # right/left faces: x = 1, x = 0
    base = orc.Problem(dim, p, list(n), perturb=0.0, kappa=None, vel=None, mass=1.0, ess_attrs=x_attrs)
    mesh = cdm.Mesh.from_arrays(ctx, base.vx, base.ev, base.bv, base.battr)
    sp = cdm.H1Space(mesh, p)
    ess = sp.essential_dofs(base.marker)
    assert np.array_equal(ess, base.ess)
    X = sp.dof_coords()[:, 0]

    def dev(a):
        t = torch.from_numpy(np.ascontiguousarray(a, np.float64)).cuda()
        torch.cuda.synchronize()
        return t

    mass_form = cdm.ConvectionDiffusionOperator(sp, mass=1.0)       # mass_form (:375-378)
    M_ref = base.csr()
    for pe in peclets:
        # forms[k] = Mass + Convection(beta, dt) + Diffusion(dt/Pe) (:391-400)
        Pk = orc.Problem(dim, p, list(n), perturb=0.0, kappa=dt / pe, vel=beta, alpha=dt, mass=1.0, ess_attrs=x_attrs)
        form = cdm.ConvectionDiffusionOperator(sp, kappa=dt / pe, vel=beta, alpha=dt, mass=1.0, ess_dofs=ess)
        solver = cdm.GMRESSolver()
        solver.SetOperator(form)
        c_ref = np.zeros(base.ndof)                                  # c^0 = 0 (:402-407)
        c_dev = dev(c_ref)
        rhs = torch.zeros_like(c_dev)
        sol = torch.zeros_like(c_dev)
        for step in range(1, nsteps + 1):
            t = step * dt
            g = np.where(base.ess_mark, exact_concentration(X, t, pe), 0.0)
            # --- CUDA path: rhs = M c^n ; BCs ; eliminate ; solve
            mass_form.MultUnconstrained(c_dev, rhs)                  # (:544)
            gd = dev(g)
            form.EliminateRHS(gd, rhs)                               # FormLinearSystem (:547-548)
            solver.Mult(rhs, sol)                                    # (:553-566)
            ctx.sync()
            assert solver.GetConverged()
            c_dev.copy_(sol)
            torch.cuda.synchronize()
            # --- oracle: the assembled-matrix path the app executes
            b = M_ref.spmv(c_ref)
            A = Pk.csr()
            A.eliminate(base.ess_mark, g, b)
            c_new, info = A.op().gmres(b, dinv=1.0 / A.diag())
            assert info["converged"] and info["iters"] == solver.GetNumIterations()
            assert np.max(np.abs(solver.history - info["hist"]) / info["hist"][0]) < 1e-10
            c_ref = c_new
            err = np.linalg.norm(c_dev.cpu().numpy() - c_ref) / max(np.linalg.norm(c_ref), 1e-300)
            assert err < 1e-10, (pe, step, err)
        # the discrete solution tracks the analytic profile (sanity, discretisation-level)
        if pe <= 10.0 and dim == 2:
            assert np.max(np.abs(c_ref - exact_concentration(X, nsteps * dt, pe))) < 0.2
