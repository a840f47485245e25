"""The MFEM-shaped C++ shim (include/cdm_mfem_shim.hpp) and the re-hosted steady driver
(examples/convdiff_steady.cpp, mirroring linear_convection_diffusion_2D.cpp:238-446)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "convdiff_steady")


def _build():
    import cdm_b200 as cdm
    if not os.path.exists(cdm.LIB_PATH):
        cdm.build()
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, capture_output=True)


def test_example_builds_and_fails_loudly_without_gpu():
    _build()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([EXE, "2", "4", "2"], capture_output=True, text=True)
    assert r.returncode == 3                      # the app's runtime-failure exit code (:441)
    assert "no usable CUDA device" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("dim,n,order,tol", [(2, 16, 2, 2e-3), (3, 12, 3, 2e-3), (2, 8, 3, 5e-3)])
def test_steady_driver_converges_to_manufactured_solution(dim, n, order, tol):
    _build()
    r = subprocess.run([EXE, str(dim), str(n), str(order)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    its = int(re.search(r"GMRES iterations: (\d+)", r.stdout).group(1))
    err = float(re.search(r"L2 error: abs [0-9.e+-]+\s+rel ([0-9.e+-]+)", r.stdout).group(1))
    assert 0 < its < 500 and err < tol, r.stdout


def _oracle_heat(orc, dim, n, p, dt, t_final, alpha=0.1):
    """the same backward-Euler loop on the CPU oracle's assembled matrices (direct solve): final L2 error"""
    import numpy as np
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    Pm = orc.Problem(dim, p, n, perturb=0.0, kappa=None, vel=None, mass=1.0)
    Pl = orc.Problem(dim, p, n, perturb=0.0, kappa=alpha * dt, vel=None, mass=1.0)
    csr = lambda c: sp.csr_matrix((c.vals, c.colind, c.rowptr), shape=(c.n, c.n))
    M, A = csr(Pm.csr()), csr(Pl.csr())
    ess = Pm.ess_mark.astype(bool)
    free = ~ess
    X, xq, xe = Pm.coords(), Pm.rule_coords(p + 1), Pm.rule_coords(p + 2)
    r2 = lambda x: ((x - 0.5) ** 2).sum(-1)
    ex = lambda x, t: np.sin(t) * np.cos(2 * r2(x))
    f = lambda x, t: (np.cos(t) * np.cos(2 * r2(x))
                      - alpha * np.sin(t) * (-16 * r2(x) * np.cos(2 * r2(x)) - 4 * dim * np.sin(2 * r2(x))))
    u = np.zeros(Pm.ndof)
    lu = spla.splu(A[free][:, free].tocsc())
    err = 0.0
    for s in range(1, int(np.ceil(t_final / dt - 1e-12)) + 1):
        t = s * dt
        rhs = M @ u + Pm.domain_lf(f(xq, t), scale=dt)
        u[ess] = ex(X, t)[ess]
        b = rhs - A @ np.where(ess, u, 0.0)
        u[free] = lu.solve(b[free])
        err = Pm.l2_error(u, ex(xe, t))
    return err


@pytest.mark.gpu
@pytest.mark.parametrize("dim,n,order", [(2, 16, 2), (3, 8, 2)])
def test_transient_heat_driver_tracks_manufactured_solution(orc, dim, n, order):
    """examples/heat_mms_transient.cpp (diffusion_mms.cpp:286-470): backward Euler with per-step mass apply,
    linear form, boundary projection, elimination, GMRES and L2 error on the device, against the same loop
    on the oracle's assembled matrices; halving dt halves the error (first order in time)."""
    _build()
    exe = os.path.join(ROOT, "examples", "heat_mms_transient")
    finals = []
    for dt in (0.05, 0.025):
        r = subprocess.run([exe, str(dim), str(n), str(order), str(dt), "1.0"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        final = float(re.search(r"Final L2 error at t=[0-9.e+-]+: ([0-9.e+-]+)", r.stdout).group(1))
        want = _oracle_heat(orc, dim, n, order, dt, 1.0)
        assert abs(final - want) <= 5e-6 * want, (final, want)       # printed with 7 digits; GMRES rtol 1e-10 vs direct solve
        finals.append(final)
    assert 0.45 < finals[1] / finals[0] < 0.55, finals


@pytest.mark.gpu
@pytest.mark.parametrize("n,order", [(6, 3), (7, 2)])
def test_generic_operator_solvers_mass_list_recover_and_integrator_level(n, order):
    """examples/shim_generic_operator.cpp: Solver::SetOperator(const Operator &) + SetPreconditioner on a user-defined
    Operator reproduce the fused drivers; two MassIntegrators on one form (diffusion_mms_ale.cpp:1018-1021);
    RecoverFEMSolution (linear_convection_diffusion_2D.cpp:377); AddMultPA / AssembleDiagonalPA on E-vectors"""
    _build()
    exe = os.path.join(ROOT, "examples", "shim_generic_operator")
    r = subprocess.run([exe, str(n), str(order)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all shim checks passed" in r.stdout
