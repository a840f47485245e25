"""SURVEY.md 8f rank 3, remainder: triangle meshes (the meshes the reference ships, Mesh/unit_square.msh) on the device
as the assembled matrix, and ILU(0) (-pc_type bjacobi -sub_pc_type ilu, Input/petsc_circle.opts:6-8) against the oracle:
pattern bit-exact, values / products <= 1e-12, GMRES histories <= 1e-10, solutions <= 1e-10."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cdm_b200 as cdm
from test_gpu_parity import Dev, make, make_op, rel, torch, ctx  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(GOLD[:-len("/golden")])


def tri(ctx, name, p, **kw):
    from oracle import tri_oracle as T
    m = cdm.Mesh.read_gmsh(ctx, os.path.join(GOLD, name + ".msh"))
    vx, ev, bv, ba = m.arrays()
    return T.TriProblem(p, vx, ev, bv, ba, **kw), m, cdm.H1Space(m, p)


@pytest.mark.parametrize("mesh", ["square_tri", "disk_tri"])
@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_triangle_matrix_and_apply_match_the_numpy_restatement(torch, ctx, mesh, p):
    P, m, sp = tri(ctx, mesh, p)
    op = cdm.ConvectionDiffusionOperator(sp, kappa=P.kappa, vel=P.vel, alpha=P.alpha, mass=P.mass, ess_dofs=P.ess)
    rowptr, colind, vals = op.assemble_csr()
    A = P.csr()
    assert np.array_equal(rowptr, A.rowptr) and np.array_equal(colind, A.colind)
    assert np.max(np.abs(vals - A.vals)) <= 1e-12 * np.max(np.abs(A.vals))
    d = Dev(torch, ctx)
    x = np.random.default_rng(p).uniform(-1, 1, P.ndof)
    xd, yd = d.up(x), d.zeros(P.ndof)
    op.MultUnconstrained(xd, yd)
    assert rel(d.down(yd), A.spmv(x)) <= 1e-12
    # constrained: z = x, z[ess] = 0, y = A z, y[ess] = x[ess]
    z = np.where(P.ess_mark, 0.0, x)
    want = np.where(P.ess_mark, x, A.spmv(z))
    op.Mult(xd, yd)
    assert rel(d.down(yd), want) <= 1e-12
    # Jacobi diagonal
    dd = d.zeros(P.ndof)
    op.AssembleDiagonal(dd)
    assert rel(d.down(dd), np.where(P.ess_mark, 1.0, A.diag())) <= 1e-12


@pytest.mark.parametrize("p", [1, 3])
def test_triangle_forms_match(torch, ctx, p):
    P, m, sp = tri(ctx, "square_tri", p)
    d = Dev(torch, ctx)
    nq = p + 2
    xq = sp.rule_coords(nq)
    assert np.max(np.abs(xq.reshape(-1, 2) - P.rule_coords(nq).reshape(-1, 2))) < 1e-15
    f = np.sin(3 * xq[..., 0]) * np.cos(2 * xq[..., 1]) + 0.5
    bd = d.zeros(P.ndof)
    sp.domain_lf(f, bd, q1d=nq)
    assert rel(d.down(bd), P.domain_lf(f, nq)) <= 1e-13
    X = P.coords()
    u = np.sin(2 * X[:, 0]) * X[:, 1]
    uex = np.sin(2 * xq[..., 0]) * xq[..., 1]
    e_dev = sp.l2_error(d.up(u), uex, q1d=nq)
    assert abs(e_dev - P.l2_error(u, uex, nq)) <= 1e-12 * max(P.l2_error(None, uex, nq), 1e-300)
    assert abs(sp.l2_error(None, uex, q1d=nq) - P.l2_error(None, uex, nq)) <= 1e-13


def _system(P, rng):
    b0 = rng.uniform(-1, 1, P.ndof)
    g = np.where(P.ess_mark, rng.uniform(-1, 1, P.ndof), 0.0)
    return b0, g


@pytest.mark.parametrize("case", [("tri", "square_tri", 2), ("tri", "disk_tri", 3), ("quad", 2, 3, 7), ("hex", 3, 2, 4)])
def test_ilu0_factors_sweeps_and_gmres_history(torch, ctx, orc, case):
    """ILU(0) of the matrix the solver sees, its two sweeps and GMRES left-preconditioned with it, against the oracle's
    sequential IKJ factorisation (the matrix is the ConstrainedOperator one: unit diagonal on essential rows)"""
    if case[0] == "tri":
        P, m, sp = tri(ctx, case[1], case[2], kappa=0.02)
        op = cdm.ConvectionDiffusionOperator(sp, kappa=P.kappa, vel=P.vel, alpha=P.alpha, mass=P.mass, ess_dofs=P.ess)
    else:
        P, m, sp = make(ctx, orc, case[1], case[2], case[3], perturb=0.1, kappa=0.02)
        op = make_op(P, sp)
    A = P.csr()
    # the matrix the device factorises: rows / columns of essential dofs eliminated, one on the diagonal
    rows = np.repeat(np.arange(A.n), np.diff(A.rowptr))
    er, ec = P.ess_mark[rows].astype(bool), P.ess_mark[A.colind].astype(bool)
    cv = np.where(er, (rows == A.colind).astype(float), np.where(ec, 0.0, A.vals))       # A's pattern kept (explicit zeros)
    Ac = orc.CSR(A.rowptr.copy(), A.colind.copy(), cv)
    f = Ac.ilu0()
    d = Dev(torch, ctx)
    op.set_option("assembly", 1)
    rng = np.random.default_rng(11)
    r = rng.uniform(-1, 1, P.ndof)
    zd = d.zeros(P.ndof)
    op.ilu_apply(d.up(r), zd)
    assert rel(d.down(zd), f.solve(r)) <= 1e-12
    # the single-launch sweeps (default) and the one-launch-per-level sweeps do the same arithmetic: bit-identical
    assert op.get_option("ilu_sweep") == 1
    op.set_option("ilu_sweep", 0)
    z0 = d.zeros(P.ndof)
    op.ilu_apply(d.up(r), z0)
    assert np.array_equal(d.down(z0), d.down(zd))
    op.set_option("ilu_sweep", 1)
    lf, lb = op.ilu_levels()
    assert 1 < lf < P.ndof and 1 < lb < P.ndof
    # GMRES(30), left-preconditioned, zero initial guess
    b0, g = _system(P, rng)
    bd, gd, xd = d.up(b0), d.up(g), d.zeros(P.ndof)
    op.EliminateRHS(gd, bd)
    b1 = d.down(bd)
    x1, info = Ac.gmres_ilu(b1, f, restart=30, max_it=500)
    s = cdm.GMRESSolver(pc="ilu")
    s.SetOperator(op)
    s.Mult(bd, xd)
    assert info["converged"] and s.GetConverged() and s.GetNumIterations() == info["iters"]
    assert np.max(np.abs(s.history - info["hist"]) / info["hist"][0]) <= 1e-10
    assert rel(d.down(xd), x1) <= 1e-10
    # and it is the stronger preconditioner (why the reference selects it): fewer iterations than Jacobi
    sj = cdm.GMRESSolver(max_it=2000)
    sj.SetOperator(op)
    xj = d.zeros(P.ndof)
    sj.Mult(bd, xj)
    assert s.GetNumIterations() < sj.GetNumIterations()


def test_steady_driver_runs_the_shipped_style_input(torch, tmp_path):
    """examples/convdiff_from_yaml.py = linear_convection_diffusion_2D <input.yaml>: flat YAML + PETSc options + Gmsh
    triangles, MMS error at the discretisation level, error CSV and ParaView files written"""
    os.makedirs(tmp_path / "Input"); os.makedirs(tmp_path / "Mesh")
    import shutil
    shutil.copy(os.path.join(GOLD, "square_tri.msh"), tmp_path / "Mesh" / "unit_square.msh")
    (tmp_path / "Input" / "petsc.opts").write_text("# options\n-ksp_type gmres\n-ksp_rtol 1e-10\n-ksp_atol 1e-12\n-ksp_max_it 2000\n"
                                                   "-pc_type bjacobi\n-sub_ksp_type preonly\n-sub_pc_type ilu\n")
    errs = {}
    for order in (2, 3):
        (tmp_path / "Input" / "in.yaml").write_text(
            f"mesh_file: Mesh/unit_square.msh\norder: {order}\nserial_ref_levels: 0\npar_ref_levels: 0\nkappa: 0.1\ncx: 1.0\ncy: -2.0\n"
            "s: 1.0\nmode_n: 1\nmode_m: 1\npetsc_options_file: \"Input/petsc.opts\"\noutput_path: \"ParaView\"\n"
            "collection_name: \"cd2d\"\nerror_csv: \"err.csv\"\nsave_paraview: true\n")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "convdiff_from_yaml.py"), "Input/in.yaml"],
                           cwd=tmp_path, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        rows = (tmp_path / "ParaView" / "err.csv").read_text().split()
        assert rows[0] == "abs_l2,rel_l2"
        errs[order] = float(rows[1].split(",")[1])
        assert (tmp_path / "ParaView" / "cd2d" / "cd2d.pvd").exists()
        assert (tmp_path / "ParaView" / "cd2d" / "Cycle000000" / "proc000000.vtu").exists()
    assert errs[2] < 2e-3 and errs[3] < errs[2] / 4                                  # 162 triangles, h ~ 1/9
    # exit codes of the reference driver: usage 1, bad input 2
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "convdiff_from_yaml.py")], cwd=tmp_path, capture_output=True)
    assert r.returncode == 1
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "convdiff_from_yaml.py"), "nope.yaml"], cwd=tmp_path, capture_output=True)
    assert r.returncode == 2
