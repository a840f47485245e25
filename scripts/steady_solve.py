#!/usr/bin/env python
"""BASELINE config 2 as a SOLVE (not just an iteration time): the steady manufactured-solution problem of
linear_convection_diffusion_2D.cpp:311-392 extended to 3D hexes, solved with the two option files the reference ships:

  Input/petsc.opts         GMRES(30) + Jacobi, max_it 500      (matrix-free apply; the assembled apply for comparison)
  Input/petsc_circle.opts  GMRES(30) + block-Jacobi/ILU(0), max_it 2000   (assembled matrix; single-launch and per-level sweeps)

  python scripts/steady_solve.py [--n 32] [--order 3] [--max-it 2000] [--skip-levels]

One JSON line per solver: iterations, converged, final preconditioned residual, seconds, ms per iteration, relative L2 error.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cdm_b200 as cdm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=32)
ap.add_argument("--order", type=int, default=3)
ap.add_argument("--max-it", type=int, default=2000)
ap.add_argument("--skip-levels", action="store_true")
ap.add_argument("--skip-assembled", action="store_true")
args = ap.parse_args()
KAPPA, VEL, MASS = 0.1, (1.0, -2.0, 0.5), 1.0

ctx = cdm.Context(0)
mesh = cdm.Mesh.cartesian(ctx, 3, args.n, perturb=0.1)
sp = cdm.H1Space(mesh, args.order)
ess = sp.essential_dofs(np.ones(6, np.int32))
pi = np.pi
exact = lambda X: np.sin(pi * X[..., 0]) * np.sin(pi * X[..., 1]) * np.sin(pi * X[..., 2])


def forcing(X):
    sx, sy, sz = (np.sin(pi * X[..., i]) for i in range(3))
    cx, cy, cz = (np.cos(pi * X[..., i]) for i in range(3))
    return (KAPPA * 3 * pi * pi + MASS) * sx * sy * sz + pi * (VEL[0] * cx * sy * sz + VEL[1] * sx * cy * sz + VEL[2] * sx * sy * cz)


op = cdm.ConvectionDiffusionOperator(sp, kappa=KAPPA, vel=VEL, mass=MASS, ess_dofs=ess)
b = torch.zeros(sp.ndof, dtype=torch.float64, device="cuda")
u = torch.zeros_like(b)
torch.cuda.synchronize()
q_lf = args.order + 2
sp.domain_lf(forcing(sp.rule_coords(q_lf)), b, q1d=q_lf)
Xd = sp.dof_coords()
sp.project_dofs(ess, exact(Xd[ess]), u)
op.EliminateRHS(u, b)
q_err = args.order + 3
ex_q = exact(sp.rule_coords(q_err))
ex_norm = sp.l2_error(None, ex_q, q1d=q_err)
print(json.dumps({"order": args.order, "n": args.n, "dofs": sp.ndof, "elements": sp.ne}), flush=True)


def run(name, pc, assembly, max_it, sweep=None):
    t0 = time.perf_counter()
    op.set_option("assembly", assembly)
    if sweep is not None:
        op.set_option("ilu_sweep", sweep)
    s = cdm.GMRESSolver(cdm.GMRES_PETSC, 30, max_it, 1e-10, 1e-12, pc=pc)
    s.SetOperator(op)
    x = torch.zeros_like(b)
    torch.cuda.synchronize()
    if pc == "ilu":
        op.ilu_levels()                                        # factorise outside the solve timing
    ctx.sync()
    t_setup = time.perf_counter() - t0
    l0 = ctx.launches
    t0 = time.perf_counter()
    s.Mult(b, x)
    ctx.sync()
    dt = time.perf_counter() - t0
    l1 = ctx.launches
    it = s.GetNumIterations()
    rec = {"solver": name, "iters": it, "converged": bool(s.GetConverged()), "final_norm": s.GetFinalNorm(), "max_it": max_it,
           "solve_s": round(dt, 4), "ms_per_iter": round(1e3 * dt / max(it, 1), 4), "setup_s": round(t_setup, 3),
           "launches": l1 - l0, "rel_l2_error": sp.l2_error(x, ex_q, q1d=q_err) / ex_norm}
    if pc == "ilu":
        rec["ilu_levels"] = op.ilu_levels()
    print(json.dumps(rec), flush=True)


run("gmres30+jacobi, matrix-free (Input/petsc.opts)", "jacobi", 0, 500)
run("gmres30+jacobi, matrix-free, max_it as petsc_circle.opts", "jacobi", 0, args.max_it)
if not args.skip_assembled:
    run("gmres30+jacobi, assembled SpMV", "jacobi", 1, args.max_it)
    run("gmres30+bjacobi/ilu0, assembled, single-launch sweeps (Input/petsc_circle.opts)", "ilu", 1, args.max_it, sweep=1)
    if not args.skip_levels:
        run("gmres30+bjacobi/ilu0, assembled, one launch per level", "ilu", 1, args.max_it, sweep=0)
