#!/usr/bin/env python
"""The reference's steady driver, input files included: `linear_convection_diffusion_2D <input.yaml>`
(linear_convection_diffusion_2D.cpp:238-446) on the B200 library.

    python examples/convdiff_from_yaml.py Input/input_2d.yaml        (run from the directory the YAML's paths refer to)

Same sequence as the app's main(): LoadParams (:62-127) -> PETSc options file (:268-282) -> Mesh(mesh_file, 1, 1) +
UniformRefinement (:290-298: serial_ref_levels + par_ref_levels uniform refinements, cdm_mesh_uniform_refine)
-> H1 space (:311-312) -> all boundary attributes essential (:319-322) -> Diffusion + Convection + Mass (:335-339) ->
DomainLFIntegrator(f) (:341-343) -> ProjectBdrCoefficient(u_exact) (:345-347) -> FormLinearSystem (:351) -> KSP solve
(:368-374) -> L2 errors with rules of order max(2, 2p+3) (:383-392) -> error CSV (:405-419) -> ParaView (:421-433).
Exit codes as in the reference: 0 ok, 1 usage, 2 bad input, 3 runtime failure.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv):
    if len(argv) != 2:
        print(f"usage: {argv[0]} <input.yaml>", file=sys.stderr)
        return 1
    cdm = importlib.import_module("continuum-mechanics-mfem_b200")
    try:
        if not os.path.exists(argv[1]):
            raise RuntimeError("YAML input file not found: " + argv[1])
        cfg = cdm.Config(argv[1])
        if not cfg.has("mesh_file") or not cfg.get("mesh_file"):
            raise RuntimeError("Missing required YAML key: mesh_file")
        p = dict(mesh_file=cfg.get("mesh_file"), order=cfg.get("order", 1, int), sref=cfg.get("serial_ref_levels", 0, int),
                 pref=cfg.get("par_ref_levels", 0, int), kappa=cfg.get("kappa", 0.1, float), s=cfg.get("s", 1.0, float),
                 n=cfg.get("mode_n", 3, int), m=cfg.get("mode_m", 3, int),
                 petsc=cfg.get("petsc_options_file", "Input/petsc.opts"), out=cfg.get("output_path", "ParaView"),
                 name=cfg.get("collection_name", "convection_diffusion_2D"), csv=cfg.get("error_csv", "error_history_2D.csv"),
                 save=cfg.get("save_paraview", True, bool))
        c = [cfg.get("cx", 1.0, float), cfg.get("cy", -2.0, float)]
        if cfg.has("convection"):
            c = cfg.get("convection", None, list)
            if len(c) != 2:
                raise RuntimeError("YAML key convection must be a sequence of exactly 2 values.")
        if p["order"] < 1:
            raise RuntimeError("order must be >= 1.")
        if p["kappa"] <= 0.0:
            raise RuntimeError("kappa must be > 0.")
        if p["n"] <= 0 or p["m"] <= 0:
            raise RuntimeError("mode_n and mode_m must be positive integers.")
        if p["sref"] < 0 or p["pref"] < 0:
            raise RuntimeError("serial_ref_levels and par_ref_levels must be >= 0.")
    except Exception as e:                                       # noqa: BLE001  (:255-262 -> exit code 2)
        print(e, file=sys.stderr)
        return 2
    try:
        import torch
        if os.path.exists(p["petsc"]):
            solver = cdm.GMRESSolver.from_petsc_options(p["petsc"])
        else:
            print(f"PETSc options file not found: {p['petsc']}. Proceeding without options file.", file=sys.stderr)
            solver = cdm.GMRESSolver(cdm.GMRES_PETSC, 30, 10000, 1e-5, 1e-50, pc="ilu")       # KSP / PC defaults
        if getattr(solver, "ksp_type", "gmres") != "gmres":
            raise RuntimeError("only -ksp_type gmres is wired into this driver")
        ctx = cdm.Context(0)
        mesh = cdm.Mesh.read_gmsh(ctx, p["mesh_file"], refine=True)
        if mesh.dim != 2:
            raise RuntimeError("The mesh must be 2D.")
        mesh = mesh.uniform_refine(p["sref"] + p["pref"])          # one rank: the serial and the parallel levels (:295-304)
        sp = cdm.H1Space(mesh, p["order"])
        print(f"Global true dofs: {sp.ndof}")
        _, _, _, battr = mesh.arrays()
        ess = sp.essential_dofs(np.ones(int(battr.max()), np.int32))
        kap, s, (cx, cy), n, m = p["kappa"], p["s"], c, p["n"], p["m"]
        an, am = n * np.pi, m * np.pi
        exact = lambda X: np.sin(an * X[..., 0]) * np.sin(am * X[..., 1])                       # (:159-170)
        forcing = lambda X: (kap * (an * an + am * am) * np.sin(an * X[..., 0]) * np.sin(am * X[..., 1])          # (:177-215)
                             + cx * an * np.cos(an * X[..., 0]) * np.sin(am * X[..., 1])
                             + cy * am * np.sin(an * X[..., 0]) * np.cos(am * X[..., 1])
                             + s * np.sin(an * X[..., 0]) * np.sin(am * X[..., 1]))
        a = cdm.ConvectionDiffusionOperator(sp, kappa=kap, vel=(cx, cy), mass=s, ess_dofs=ess)
        dev = lambda v: torch.from_numpy(np.ascontiguousarray(v, np.float64)).cuda()
        b = torch.zeros(sp.ndof, dtype=torch.float64, device="cuda")
        u = torch.zeros_like(b)
        X = torch.zeros_like(b)
        torch.cuda.synchronize()
        q_lf = p["order"] + 1                                     # DomainLFIntegrator: order 2p
        sp.domain_lf(forcing(sp.rule_coords(q_lf)), b, q1d=q_lf)
        Xd = sp.dof_coords()
        sp.project_dofs(ess, exact(Xd[ess]), u)
        a.EliminateRHS(u, b)
        solver.SetOperator(a)
        solver.Mult(b, X)
        ctx.sync()
        if not solver.GetConverged():
            raise RuntimeError(f"PETSc solver did not converge. Iterations={solver.GetNumIterations()}, residual={solver.GetFinalNorm()}")
        q_err = max(2, 2 * p["order"] + 3) // 2 + 1
        ex_q = exact(sp.rule_coords(q_err))
        abs_l2 = sp.l2_error(X, ex_q, q1d=q_err)
        exact_l2 = sp.l2_error(None, ex_q, q1d=q_err)
        rel_l2 = abs_l2 / exact_l2 if exact_l2 > 1e-14 else 0.0
        print(f"KSP iterations: {solver.GetNumIterations()}, final residual {solver.GetFinalNorm():.3e}")
        print(f"L2 error (absolute): {abs_l2:.16g}")
        print(f"L2 error (relative): {rel_l2:.16g}")
        os.makedirs(p["out"], exist_ok=True)
        with open(os.path.join(p["out"], p["csv"]), "w") as fh:
            fh.write("abs_l2,rel_l2\n%.16g,%.16g\n" % (abs_l2, rel_l2))
        if p["save"]:
            sp.write_paraview(p["out"], p["name"], {"u": X.cpu().numpy(), "u_exact": exact(Xd)}, cycle=0, time=0.0)
    except Exception as e:                                       # noqa: BLE001  (:435-442 -> exit code 3)
        print("Error:", e, file=sys.stderr)
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
