#!/usr/bin/env python
"""BASELINE config 3 in its own terms: backward-Euler convection-diffusion, 3D hex order 2, ~30 M dofs, on 1 GPU or
(under torchrun) on N GPUs with the mesh partitioned into boxes.
The time loop of linear_convection_diffusion_1D.cpp:537-572 for one Peclet block:
    (M + dt C(beta) + (dt/Pe) K) c^{n+1} = M c^n,   Dirichlet (analytic erfc profile) on x = 0, 1,
per step: mass apply, boundary projection, RHS elimination, GMRES(30)+Jacobi (rtol 1e-10, atol 1e-12), all on
the device.  Prints one JSON line: per-step time and its split, Krylov iterations, time per iteration.

  python scripts/transient_bench.py [--n 155] [--order 2] [--steps 4] [--pe 10] [--dt 1e-3]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P scripts/transient_bench.py --n 156
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
from scipy.special import erfc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import cdm_b200 as cdm  # noqa: E402


def exact_concentration(x, t, pe):
    """ExactConcentration (linear_convection_diffusion_1D.cpp:146-166)"""
    diff = t / pe
    root = math.sqrt(diff)
    a1, a2 = (x - t) / (2 * root), (x + t) / (2 * root)
    t3 = 0.5 * (1 + pe * x + pe * t) * np.exp(np.minimum(pe * x, 700.0)) * erfc(a2)
    c = 0.5 * erfc(a1) + math.sqrt(t * pe / math.pi) * np.exp(-((x - t) ** 2) / (4 * diff)) - t3
    return np.where(np.isfinite(c), c, 0.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", "--elems", dest="n", type=int, default=155)   # --elems: unambiguous under torch.distributed.run
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--pe", type=float, default=10.0)
    ap.add_argument("--dt", type=float, default=1e-3)
    ap.add_argument("--halo", type=int, default=2)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    ctx = cdm.Context(local_rank)
    parts = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_PROTO", "LL")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(cdm.Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_init(rank, world, uid.cpu().numpy().tobytes())
    t0 = time.perf_counter()
    mesh = cdm.Mesh.cartesian(ctx, 3, args.n, perturb=0.0)
    if world > 1:
        mesh = mesh.partition_box(parts, rank)
    sp = cdm.H1Space(mesh, args.order)
    marker = np.zeros(6, np.int32)
    marker[[2, 4]] = 1                                               # attributes 3, 5: x = 1 and x = 0 (:214-258)
    ess = sp.essential_dofs(marker)
    X = sp.dof_coords()[ess, 0]
    t_host = time.perf_counter() - t0
    beta, dt, pe = (1.0, 0.0, 0.0), args.dt, args.pe
    mass_form = cdm.ConvectionDiffusionOperator(sp, mass=1.0)                                   # (:375-378)
    form = cdm.ConvectionDiffusionOperator(sp, kappa=dt / pe, vel=beta, alpha=dt, mass=1.0, ess_dofs=ess)   # (:391-400)
    for o in (mass_form, form):
        o.set_option("tail", 1)                                      # vectors carry the ghost tail: no T<->L copies
        if world > 1:
            o.set_option("halo", args.halo)
    solver = cdm.GMRESSolver()                                       # Input/petsc.opts: gmres, 1e-10, 1e-12, 500, jacobi
    solver.SetOperator(form)
    n = sp.ndof
    c = torch.zeros(n, dtype=torch.float64, device="cuda")           # c^0 = 0 (:402-407)
    rhs, sol, g = torch.zeros_like(c), torch.zeros_like(c), torch.zeros_like(c)
    essd = torch.from_numpy(ess).cuda()
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream)
    its, t_step, t_solve = [], [], []
    for step in range(1, args.steps + 1):
        t = step * dt
        gv = exact_concentration(X, t, pe)                           # host evaluation of the Dirichlet data
        ctx.sync()
        w0 = time.perf_counter()
        mass_form.MultUnconstrained(c, rhs)                          # rhs = M c^n (:544)
        sp.project_dofs(essd, gv, g)                                 # ProjectBdrCoefficient (:545-546)
        form.EliminateRHS(g, rhs)                                    # FormLinearSystem (:547-548)
        solver.Mult(rhs, sol)                                        # (:553-566)
        ctx.sync()
        w1 = time.perf_counter()
        if not solver.GetConverged():
            raise SystemExit(f"solver did not converge at step {step}")
        with torch.cuda.stream(stream):
            c.copy_(sol)
        ctx.sync()
        its.append(solver.GetNumIterations())
        t_step.append((w1 - w0) * 1e3)
        t_solve.append(solver.res.seconds * 1e3)
    own = ess < sp.ntrue
    cx = c[essd].cpu().numpy()[own]
    X = X[own]
    tot = torch.tensor([sp.ntrue, sp.ne, int(own.sum())], dtype=torch.int64, device="cuda")
    red = torch.tensor([t_step, t_solve], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)               # every time is the max over the ranks
    t_step, t_solve = red[0].tolist(), red[1].tolist()
    n = int(tot[0])
    line = {"workload": f"backward Euler, 3D hex order {args.order}, {args.n}^3 elements, Pe={pe}, dt={dt}",
            "n_gpus": world, "partition": "x".join(map(str, parts)),
            "shared_dof_protocol": form.get_option("halo") if world > 1 else None,
            "dofs": n, "elements": int(tot[1]), "essential_dofs": int(tot[2]), "steps": args.steps,
            "gmres_iterations_per_step": its, "ms_per_step": [round(v, 3) for v in t_step],
            "solve_ms_per_step": [round(v, 3) for v in t_solve],
            "ms_per_krylov_iteration": round(sum(t_solve) / max(sum(its), 1), 4),
            "gdofs_per_krylov_iteration": round(n / (sum(t_solve) / max(sum(its), 1)) / 1e6, 3),
            "host_setup_s": round(t_host, 2),
            "bc_check": float(np.max(np.abs(cx - exact_concentration(X, args.steps * dt, pe))))}
    if world > 1:
        bc = torch.tensor([line["bc_check"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(bc, op=dist.ReduceOp.MAX)
        line["bc_check"] = float(bc[0])
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
